"""CPU tests of the KERNEL BODIES: csrc/rbd_sweeps.h compiled by g++ (tests/emul, test-only shim) against the C oracle.
Same templates, same stack/branch-slot layout, same tiled addressing as the sm_100a kernels; shared memory is a
poisoned heap array so that a read-before-write shows up as 1e300 garbage."""
import ctypes as C

import numpy as np
import pytest

import helpers
from helpers import P, TOL_FP64, from_tiled, rel_err, to_tiled

MODELS = helpers.all_models()
IDS = [m.name for m in MODELS]
MODELS_MAP = MODELS + helpers.subset_models()          # + permuted strict subsets of the frames as mapped outputs
IDS_MAP = [m.name for m in MODELS_MAP]


def _cdesc(d):
    from sofa.rigidbodydynamics_b200._capi import make_desc
    return make_desc(d)


def _emul_eval(emul, d, q, v, a, tile, block, want, use_tmem=1):
    n = q.shape[0]
    cd, keep = _cdesc(d)
    nt = (n + tile - 1) // tile
    K = {"oMi": 12 * (d.njoints - 1), "J": 6 * d.nv, "M": d.nv * (d.nv + 1) // 2, "tau": d.nv}
    outs = {k: (np.full((nt, K[k], tile), np.nan) if k in want else None) for k in K}
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    rc = emul.emul_eval(C.addressof(cd), n, tile, block, use_tmem, P(qs), P(vs), P(as_), P(outs["oMi"]), P(outs["J"]),
                        P(outs["M"]), P(outs["tau"]))
    assert rc == 0, emul.emul_last_error()
    return {k: (from_tiled(o, n) if o is not None else None) for k, o in outs.items()}


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile,block,use_tmem", [(32, 0, 1), (32, 64, 0), (7, 32, 1), (100, 96, 0), (32, 256, 1)])
def test_eval_bodies_match_oracle(emul, d, tile, block, use_tmem):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n = 70
    q, v, a = mdl.random_state(d, n, seed=3)
    if d.name == "talos_reduced" and block == 256 and use_tmem:
        pytest.skip("state of 256 Talos instances exceeds one SM (planner refuses: covered by test_stack_budget)")
    got = _emul_eval(emul, d, q, v, a, tile, block, ("oMi", "J", "M", "tau"), use_tmem)
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    for k in ref:
        assert rel_err(got[k], ref[k]) <= TOL_FP64, k


@pytest.mark.parametrize("want", [("M",), ("tau",), ("oMi", "J"), ("J", "tau"), ("oMi", "M"), ("oMi",)])
@pytest.mark.parametrize("name", ["talos_reduced", "random_3", "random_2"])
@pytest.mark.parametrize("use_tmem", [0, 1])
def test_eval_partial_outputs(emul, name, want, use_tmem):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = [m for m in MODELS if m.name == name][0]
    q, v, a = mdl.random_state(d, 40, seed=4)
    got = _emul_eval(emul, d, q, v, a, 32, 0, want, use_tmem)
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    for k in want:
        assert rel_err(got[k], ref[k]) <= TOL_FP64, k


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile,use_tmem,want_tau", [(32, 1, True), (13, 0, False)])
def test_eval_sparse_mass_format(emul, d, tile, use_tmem, want_tau):
    """Support-only M (rbd_eval_soa_sparse): the same sweep driven by the compact column programs writes exactly the
    structurally non-zero entries of the oracle's packed upper triangle, in the order of rbd_model_mass_pattern, and
    every entry it leaves out is an exact zero of the oracle's M."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n = 50
    q, v, a = mdl.random_state(d, n, seed=11)
    cd, keep = _cdesc(d)
    col_ptr_ref, row_ref, packed = helpers.mass_pattern_reference(d)
    col_ptr = np.zeros(d.nv + 1, dtype=np.int32); row_idx = np.zeros(max(1, packed.size), dtype=np.int32)
    nnz = emul.emul_mass_pattern(C.addressof(cd), P(col_ptr), P(row_idx))
    assert nnz == packed.size
    assert np.array_equal(col_ptr, col_ptr_ref) and np.array_equal(row_idx[:nnz], row_ref)
    nt = (n + tile - 1) // tile
    Mv = np.full((nt, max(1, nnz), tile), np.nan)
    tau = np.full((nt, d.nv, tile), np.nan) if want_tau else None
    oMi = np.full((nt, 12 * (d.njoints - 1), tile), np.nan)
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    rc = emul.emul_eval_ex(C.addressof(cd), n, tile, 0, use_tmem, P(qs), P(vs), P(as_), P(oMi), None, P(Mv), P(tau), 1, -2)
    assert rc == 0, emul.emul_last_error()
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    got = from_tiled(Mv, n)[:, :nnz]
    assert rel_err(got, ref["M"][:, packed]) <= TOL_FP64
    rest = np.ones(d.nv * (d.nv + 1) // 2, dtype=bool); rest[packed] = False
    assert np.all(ref["M"][:, rest] == 0.0)               # what the format drops is structurally zero
    assert rel_err(from_tiled(oMi, n), ref["oMi"]) <= TOL_FP64
    if want_tau:
        assert rel_err(from_tiled(tau, n), ref["tau"]) <= TOL_FP64


def _emul_eval_variant(emul, d, q, v, a, tile, variant, want=("oMi", "J", "M", "tau"), sparse=0):
    n = q.shape[0]
    cd, keep = _cdesc(d)
    nt = (n + tile - 1) // tile
    col_ptr = np.zeros(d.nv + 1, dtype=np.int32)
    nnz = emul.emul_mass_pattern(C.addressof(cd), P(col_ptr), None)
    K = {"oMi": 12 * (d.njoints - 1), "J": 6 * d.nv, "M": nnz if sparse else d.nv * (d.nv + 1) // 2, "tau": d.nv}
    outs = {k: (np.full((nt, max(1, K[k]), tile), np.nan) if k in want else None) for k in K}
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    rc = emul.emul_eval_ex(C.addressof(cd), n, tile, 0, 1, P(qs), P(vs), P(as_), P(outs["oMi"]), P(outs["J"]),
                           P(outs["M"]), P(outs["tau"]), sparse, variant)
    assert rc == 0, emul.emul_last_error()
    return {k: (from_tiled(o, n)[:, :K[k]] if o is not None else None) for k, o in outs.items()}


VARIANT_NAMES = {0: "generic", 1: "simple", 2: "chain", 3: "ffroot", 4: "ffroot8"}


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 9])
def test_eval_specialised_instantiations(emul, d, tile):
    """rbd::simple / rbd::chain / rbd::ffroot (the sweep compiled with joint-type knowledge; rbd_sweeps.h) compute
    bit-for-bit what the generic sweep computes, for every model that qualifies; the instantiation the planner picks
    matches the oracle; BASELINE robots get the instantiation they were built for."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n = 70
    q, v, a = mdl.random_state(d, n, seed=21)
    cd, keep = _cdesc(d)
    w = C.c_int(0)
    picked = emul.emul_eval_variant(C.addressof(cd), 1, 1, C.byref(w))
    assert picked >= 0
    expect = {"panda7": (2, 12), "solo12": (3, 12), "talos_reduced": (4, 8)}.get(d.name)
    if expect:
        assert (picked, w.value) == expect, (VARIANT_NAMES[picked], w.value)
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    base = _emul_eval_variant(emul, d, q, v, a, tile, -1)
    got = _emul_eval_variant(emul, d, q, v, a, tile, -2)
    for k in ref:
        assert rel_err(got[k], ref[k]) <= TOL_FP64, k
        assert np.array_equal(got[k], base[k]), k
    ran = 0
    for variant in (1, 2, 3, 4):
        cdv, _k = _cdesc(d)
        nt = (n + tile - 1) // tile
        probe = np.full((nt, max(1, d.nv), tile), np.nan)
        qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
        rc = emul.emul_eval_ex(C.addressof(cdv), n, tile, 0, 1, P(qs), P(vs), P(as_), None, None, None, P(probe), 0, variant)
        if rc == -102:
            continue                                        # the model does not qualify for this instantiation
        assert rc == 0, emul.emul_last_error()
        ran += 1
        for want, sparse in ((("oMi", "J", "M", "tau"), 0), (("M",), 1), (("tau",), 0), (("oMi", "M"), 0), (("J",), 0)):
            g = _emul_eval_variant(emul, d, q, v, a, tile, variant, want, sparse)
            b = _emul_eval_variant(emul, d, q, v, a, tile, -1, want, sparse)
            for k in want:
                assert np.array_equal(g[k], b[k]), (VARIANT_NAMES[variant], want, k)
    if d.name in ("panda7", "solo12", "talos_reduced"):
        assert ran >= 2


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("use_tmem", [0, 1])
def test_eval_f32_bodies_within_1e4(emul, d, use_tmem):
    """Optional FP32 mode (BASELINE north_star: tolerance 1e-4): the same sweep compiled with real = float."""
    import oracle
    from helpers import TOL_FP32
    from sofa.rigidbodydynamics_b200 import model as mdl
    n, tile = 70, 32
    q, v, a = mdl.random_state(d, n, seed=3)
    cd, keep = _cdesc(d)
    nt = (n + tile - 1) // tile
    K = {"oMi": 12 * (d.njoints - 1), "J": 6 * d.nv, "M": d.nv * (d.nv + 1) // 2, "tau": d.nv}
    outs = {k: np.full((nt, K[k], tile), np.nan, dtype=np.float32) for k in K}
    qs, vs, as_ = (to_tiled(x.astype(np.float32), tile) for x in (q, v, a))
    rc = emul.emul_eval_f32(C.addressof(cd), n, tile, 0, use_tmem, P(qs), P(vs), P(as_), P(outs["oMi"]), P(outs["J"]),
                            P(outs["M"]), P(outs["tau"]))
    assert rc == 0, emul.emul_last_error()
    ref, _ = oracle.eval_batch(d, q.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64),
                               a.astype(np.float32).astype(np.float64), 1)
    for k in ref:
        assert rel_err(from_tiled(outs[k], n).astype(np.float64), ref[k]) <= TOL_FP32, k


def _split(d, q, v):
    if d.has_freeflyer_root:
        return (np.ascontiguousarray(q[:, 7:]), np.ascontiguousarray(q[:, :7]),
                np.ascontiguousarray(v[:, 6:]), np.ascontiguousarray(v[:, :6]))
    return q, None, v, None


@pytest.mark.parametrize("d", MODELS_MAP, ids=IDS_MAP)
def test_mapping_bodies_match_oracle(emul, d):
    """apply -> applyJ -> applyJT -> applyJT(rows): the component's call order (KCM.inl:338-639)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n, tile, block, ctile = 45, 16, 64, 64
    q, v, _ = mdl.random_state(d, n, seed=8)
    qj, root, dq, rootv = _split(d, q, v)
    cd, keep = _cdesc(d)
    T = lambda x: None if x is None else to_tiled(x, tile)
    nt = (n + tile - 1) // tile
    qcache = np.full(((n + ctile - 1) // ctile, d.nq, ctile), np.nan)
    pos = np.full((nt, 7 * d.n_out, tile), np.nan)
    tq, tr = T(qj), T(root)
    assert emul.emul_apply(C.addressof(cd), n, tile, block, P(tq), P(tr), P(qcache), ctile, P(pos)) == 0
    assert rel_err(from_tiled(qcache, n), q) == 0.0
    vel = np.full((nt, 6 * d.n_out, tile), np.nan)
    tdq, trv = T(dq), T(rootv)
    assert emul.emul_applyJ(C.addressof(cd), n, tile, block, P(qcache), ctile, P(tdq), P(trv), P(vel)) == 0
    rng = np.random.default_rng(5)
    w = rng.uniform(-1, 1, (n, 6 * d.n_out))
    tau0 = rng.uniform(-1, 1, (n, d.nv_joints)); root0 = rng.uniform(-1, 1, (n, 6))
    tw, ttau = T(w), T(tau0)
    trt = T(root0) if root is not None else None
    assert emul.emul_applyJT(C.addressof(cd), n, tile, block, P(qcache), ctile, P(tw), P(ttau), P(trt)) == 0
    # the same product through the storage-plan sweep (shared-only and shared + tensor memory): identical accumulation
    for use_tmem in (0, 1):
        ttau2 = T(tau0); trt2 = T(root0) if root is not None else None
        assert emul.emul_applyJT_plan(C.addressof(cd), n, tile, use_tmem, P(qcache), ctile, P(tw), P(ttau2), P(trt2)) == 0
        assert rel_err(from_tiled(ttau2, n), from_tiled(ttau, n)) <= 1e-13
        if root is not None:
            assert rel_err(from_tiled(trt2, n), from_tiled(trt, n)) <= 1e-13
    # ... and through the ring variant (plan_jt_ring; the device ring is replaced by a direct reader)
    # (both sweeps: depth stack, and forward projection where the planner allows it)
    for use_tmem, allow_fwd in ((0, 1), (1, 0), (1, 1), (1, 2), (1, 3)):   # 3: the unrolled sweep of 6- / 7-joint chains
        ttau3 = T(tau0); trt3 = T(root0) if root is not None else None
        ns = C.c_int(-1); variant = C.c_int(-1)
        rc = emul.emul_applyJT_ring(C.addressof(cd), n, tile, use_tmem, P(qcache), ctile, P(tw), P(ttau3), P(trt3), C.byref(ns),
                                    allow_fwd, C.byref(variant))
        assert rc in (0, 1) and (rc == 0) == (ns.value > 0)
        assert not (variant.value == 1 and allow_fwd in (0, 3))
        assert not (allow_fwd == 3 and d.name == "panda7" and rc != 0)
        if rc == 0:
            assert rel_err(from_tiled(ttau3, n), from_tiled(ttau, n)) <= 1e-13
            if root is not None:
                assert rel_err(from_tiled(trt3, n), from_tiled(trt, n)) <= 1e-13
    pos = from_tiled(pos, n).reshape(n, d.n_out, 7); vel = from_tiled(vel, n).reshape(n, d.n_out, 6)
    tau = from_tiled(ttau, n); rt = from_tiled(trt, n) if trt is not None else None
    orc = oracle.Oracle(d)
    # constraint rows: 2 rows per instance
    row_ptr = [0]; row_inst = []; cols = []; vals = []
    for i in range(n):
        for _ in range(2):
            nnz = int(rng.integers(0, 4))
            for _ in range(nnz):
                cols.append(int(rng.integers(0, d.n_out))); vals.append(rng.uniform(-1, 1, 6))
            row_ptr.append(len(cols)); row_inst.append(i)
    row_ptr = np.array(row_ptr, dtype=np.int32); row_inst = np.array(row_inst, dtype=np.int32)
    cols = np.array(cols, dtype=np.int32); vals = np.array(vals, dtype=np.float64).reshape(-1, 6)
    nrows = len(row_inst)
    rows = np.full((nrows, d.nv), np.nan); keepf = np.zeros(nrows, dtype=np.int32)
    assert emul.emul_applyJT_rows(C.addressof(cd), nrows, block, P(qcache), ctile, P(row_ptr), P(row_inst), P(cols),
                                  P(vals), P(rows), P(keepf)) == 0
    # the same rows from cached Jacobians (the library's default row path)
    rows_c = np.full((nrows, d.nv), np.nan); keep_c = np.zeros(nrows, dtype=np.int32)
    assert emul.emul_applyJT_rows_cached(C.addressof(cd), n, P(qcache), ctile, nrows, P(row_ptr), P(row_inst), P(cols),
                                         P(vals), P(rows_c), P(keep_c)) == 0, emul.emul_last_error()
    assert rel_err(rows_c, rows) <= 1e-13 and np.array_equal(keep_c, keepf)
    # dense LOCAL_WORLD_ALIGNED frame Jacobians from the same cache: the oracle's getFrameJacobian, and J_k dq == applyJ
    pi = np.array([0, 0, n - 1, n // 2], dtype=np.int32)
    po = np.array([0, d.n_out - 1, d.n_out // 2, min(d.njoints, d.n_out - 1)], dtype=np.int32)
    Jd = np.full((len(pi), d.nv, 6), np.nan)
    assert emul.emul_frame_jacobians(C.addressof(cd), n, P(qcache), ctile, len(pi), P(pi), P(po), P(Jd)) == 0
    for t in range(len(pi)):
        i, k = int(pi[t]), int(po[t])
        orc.apply(qj[i], None if root is None else root[i])
        Jref = orc.frame_jacobian_lwa(int(d.out_frames[k]))
        assert rel_err(Jd[t].T, Jref) <= TOL_FP64
        full_dq = dq[i] if rootv is None else np.concatenate([rootv[i], dq[i]])
        assert rel_err(Jd[t].T @ full_dq, vel[i, k]) <= TOL_FP64
    # the ASSEMBLED mapping Jacobian (rbd_getJ_*): block rows of a permuted selection of outputs, non-zero columns only
    sel = rng.permutation(d.n_out)[:max(1, d.n_out // 2)].astype(np.int32)
    blk = np.zeros(len(sel) + 1, dtype=np.int32)
    assert emul.emul_getJ(C.addressof(cd), n, P(qcache), ctile, len(sel), P(sel), P(blk), None, None) == 0
    gcols = np.zeros(max(1, int(blk[-1])), dtype=np.int32)
    gval = np.full((n, max(1, int(blk[-1])), 6), np.nan)
    assert emul.emul_getJ(C.addressof(cd), n, P(qcache), ctile, len(sel), P(sel), P(blk), P(gcols), P(gval)) == 0
    for i in (0, n - 1):
        orc.apply(qj[i], None if root is None else root[i])
        for s_, k in enumerate(sel):
            Jref = orc.frame_jacobian_lwa(int(d.out_frames[int(k)]))            # dense 6 x nv
            cs = gcols[blk[s_]:blk[s_ + 1]]
            assert np.all(np.diff(cs) > 0)                                       # ascending dof columns
            assert rel_err(gval[i, blk[s_]:blk[s_ + 1]].T, Jref[:, cs]) <= TOL_FP64
            rest = np.ones(d.nv, dtype=bool); rest[cs] = False
            assert np.all(Jref[:, rest] == 0.0)                                  # what the pattern leaves out is zero
    for i in range(n):
        rp = orc.apply(qj[i], None if root is None else root[i])
        assert rel_err(pos[i][:, :3], rp[:, :3]) <= TOL_FP64
        assert helpers.quat_err(pos[i][:, 3:], rp[:, 3:]) <= 1e-9
        assert rel_err(vel[i], orc.applyJ(dq[i], None if rootv is None else rootv[i])) <= TOL_FP64
        rtau, rroot = orc.applyJT(w[i], tau0[i], root0[i] if root is not None else None)
        assert rel_err(tau[i], rtau) <= TOL_FP64
        if root is not None:
            assert rel_err(rt[i], rroot) <= TOL_FP64
        for r in (2 * i, 2 * i + 1):
            b, e = row_ptr[r], row_ptr[r + 1]
            rrow, rkeep = orc.applyJT_row(cols[b:e], vals[b:e])
            assert rel_err(rows[r], rrow) <= TOL_FP64
            assert bool(keepf[r]) == rkeep


def test_quaternion_branches(emul):
    """rot_to_quat follows Eigen's branch order (Conversions.h:60): same quaternion SIGN as the oracle (not just the
    same rotation) on rotations that hit all four branches.  Not bitwise: the kernels' sincos differs from libm by
    an ulp."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.random_tree(11, 1, all_types=False)   # a single revolute joint: outputs rotate with q
    rng = np.random.default_rng(2)
    n = 400
    q = rng.uniform(-np.pi, np.pi, (n, d.nq))
    q[:8, 0] = [0.0, np.pi, -np.pi, np.pi / 2, -np.pi / 2, 3.0, -3.0, 1e-9]
    cd, keep = _cdesc(d)
    tq = to_tiled(q, 32)
    pos = np.full((tq.shape[0], 7 * d.n_out, 32), np.nan)
    assert emul.emul_apply(C.addressof(cd), n, 32, 32, P(tq), None, None, 32, P(pos)) == 0
    pos = from_tiled(pos, n).reshape(n, d.n_out, 7)
    orc = oracle.Oracle(d)
    bad = 0
    for i in range(n):
        ref = orc.apply(q[i])
        assert helpers.quat_err(pos[i][:, 3:], ref[:, 3:]) <= 1e-12
        if np.max(np.abs(pos[i] - ref)) > 1e-12:
            bad += 1          # sign flipped: only legitimate exactly on a branch boundary (e.g. q = +-pi)
    assert bad <= 3


def test_stack_budget(emul):
    """Storage plan of the fused eval (DESIGN.md 'state budget'): resident warps per SM with and without tensor memory."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    info = (C.c_int * 8)()
    expect = {"panda7": (8, 6), "solo12": (8, 7), "talos_reduced": (6, 3)}
    for d in (mdl.panda7(), mdl.solo12(), mdl.talos_reduced()):
        cd, keep = _cdesc(d)
        assert emul.emul_model_info(C.addressof(cd), info) == 0
        smem_d, tmem_d, warps_tm, warps_sm = info[2], info[3], info[4], info[5]
        assert warps_tm >= expect[d.name][0] and warps_sm >= expect[d.name][1], (d.name, list(info))
        assert warps_tm >= warps_sm


def test_sincos_accuracy(emul):
    """The kernels' own sincos (Cody-Waite + fdlibm kernels) against 40-digit mpmath over the fast-path range and
    through the library fallback beyond it."""
    import mpmath
    mpmath.mp.dps = 40
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 3000), rng.uniform(-1e5, 1e5, 3000), rng.uniform(-1e9, 1e9, 200),
                        np.array([0.0, -0.0, np.pi / 2, np.pi, -np.pi, 1e-300, 0.7853981633974483, 99999.999, 1e5, 3e5]),
                        np.arange(-40, 41) * (np.pi / 4)])
    x = np.ascontiguousarray(x)
    s = np.zeros_like(x); c = np.zeros_like(x)
    emul.emul_sincos(x.shape[0], P(x), P(s), P(c))
    rs = np.array([float(mpmath.sin(mpmath.mpf(float(t)))) for t in x])
    rc = np.array([float(mpmath.cos(mpmath.mpf(float(t)))) for t in x])
    assert np.max(np.abs(s - rs)) <= 4.5e-16, np.max(np.abs(s - rs))
    assert np.max(np.abs(c - rc)) <= 4.5e-16, np.max(np.abs(c - rc))
    assert np.max(np.abs(s * s + c * c - 1.0)) <= 1e-15


def _unpack_upper(Mp, nv):
    """packed upper triangle by columns -> full symmetric [n, nv, nv]"""
    n = Mp.shape[0]
    M = np.zeros((n, nv, nv))
    for c in range(nv):
        for r in range(c + 1):
            M[:, r, c] = M[:, c, r] = Mp[:, c * (c + 1) // 2 + r]
    return M


@pytest.mark.parametrize("d", MODELS, ids=IDS)
def test_mass_forcefield_bodies(emul, d):
    """rbd_addMDx_soa / rbd_addForce_soa bodies (RNEA sweep with absent v or a, gravity switch, accumulation):
    res += f M(q) dx against the oracle's CRBA matrix; f -= rnea(q, v, 0) and the gravity-only form against its RNEA."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n, tile = 40, 16
    q, v, dx = mdl.random_state(d, n, seed=8)
    cd, keep = _cdesc(d)
    rng = np.random.default_rng(9)
    r0 = rng.uniform(-1, 1, (n, d.nv))
    zero = np.zeros_like(v)
    ref, _ = oracle.eval_batch(d, q, zero, zero, 1)
    M = _unpack_upper(ref["M"], d.nv)
    T = lambda x: to_tiled(x, tile)
    qs = T(q)
    res = T(r0)
    tdx, tv = T(dx), T(v)   # (P() hands out a raw address: the arrays must outlive the calls)
    assert emul.emul_mass_force(C.addressof(cd), n, tile, P(qs), None, P(tdx), P(res), 0.75, 0) == 0
    assert rel_err(from_tiled(res, n), r0 + 0.75 * np.einsum("nij,nj->ni", M, dx)) <= TOL_FP64
    f = T(r0)
    assert emul.emul_mass_force(C.addressof(cd), n, tile, P(qs), P(tv), None, P(f), -1.0, 1) == 0
    bias, _ = oracle.eval_batch(d, q, v, zero, 1)
    assert rel_err(from_tiled(f, n), r0 - bias["tau"]) <= TOL_FP64
    g = T(r0)
    assert emul.emul_mass_force(C.addressof(cd), n, tile, P(qs), None, None, P(g), -1.0, 1) == 0
    assert rel_err(from_tiled(g, n), r0 - ref["tau"]) <= TOL_FP64


def test_applyJT_ring_planner_choices(emul):
    """plan_jt_ring: small robots run the depth-stack sweep in 12-warp blocks, Talos needs the forward projection to get
    12 warps, deep chains fall back to 8 warps or to the register path; every plan respects the SM's budgets (227 KB of
    shared memory per block, 512 tensor-memory columns, 12 levels for the forward projection)."""
    from sofa.rigidbodydynamics_b200 import model as mdl

    def plan(d, use_tmem=1, mode=1):
        cd, keep = _cdesc(d)
        out = (C.c_int * 6)()
        assert emul.emul_jt_ring_plan(C.addressof(cd), use_tmem, mode, out) == 0
        return dict(zip(("stages", "fwd", "warps", "tmem", "smem", "bytes"), list(out)))

    solo, panda, talos = plan(mdl.solo12()), plan(mdl.panda7()), plan(mdl.talos_reduced())
    assert (solo["fwd"], solo["warps"]) == (0, 12) and solo["stages"] >= 2
    assert (panda["fwd"], panda["warps"]) == (0, 12)
    assert (talos["fwd"], talos["warps"]) == (1, 12) and talos["tmem"] <= 85 and talos["stages"] >= 2
    t8 = plan(mdl.talos_reduced(), mode=0)           # forward projection not allowed: 8 warps, depth stack
    assert (t8["fwd"], t8["warps"]) == (0, 8) and t8["tmem"] <= 128
    assert plan(mdl.talos_reduced(), use_tmem=0)["stages"] == 0    # no tensor memory: register path
    chain = plan(mdl.random_tree(7, 40, branch_prob=0.0))           # 40-joint serial chain: deeper than any plan
    assert chain["stages"] == 0 or (chain["fwd"] == 0 and chain["tmem"] <= 128)
    for p in (solo, panda, talos, t8):
        assert p["bytes"] <= 227 * 1024 and 2 * p["tmem"] * ((p["warps"] + 3) // 4) <= 512


@pytest.mark.parametrize("d", MODELS, ids=IDS)
def test_applyDJT_bodies_match_finite_differences(emul, d):
    """applyDJT (geometric stiffness, optional extension — the reference's applyDJT is a no-op): the analytic one-sweep
    derivative equals the central finite difference of applyJT along dq, (J(q+e dq)^T w - J(q-e dq)^T w) / 2e."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    if any(int(d.jtype[i]) == mdl.J_FF and int(d.parents[i]) != 0 for i in range(1, d.njoints)):
        pytest.skip("free-flyer below the root: not supported by applyDJT")
    n, tile, block, ctile = 12, 4, 32, 4
    q, v, _ = mdl.random_state(d, n, seed=13)
    cd, keep = _cdesc(d)
    rng = np.random.default_rng(14)
    w = rng.uniform(-1, 1, (n, 6 * d.n_out))
    T = lambda x: None if x is None else to_tiled(x, tile)
    ff = bool(d.has_freeflyer_root)
    tw = T(w)

    def jt(qx):
        qj, root, _, _ = _split(d, qx, v)
        qcache = np.full(((n + ctile - 1) // ctile, d.nq, ctile), np.nan)
        pos = np.full(((n + tile - 1) // tile, 7 * d.n_out, tile), np.nan)
        tqj, troot = T(qj), T(root)   # (P() hands out a raw address: the arrays must outlive the calls)
        assert emul.emul_apply(C.addressof(cd), n, tile, block, P(tqj), P(troot), P(qcache), ctile, P(pos)) == 0
        tau = T(np.zeros((n, d.nv_joints))); rt = T(np.zeros((n, 6))) if ff else None
        assert emul.emul_applyJT(C.addressof(cd), n, tile, block, P(qcache), ctile, P(tw), P(tau), P(rt)) == 0
        tj = from_tiled(tau, n)
        return (np.concatenate([from_tiled(rt, n), tj], axis=1) if ff else tj), qcache

    eps = 1e-6
    fd = (jt(helpers.integrate(d, q, v, eps))[0] - jt(helpers.integrate(d, q, v, -eps))[0]) / (2 * eps)
    _, qcache = jt(q)
    _, _, dq, rootv = _split(d, q, v)
    out = T(np.ones((n, d.nv_joints))); ort = T(np.ones((n, 6))) if ff else None      # accumulates onto ones
    tdq, trv = T(dq), T(rootv)
    assert emul.emul_applyDJT(C.addressof(cd), n, tile, P(qcache), ctile, P(tdq), P(trv), P(tw), 0.5, P(out), P(ort)) == 0
    got = from_tiled(out, n)
    if ff:
        got = np.concatenate([from_tiled(ort, n), got], axis=1)
    assert rel_err(got, 1.0 + 0.5 * fd) <= 1e-6
    # the ring kernel's forward-projection sweep gives the same numbers (where the planner has a plan for the model)
    out2 = T(np.ones((n, d.nv_joints))); ort2 = T(np.ones((n, 6))) if ff else None
    rc = emul.emul_applyDJT_ring(C.addressof(cd), n, tile, P(qcache), ctile, P(tdq), P(trv), P(tw), 0.5, P(out2), P(ort2))
    assert rc in (0, 1)
    if rc == 0:
        got2 = from_tiled(out2, n)
        if ff:
            got2 = np.concatenate([from_tiled(ort2, n), got2], axis=1)
        assert rel_err(got2, got) <= 1e-12


@pytest.mark.parametrize("name", ["solo12", "talos_reduced", "random_2", "schunk_svh_hand_right_ff", "panda7"])
def test_world_jacobian_by_finite_differences(emul, name):
    """Free-flyer velocity frame, [lin; ang] ordering and column ownership of the world joint Jacobian pinned by central
    differences of the kernel's own joint placements along pinocchio's integrate (helpers.fd_world_jacobian_error)."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = [m for m in MODELS if m.name == name][0]
    q, v, a = mdl.random_state(d, 3, seed=31)

    def eval_fn(qb):
        zeros = np.zeros((qb.shape[0], d.nv))
        got = _emul_eval(emul, d, qb, zeros, zeros, 32, 0, ("oMi", "J"))
        return got["oMi"], got["J"]

    assert helpers.fd_world_jacobian_error(eval_fn, d, q) <= 1e-7


# which forward-dynamics body: the first version (global-memory record), the fused sweep on the planner's plan, the fused
# sweep without tensor memory, and on blocks of 3 warps (other split of the levels between the two spaces)
ABA_BODIES = ["v1", "fused", "fused_smem_only", "fused_w3"]


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 11])
@pytest.mark.parametrize("body", ABA_BODIES)
def test_aba_body_matches_oracle_and_golden(emul, d, tile, body):
    """aba_instance (world-frame articulated-body sweep, scratch record per thread) and aba2_instance (passes 1 + 2 fused
    on depth-level records in shared / tensor memory) against the oracle's ABA (pinocchio's LOCAL-convention algorithm),
    the mpmath forward-dynamics golden vectors, and the round trip through the oracle's RNEA.  Tolerance 1e-9: the
    conditioning of M(q) enters (the golden solve is exact to 50 digits)."""
    import json
    import os
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    cd, keep = _cdesc(d)
    with open(os.path.join(helpers.GOLDEN, "golden_aba.json")) as f:
        gold = json.load(f)[d.name]["states"]
    n = 40
    q, v, a = mdl.random_state(d, n, seed=51)
    rng = np.random.default_rng(52)
    tau = rng.uniform(-2, 2, (n, d.nv))
    for s, st in enumerate(gold):                          # the golden states ride along as the first instances
        q[s], v[s], tau[s] = np.array(st["q"]), np.array(st["v"]), np.array(st["tau"])
    orc = oracle.Oracle(d)
    tau[n - 1] = orc.rnea(q[n - 1], v[n - 1], a[n - 1])    # round trip: aba(q, v, rnea(q, v, a)) == a
    nt = (n + tile - 1) // tile
    ddq = np.full((nt, max(1, d.nv), tile), np.nan)
    qs, vs, ts = to_tiled(q, tile), to_tiled(v, tile), to_tiled(tau, tile)
    if body == "v1":
        assert emul.emul_aba(C.addressof(cd), n, tile, P(qs), P(vs), P(ts), P(ddq)) == 0, emul.emul_last_error()
    else:
        plan = (C.c_int * 4)()
        rc = emul.emul_aba2(C.addressof(cd), n, tile, 0 if body == "fused_smem_only" else 1, 3 if body == "fused_w3" else 0,
                            P(qs), P(vs), P(ts), P(ddq), plan)
        assert rc == 0, emul.emul_last_error()
        if body == "fused" and d.name == "talos_reduced":
            assert list(plan)[:3] == [8, 105, 105]          # 8 warps per SM: 45 + 45 + 15 shared, 7 x 15 tensor memory
    got = from_tiled(ddq, n)[:, :d.nv]
    ref = np.stack([orc.aba(q[i], v[i], tau[i]) for i in range(n)])
    assert rel_err(got, ref) <= 1e-9
    for s, st in enumerate(gold):
        assert rel_err(got[s], np.array(st["qdd"])) <= 1e-9
    assert rel_err(got[n - 1], a[n - 1]) <= 1e-9


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 11])
@pytest.mark.parametrize("body", ["planned", "smem_only", "w3"])
def test_minv_body_matches_oracle(emul, d, tile, body):
    """minv_instance (M^-1 as forward dynamics of unit torques on the shared articulated-body quantities: backward walks
    along the path of each column, one forward sweep per block of columns) against the inverse of the oracle's CRBA mass
    matrix, the oracle's ABA on unit torques, and M^-1 M == I.  Packed upper triangle, like M."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    if d.nv == 0:
        pytest.skip("no dofs")
    cd, keep = _cdesc(d)
    n = 37
    q, _v, _a = mdl.random_state(d, n, seed=77)
    nt = (n + tile - 1) // tile
    mf = d.nv * (d.nv + 1) // 2
    Mv = np.full((nt, mf, tile), np.nan)
    qs = to_tiled(q, tile)
    plan = (C.c_int * 4)()
    rc = emul.emul_minv(C.addressof(cd), n, tile, 0 if body == "smem_only" else 1, 3 if body == "w3" else 0, P(qs), P(Mv), plan)
    assert rc == 0, emul.emul_last_error()
    got = from_tiled(Mv, n)[:, :mf]
    assert np.isfinite(got).all()
    orc = oracle.Oracle(d)
    iu = np.triu_indices(d.nv)
    for i in range(n):
        G = np.zeros((d.nv, d.nv))
        for c in range(d.nv):
            G[:c + 1, c] = got[i, c * (c + 1) // 2: c * (c + 1) // 2 + c + 1]
        G = np.triu(G) + np.triu(G, 1).T
        ref = orc.minv(q[i])
        scale = max(1.0, float(np.max(np.abs(ref))))
        assert np.max(np.abs(G - ref)) <= 1e-9 * scale, (i, np.max(np.abs(G - ref)) / scale)
        M = orc.crba(q[i]); M = np.triu(M) + np.triu(M, 1).T
        assert np.max(np.abs(G @ M - np.eye(d.nv))) <= 1e-8
        if i < 2:       # the articulated-body restatement on unit torques (v = 0: the gravity term cancels in the difference)
            z = np.zeros(d.nv)
            a0 = orc.aba(q[i], z, z)
            for c in (0, d.nv // 2, d.nv - 1):
                e = np.zeros(d.nv); e[c] = 1.0
                assert np.max(np.abs(orc.aba(q[i], z, e) - a0 - G[:, c])) <= 1e-9 * scale


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 11])
@pytest.mark.parametrize("body", ["planned", "smem_only", "w3"])
def test_drnea_body_matches_oracle(emul, d, tile, body):
    """drnea_instance (derivatives of inverse dynamics regrouped per (joint, ancestor) pair on the depth-first sweep,
    compact composite forms: 9 + 9 + 6 doubles instead of dense 6 x 6 matrices) against the oracle's restatement of
    pinocchio::computeRNEADerivatives (dense, two full passes), the 50-digit finite-difference golden vectors, and exact
    structural zeros.  M (dtau_da) in the packed format of the CRBA entry point."""
    import json
    import os
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    if d.nv == 0:
        pytest.skip("no dofs")
    cd, keep = _cdesc(d)
    with open(os.path.join(helpers.GOLDEN, "golden_drnea.json")) as f:
        gold = json.load(f)[d.name]
    n = 37
    q, v, a = mdl.random_state(d, n, seed=91)
    q[0], v[0], a[0] = np.array(gold["q"]), np.array(gold["v"]), np.array(gold["a"])
    nv = d.nv
    nt = (n + tile - 1) // tile
    DQ = np.full((nt, nv * nv, tile), np.nan); DV = np.full((nt, nv * nv, tile), np.nan)
    Mp = np.full((nt, nv * (nv + 1) // 2, tile), np.nan)
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    plan = (C.c_int * 4)()
    rc = emul.emul_drnea(C.addressof(cd), n, tile, 0 if body == "smem_only" else 1, 3 if body == "w3" else 0,
                         P(qs), P(vs), P(as_), P(DQ), P(DV), P(Mp) if body != "w3" else None, plan, 0)
    assert rc == 0, emul.emul_last_error()
    gq = from_tiled(DQ, n)[:, :nv * nv].reshape(n, nv, nv); gv = from_tiled(DV, n)[:, :nv * nv].reshape(n, nv, nv)
    assert np.isfinite(gq).all() and np.isfinite(gv).all()          # every entry written, structural zeros included
    orc = oracle.Oracle(d)
    related = helpers.support_relation(d)
    for i in range(n):
        rq, rv, rM = orc.rnea_derivatives(q[i], v[i], a[i])
        assert rel_err(gq[i], rq) <= 1e-10, (i, rel_err(gq[i], rq))
        assert rel_err(gv[i], rv) <= 1e-10, (i, rel_err(gv[i], rv))
        assert np.all(gq[i][~related] == 0.0) and np.all(gv[i][~related] == 0.0)
        if body != "w3":
            gM = from_tiled(Mp, n)[i, :nv * (nv + 1) // 2]
            assert np.isfinite(gM).all()
            assert rel_err(oracle.unpack_upper(gM, nv), np.triu(rM) + np.triu(rM, 1).T) <= 1e-10
    assert rel_err(gq[0], np.array(gold["dtau_dq"])) <= 1e-10
    assert rel_err(gv[0], np.array(gold["dtau_dv"])) <= 1e-10
    # support-only format: the pattern is the support relation, row by row with ascending columns, and the values are
    # the dense entries at the pattern's positions (bit for bit: same sweep, other addresses)
    row_ptr = np.zeros(nv + 1, dtype=np.int32); nnz = C.c_int(0)
    assert emul.emul_drnea_pattern(C.addressof(cd), P(row_ptr), None, C.byref(nnz)) == 0
    col_idx = np.zeros(max(1, nnz.value), dtype=np.int32)
    assert emul.emul_drnea_pattern(C.addressof(cd), P(row_ptr), P(col_idx), C.byref(nnz)) == 0
    assert nnz.value == int(related.sum()) == row_ptr[nv]
    for r in range(nv):
        cols = col_idx[row_ptr[r]:row_ptr[r + 1]]
        assert np.all(np.diff(cols) > 0) and np.array_equal(np.flatnonzero(related[r]), cols)
    SQ = np.full((nt, max(1, nnz.value), tile), np.nan); SV = np.full((nt, max(1, nnz.value), tile), np.nan)
    rc = emul.emul_drnea(C.addressof(cd), n, tile, 0 if body == "smem_only" else 1, 3 if body == "w3" else 0,
                         P(qs), P(vs), P(as_), P(SQ), P(SV), None, plan, 1)
    assert rc == 0, emul.emul_last_error()
    rows = np.repeat(np.arange(nv), np.diff(row_ptr))
    assert np.array_equal(from_tiled(SQ, n)[:, :nnz.value], gq[:, rows, col_idx[:nnz.value]])
    assert np.array_equal(from_tiled(SV, n)[:, :nnz.value], gv[:, rows, col_idx[:nnz.value]])
    if body == "planned" and d.name == "talos_reduced":
        assert nnz.value == 700          # of 1 444: fewer than half of the entries are structurally non-zero
        # 8 warps per SM: hot records 24 + 9 x 18 doubles on chip (+ input ring and staging: 18), cold ones 42 + 9 + 42 + 7 x 9 in scratch
        assert list(plan) == [8, 96, 108, 156]


def test_block_size_per_launch(emul):
    """pick_warps_per_block: a persistent kernel planned for W warps is launched with the block size that leaves the fewest
    warps idle in the last round (cost of a round ~ 2.5 + w): Panda's 2 048 tiles on 148 SMs are two full rounds of 7-warp
    blocks; large batches and the sizing call (no tiles) keep W; never fewer than 4 warps, never more than W."""
    assert emul.emul_pick_wpb(12, 2048, 148) == 7 and emul.emul_pick_wpb(8, 2048, 148) == 7
    assert emul.emul_pick_wpb(12, 32768, 148) == 12 and emul.emul_pick_wpb(8, 32768, 148) == 8
    assert emul.emul_pick_wpb(12, 8192, 148) == 12
    assert emul.emul_pick_wpb(8, 0, 148) == 8 and emul.emul_pick_wpb(12, 1, 148) == 4
    for W in (4, 8, 12):
        for nt in (1, 100, 1183, 1184, 1185, 5000, 100000):
            w = emul.emul_pick_wpb(W, nt, 148)
            assert 4 <= w <= W
            rounds = lambda x: -(-nt // (148 * x))
            assert rounds(w) * (2.5 + w) <= rounds(W) * (2.5 + W) + 1e-9


@pytest.mark.parametrize("nmoving,free_flyer", [(40, False), (45, True), (63, False)])
def test_drnea_deep_chains(emul, nmoving, free_flyer):
    """Long serial chains: the hot level records no longer fit 8 warps per SM — the planner takes smaller blocks (levels in
    shared AND tensor memory, more of each per thread) until they fit, and refuses (-100, RBD_ERR_UNSUPPORTED in the
    library) when even one warp per SM cannot hold them.  Results against the oracle."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.random_tree(100 + nmoving, nmoving, free_flyer=free_flyer, branch_prob=0.0, all_types=False)
    cd, keep = _cdesc(d)
    n, tile, nv = 5, 32, d.nv
    q, v, a = mdl.random_state(d, n, seed=3)
    DQ = np.full((1, nv * nv, tile), np.nan); DV = np.full((1, nv * nv, tile), np.nan)
    plan = (C.c_int * 4)()
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)      # (kept alive across the call)
    rc = emul.emul_drnea(C.addressof(cd), n, tile, 1, 0, P(qs), P(vs), P(as_), P(DQ), P(DV), None, plan, 0)
    if nmoving == 63:
        assert rc == -100          # 62 levels x 18 doubles: more than one warp's share of shared + tensor memory
        return
    assert rc == 0, emul.emul_last_error()
    assert 1 <= list(plan)[0] < 8
    orc = oracle.Oracle(d)
    gq = from_tiled(DQ, n)[:, :nv * nv].reshape(n, nv, nv); gv = from_tiled(DV, n)[:, :nv * nv].reshape(n, nv, nv)
    for i in range(n):
        rq, rv, _ = orc.rnea_derivatives(q[i], v[i], a[i])
        assert rel_err(gq[i], rq) <= 1e-9 and rel_err(gv[i], rv) <= 1e-9, (i, rel_err(gq[i], rq), rel_err(gv[i], rv))


@pytest.mark.parametrize("tile", [32, 13])
@pytest.mark.parametrize("nj", [6, 7, 5])
def test_arm_body_matches_oracle_and_generic_sweep(emul, tile, nj):
    """rbd::chain::arm_instance (serial chains with a compile-time joint count: both passes unrolled, subspace columns in
    registers, f | Y stack only) against the oracle and against the generic sweep (same formulas: equal to rounding).  A
    5-joint chain has no instantiation (-102: the library then runs the generic chain sweep)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.panda7()
    if nj != 7:     # a shorter arm: the first nj joints of the Panda chain, frames re-derived by the builder
        b = mdl._Builder(f"arm{nj}")
        b.add_frame("root_joint", 0, mdl._se3(), mdl.F_FIXED_JOINT)
        for i in range(1, nj + 1):
            jid = b.add_joint(i - 1, int(d.jtype[i]), d.placement[i], f"j{i}")
            b.add_frame(f"j{i}", jid, mdl._se3(), mdl.F_JOINT)
            b.inertia[jid] = d.inertia[i].copy()
            b.add_frame(f"b{i}", jid, mdl._se3(), mdl.F_BODY)
        d = b.finish()
    n = 45
    q, v, a = mdl.random_state(d, n, seed=5)
    cd, keep = _cdesc(d)
    nt = (n + tile - 1) // tile
    K = {"oMi": 12 * (d.njoints - 1), "J": 6 * d.nv, "M": d.nv * (d.nv + 1) // 2, "tau": d.nv}
    outs = {k: np.full((nt, K[k], tile), np.nan) for k in K}
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    rc = emul.emul_eval_arm(C.addressof(cd), n, tile, P(qs), P(vs), P(as_), P(outs["oMi"]), P(outs["J"]), P(outs["M"]), P(outs["tau"]))
    if nj == 5:
        assert rc == -102
        return
    assert rc == 0, emul.emul_last_error()
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    base = _emul_eval_variant(emul, d, q, v, a, tile, -1)
    for k in K:
        got = from_tiled(outs[k], n)[:, :K[k]]
        assert np.isfinite(got).all(), k
        assert rel_err(got, ref[k]) <= TOL_FP64, k
        assert rel_err(got, base[k]) <= 1e-13, k
    # oMi and J are optional
    o2 = {k: np.full((nt, K[k], tile), np.nan) for k in ("M", "tau")}
    assert emul.emul_eval_arm(C.addressof(cd), n, tile, P(qs), P(vs), P(as_), None, None, P(o2["M"]), P(o2["tau"])) == 0
    for k in o2:
        assert np.array_equal(o2[k], outs[k], equal_nan=True), k

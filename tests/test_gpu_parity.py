"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (ctypes -> librbd_b200.so), against the
CPU oracle on the same seeded inputs, against the committed mpmath golden vectors, and — at BASELINE sizes — through
size-independent identities.  Tolerance: 1e-10 relative (BASELINE.json north_star, FP64)."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import TOL_FP64, rel_err, to_tiled, from_tiled

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def api(rbd_lib):
    from sofa.rigidbodydynamics_b200 import api as a
    assert a.device_count() >= 1, "librbd_b200.so sees no CUDA device: the product path has no fallback"
    return a


MODELS = helpers.all_models()
IDS = [m.name for m in MODELS]
SUBSETS = helpers.subset_models()                      # permuted strict subsets of the frames as mapped outputs
MODELS_MAP = MODELS + SUBSETS
IDS_MAP = [m.name for m in MODELS_MAP]


def _dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _eval_gpu(torch, api, d, q, v, a, tile, want=("oMi", "J", "M", "tau"), block=0, tmem=True):
    n = q.shape[0]
    m = api.Model(d); ws = api.Workspace(m, max(n, 1))
    if block:
        ws.set_eval_block(block)
    ws.set_eval_tmem(tmem)
    I = m.info
    nt = (n + tile - 1) // tile
    K = {"oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": I.nv}
    outs = {k: (torch.full((nt, K[k], tile), float("nan"), dtype=torch.float64, device="cuda") if k in want else None)
            for k in K}
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    ws.eval_soa(n, tile, qs, vs, as_, outs["oMi"], outs["J"], outs["M"], outs["tau"])
    ws.synchronize()
    res = {k: (from_tiled(o.cpu().numpy(), n) if o is not None else None) for k, o in outs.items()}
    return res, ws.launch_count


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile,tmem", [(32, True), (32, False), (1000, True)])
def test_eval_matches_oracle(torch_mod, api, d, tile, tmem):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    n = 777
    q, v, a = mdl.random_state(d, n, seed=11)
    got, launches = _eval_gpu(torch_mod, api, d, q, v, a, tile, tmem=tmem)
    assert launches == 1
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for k in ("oMi", "J", "M", "tau"):
        assert rel_err(got[k], ref[k]) <= TOL_FP64, k


@pytest.mark.parametrize("want", [("M",), ("tau",), ("oMi", "J"), ("J", "tau"), ("oMi", "M")])
def test_eval_partial_outputs(torch_mod, api, want):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.talos_reduced()
    q, v, a = mdl.random_state(d, 300, seed=5)
    got, _ = _eval_gpu(torch_mod, api, d, q, v, a, 64, want=want)
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for k in ("oMi", "J", "M", "tau"):
        if k in want:
            assert rel_err(got[k], ref[k]) <= TOL_FP64, k
        else:
            assert got[k] is None


@pytest.mark.parametrize("block,tmem", [(32, False), (64, False), (96, False), (64, True), (128, True), (192, True)])
def test_eval_block_sizes(torch_mod, api, block, tmem):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.talos_reduced()
    q, v, a = mdl.random_state(d, 333, seed=6)
    got, _ = _eval_gpu(torch_mod, api, d, q, v, a, 333, block=block, tmem=tmem)
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for k in ("oMi", "J", "M", "tau"):
        assert rel_err(got[k], ref[k]) <= TOL_FP64, k


def test_block_too_large_is_an_error(torch_mod, api):
    """256 Talos instances do not fit one SM: the planner refuses instead of launching something slower."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.talos_reduced()
    q, v, a = mdl.random_state(d, 8, seed=6)
    with pytest.raises(api.RbdError) as e:
        _eval_gpu(torch_mod, api, d, q, v, a, 32, block=256, tmem=False)
    assert e.value.code == -5   # RBD_ERR_UNSUPPORTED


def test_golden_vectors(torch_mod, api):
    """CUDA path against the committed 50-digit mpmath vectors (tests/golden/make_golden.py)."""
    with open(os.path.join(helpers.GOLDEN, "golden_vectors.json")) as f:
        G = json.load(f)
    byname = {m.name: m for m in MODELS}
    from oracle import unpack_upper
    checked = 0
    for name, g in G.items():
        d = byname[name]
        q = np.array([s["q"] for s in g["states"]]); v = np.array([s["v"] for s in g["states"]])
        a = np.array([s["a"] for s in g["states"]])
        got, _ = _eval_gpu(torch_mod, api, d, q, v, a, 32)
        for i, s in enumerate(g["states"]):
            assert rel_err(got["oMi"][i], np.array(s["oMi"]).reshape(-1)) <= TOL_FP64
            assert rel_err(got["J"][i], np.array(s["J"])) <= TOL_FP64
            assert rel_err(unpack_upper(got["M"][i], d.nv), np.array(s["M"]).reshape(d.nv, d.nv)) <= TOL_FP64
            assert rel_err(got["tau"][i], np.array(s["tau"])) <= TOL_FP64
            checked += 1
    assert checked >= 20


# ---- mapping path -------------------------------------------------------------------------------------------
def _split(d, q, v):
    if d.has_freeflyer_root:
        return (np.ascontiguousarray(q[:, 7:]), np.ascontiguousarray(q[:, :7]),
                np.ascontiguousarray(v[:, 6:]), np.ascontiguousarray(v[:, :6]))
    return q, None, v, None


@pytest.mark.parametrize("d", MODELS_MAP, ids=IDS_MAP)
@pytest.mark.parametrize("n,tile,ring", [(200, 64, True), (200, 32, True), (200, 32, False), (4000, 32, True)])
def test_mapping_soa_matches_oracle(torch_mod, api, d, n, tile, ring):
    """apply -> applyJ -> applyJT on device tiled-SoA buffers (the component's call order, KCM.inl:338-513).
    tile 32 sends the full tiles of applyJT through the ring kernel (bulk-copy staged wrenches) and the partial last
    tile through the register path; `ring` False pins the register path for all of it."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    q, v, _ = mdl.random_state(d, n, seed=21)
    qj, root, dq, rootv = _split(d, q, v)
    rng = np.random.default_rng(3)
    w = rng.uniform(-1, 1, (n, d.n_out * 6))
    tau0 = rng.uniform(-1, 1, (n, d.nv_joints)); root0 = rng.uniform(-1, 1, (n, 6))
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_applyJT_ring(ring)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")
    vel = torch.zeros((nt, 6 * d.n_out, tile), dtype=torch.float64, device="cuda")
    tau = T(tau0); rt = T(root0) if root is not None else None
    # if qj has zero fields (1-dof FF-only models do not occur here) the pointer is still valid
    ws.apply_soa(n, tile, T(qj), T(root), pos)
    ws.applyJ_soa(n, tile, T(dq), T(rootv), vel)
    ws.applyJT_soa(n, tile, T(w), tau, rt)
    ws.synchronize()
    pos = from_tiled(pos.cpu().numpy(), n).reshape(n, d.n_out, 7)
    vel = from_tiled(vel.cpu().numpy(), n).reshape(n, d.n_out, 6)
    tau = from_tiled(tau.cpu().numpy(), n)
    rt = from_tiled(rt.cpu().numpy(), n) if rt is not None else None
    orc = oracle.Oracle(d)
    for i in list(range(0, n, 7 if n < 1000 else 97)) + [n - 1]:
        rp = orc.apply(qj[i], None if root is None else root[i])
        assert rel_err(pos[i][:, :3], rp[:, :3]) <= TOL_FP64
        # identical branch order as Eigen's Quaternion(Matrix3): equality, not just up to sign — except when the
        # trace sits within rounding of the branch point
        assert helpers.quat_err(pos[i][:, 3:], rp[:, 3:]) <= 1e-9
        rv = orc.applyJ(dq[i], None if rootv is None else rootv[i])
        assert rel_err(vel[i], rv) <= TOL_FP64
        rtau, rroot = orc.applyJT(w[i], tau0[i], root0[i] if root is not None else None)
        assert rel_err(tau[i], rtau) <= TOL_FP64
        if root is not None:
            assert rel_err(rt[i], rroot) <= TOL_FP64


def test_applyJ_before_apply_is_an_error(api):
    from sofa.rigidbodydynamics_b200 import model as mdl
    import torch
    d = mdl.panda7()
    m = api.Model(d); ws = api.Workspace(m, 8)
    x = torch.zeros((1, d.nv, 8), dtype=torch.float64, device="cuda")
    o = torch.zeros((1, 6 * d.n_out, 8), dtype=torch.float64, device="cuda")
    with pytest.raises(api.RbdError) as e:
        ws.applyJ_soa(8, 8, x, None, o)
    assert e.value.code == -4   # RBD_ERR_NO_APPLY (KCM.inl:365-372: applyJ relies on the previous apply)


@pytest.mark.parametrize("d", [m for m in MODELS if m.name in ("panda7", "solo12", "talos_reduced", "random_2",
                                                                "schunk_svh_hand_right_regularized_inertia")],
                         ids=lambda m: m.name)
def test_component_mirror_host_path(api, d):
    """KinematicChainMapping mirror over the *_host entry points (SOFA instance-major layout, several chunks)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200.mapping import KinematicChainMapping
    n = 20000 if d.nv < 20 else 17000      # > one 16384-instance chunk: exercises the two-slot pipeline
    q, v, _ = mdl.random_state(d, n, seed=31)
    qj, root, dq, rootv = _split(d, q, v)
    kcm = KinematicChainMapping(d, n_instances=n)
    pos = np.zeros((n, d.n_out, 7)); vel = np.zeros((n, d.n_out, 6))
    in2 = [] if root is None else [np.ascontiguousarray(root[:, None, :])]
    in2v = [] if rootv is None else [np.ascontiguousarray(rootv[:, None, :])]
    kcm.apply([pos], [qj], in2)
    kcm.applyJ([vel], [dq], in2v)
    rng = np.random.default_rng(9)
    w = rng.uniform(-1, 1, (n, d.n_out, 6))
    f1 = rng.uniform(-1, 1, (n, d.nv_joints)); f1_0 = f1.copy()
    f2 = rng.uniform(-1, 1, (n, 1, 6)); f2_0 = f2.copy()
    kcm.applyJT([f1], [] if root is None else [f2], [w])
    assert kcm.d_componentState == "Valid"
    orc = oracle.Oracle(d)
    for i in list(range(0, n, 1999)) + [16383, 16384, n - 1]:
        rp = orc.apply(qj[i], None if root is None else root[i])
        assert rel_err(pos[i][:, :3], rp[:, :3]) <= TOL_FP64
        assert helpers.quat_err(pos[i][:, 3:], rp[:, 3:]) <= 1e-9
        assert rel_err(vel[i], orc.applyJ(dq[i], None if rootv is None else rootv[i])) <= TOL_FP64
        rtau, rroot = orc.applyJT(w[i].reshape(-1), f1_0[i], f2_0[i, 0] if root is not None else None)
        assert rel_err(f1[i], rtau) <= TOL_FP64
        if root is not None:
            assert rel_err(f2[i, 0], rroot) <= TOL_FP64
    # constraint rows (applyJT MatrixDeriv, KCM.inl:515-639), a handful of instances
    rows = {}
    for r in range(40):
        inst = int(rng.integers(0, n))
        nnz = int(rng.integers(1, 4))
        rows[(inst, r)] = [(int(rng.integers(0, d.n_out)), rng.uniform(-1, 1, 6)) for _ in range(nnz)]
    rows[(5, 40)] = [(0, np.ones(6))]                       # universe output: empty support -> dropped row
    rows[(6, 41)] = [(d.n_out - 1, np.zeros(6))]            # zero wrench -> dropped row
    out, out_root = kcm.applyJT_matrix(rows)
    assert (5, 40) not in out and (6, 41) not in out
    for key, entries in rows.items():
        inst = key[0]
        orc.apply(qj[inst], None if root is None else root[inst])
        row, keep = orc.applyJT_row([c for c, _ in entries], np.array([x for _, x in entries]))
        assert keep == (key in out)
        if keep:
            full = np.concatenate([out_root[key], out[key]]) if root is not None else out[key]
            assert rel_err(full, row) <= TOL_FP64


def test_eval_host_matches_soa(torch_mod, api):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.solo12()
    n = 40000
    q, v, a = mdl.random_state(d, n, seed=41)
    m = api.Model(d); ws = api.Workspace(m, n)
    I = m.info
    oMi = np.zeros((n, I.omi_fields)); J = np.zeros((n, I.jac_fields)); M = np.zeros((n, I.mass_fields))
    tau = np.zeros((n, I.nv))
    ws.eval_host(n, q, v, a, oMi, J, M, tau)
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for k, g in (("oMi", oMi), ("J", J), ("M", M), ("tau", tau)):
        assert rel_err(g, ref[k]) <= TOL_FP64, k


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 1000])
def test_eval_sparse_mass_format(torch_mod, api, d, tile):
    """rbd_eval_soa_sparse: M in the support-only format equals the structurally non-zero entries of the oracle's packed
    upper triangle (order of rbd_model_mass_pattern); the other outputs are unchanged; the kernel that ran is the same
    eval instantiation (read back from the launch)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    n = 333
    q, v, a = mdl.random_state(d, n, seed=12)
    m = api.Model(d); ws = api.Workspace(m, n)
    I = m.info
    col_ptr_ref, row_ref, packed = helpers.mass_pattern_reference(d)
    col_ptr, row_idx = m.mass_pattern()
    assert I.mass_support_fields == packed.size
    assert np.array_equal(col_ptr, col_ptr_ref) and np.array_equal(row_idx, row_ref)
    assert np.array_equal(m.mass_packed_index(), packed)
    nt = (n + tile - 1) // tile
    new = lambda K: torch.full((nt, max(1, K), tile), float("nan"), dtype=torch.float64, device="cuda")
    oMi, J, Mv, tau = new(I.omi_fields), new(I.jac_fields), new(I.mass_support_fields), new(I.nv)
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    ws.eval_soa_sparse(n, tile, qs, vs, as_, oMi, J, Mv, tau)
    ws.synchronize()
    assert any(k in api.last_kernel_name() for k in ("eval_kernel", "eval_arm_kernel"))
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    assert rel_err(from_tiled(Mv.cpu().numpy(), n)[:, :packed.size], ref["M"][:, packed]) <= TOL_FP64
    for k, g in (("oMi", oMi), ("J", J), ("tau", tau)):
        assert rel_err(from_tiled(g.cpu().numpy(), n)[:, :ref[k].shape[1]], ref[k]) <= TOL_FP64, k
    # crba only, same format
    Mv2 = new(I.mass_support_fields)
    ws.crba_soa_sparse(n, tile, qs, Mv2)
    ws.synchronize()
    # (another instantiation of the sweep: the compiler may contract multiply-adds differently -> last-bit differences)
    assert rel_err(from_tiled(Mv2.cpu().numpy(), n), from_tiled(Mv.cpu().numpy(), n)) <= 1e-13


def test_eval_host_sparse_and_output_selection(api):
    """rbd_eval_host_sparse: all outputs, and the Mass / ForceField selection (Mval + tau only) — a NULL output is neither
    computed nor copied; the values do not depend on the selection."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.talos_reduced()
    n = 40000                                               # two pipelined chunks, the second one partial
    q, v, a = mdl.random_state(d, n, seed=43)
    m = api.Model(d); ws = api.Workspace(m, n)
    I = m.info
    packed = m.mass_packed_index()
    oMi = np.zeros((n, I.omi_fields)); J = np.zeros((n, I.jac_fields)); Mv = np.zeros((n, I.mass_support_fields))
    tau = np.zeros((n, I.nv))
    ws.eval_host_sparse(n, q, v, a, oMi, J, Mv, tau)
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    assert rel_err(Mv, ref["M"][:, packed]) <= TOL_FP64
    for k, g in (("oMi", oMi), ("J", J), ("tau", tau)):
        assert rel_err(g, ref[k]) <= TOL_FP64, k
    Mv2 = np.zeros_like(Mv); tau2 = np.zeros_like(tau)
    ws.eval_host_sparse(n, q, v, a, None, None, Mv2, tau2)
    assert rel_err(Mv2, Mv) <= 1e-13 and rel_err(tau2, tau) <= 1e-13   # (other instantiations of the sweep)
    Mv3 = np.zeros_like(Mv)
    ws.eval_host_sparse(n, q, None, None, None, None, Mv3, None)   # crba alone: v and a do not travel
    assert rel_err(Mv3, Mv) <= 1e-13


def test_rows_host_rejects_malformed_csr(api):
    """rbd_applyJT_rows_host validates the constraint matrix before anything reaches the device (a bad column index
    would read the per-output tables and the Jacobian cache out of bounds)."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.panda7()
    n = 4
    q, v, _ = mdl.random_state(d, n, seed=3)
    m = api.Model(d); ws = api.Workspace(m, n)
    pos = np.zeros((n, d.n_out, 7))
    ws.apply_host(n, np.ascontiguousarray(q), None, pos)
    rows = np.zeros((2, d.nv)); keep = np.zeros(2, dtype=np.int32)
    val = np.ones((2, 6))
    ok = dict(row_ptr=np.array([0, 1, 2], dtype=np.int32), row_instance=np.array([0, 3], dtype=np.int32),
              col=np.array([1, d.n_out - 1], dtype=np.int32))
    ws.applyJT_rows_host(2, ok["row_ptr"], ok["row_instance"], ok["col"], val, rows, keep)
    bad = [dict(ok, col=np.array([1, d.n_out], dtype=np.int32)), dict(ok, col=np.array([-1, 0], dtype=np.int32)),
           dict(ok, row_ptr=np.array([0, 2, 1], dtype=np.int32)), dict(ok, row_ptr=np.array([1, 1, 2], dtype=np.int32)),
           dict(ok, row_instance=np.array([0, n], dtype=np.int32))]
    for b in bad:
        with pytest.raises(api.RbdError) as e:
            ws.applyJT_rows_host(2, b["row_ptr"], b["row_instance"], b["col"], val, rows, keep)
        assert e.value.code == -1
    ws.applyJT_rows_host(2, ok["row_ptr"], ok["row_instance"], ok["col"], val, rows, keep)   # the workspace is still usable


# ---- BASELINE sizes: size-independent identities --------------------------------------------------------------
def test_full_size_identities(torch_mod, api):
    """Talos at N = 1,048,576 (BASELINE configs[3]): (1) rnea(q,v,a) - rnea(q,v,0) == M(q) a for every instance,
    (2) M symmetric-positive (diagonal > 0), (3) a strided sample equals the oracle, (4) the chunked result is
    independent of the tile size (checksum of outputs)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200.dynamics import BatchedDynamics
    torch = torch_mod
    d = mdl.talos_reduced()
    n = 1 << 20
    dyn = BatchedDynamics(d, n)
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    nq, nv = d.nq, d.nv
    q = (torch.rand((1, nq, n), generator=g, dtype=torch.float64, device="cuda") * 2 - 1) * np.pi
    q[0, 0:3] = q[0, 0:3] / np.pi
    quat = torch.randn((4, n), generator=g, dtype=torch.float64, device="cuda")
    q[0, 3:7] = quat / quat.norm(dim=0, keepdim=True)
    v = torch.rand((1, nv, n), generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    a = torch.rand((1, nv, n), generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    out = dyn.alloc_outputs()
    dyn.eval(q, v, a, out)
    tau0 = dyn.empty("tau")
    dyn.rnea(q, v, torch.zeros_like(a), tau0)
    torch.cuda.synchronize()
    # M a from the packed upper triangle
    Mp = out["M"][0]                                             # [741, n]
    Ma = torch.zeros((nv, n), dtype=torch.float64, device="cuda")
    for c in range(nv):
        base = c * (c + 1) // 2
        Ma[c] += Mp[base + c] * a[0, c]
        if c:
            col = Mp[base:base + c]                              # rows 0..c-1 of column c
            Ma[:c] += col * a[0, c]
            Ma[c] += (col * a[0, :c]).sum(dim=0)
    lhs = out["tau"][0] - tau0[0]
    scale = max(1.0, float(lhs.abs().max()))
    assert float((lhs - Ma).abs().max()) / scale <= 1e-9          # two rnea + crba roundings accumulate
    diag = torch.stack([Mp[c * (c + 1) // 2 + c] for c in range(nv)])
    assert bool((diag > 0).all())
    idx = torch.arange(0, n, 8191, device="cuda")
    qs = q[0][:, idx].T.contiguous().cpu().numpy(); vs = v[0][:, idx].T.contiguous().cpu().numpy()
    as_ = a[0][:, idx].T.contiguous().cpu().numpy()
    ref, _ = oracle.eval_batch(d, qs, vs, as_, nthreads=0)
    for k in ("oMi", "J", "M", "tau"):
        assert rel_err(out[k][0][:, idx].T.cpu().numpy(), ref[k]) <= TOL_FP64, k


def test_layout_adapters_roundtrip(torch_mod, api):
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    m = api.Model(mdl.panda7()); ws = api.Workspace(m, 8)
    for n, K, tile in ((1, 1, 1), (1000, 37, 32), (4097, 100, 4097), (33, 741, 64), (4097, 741, 32), (31, 200, 32),
                       (64, 128, 32), (33, 129, 32), (1, 1, 32), (100, 7, 32), (4096, 256, 32), (2049, 255, 32),
                       (95, 2, 32), (66, 3, 32), (32, 127, 32), (63, 130, 32), (1025, 228, 32), (7, 396, 32)):
        for shift in (0, 1):        # shift 1: base pointers 8-byte aligned only (the 16-byte kernel must not be picked)
            xb = torch.randn((n * K + shift,), dtype=torch.float64, device="cuda")
            x = xb[shift:].view(n, K)
            nt = (n + tile - 1) // tile
            sb = torch.zeros((nt * K * tile + shift,), dtype=torch.float64, device="cuda")
            s = sb[shift:].view(nt, K, tile)
            ws.aos_to_soa(n, K, tile, x, s)
            # exact image, padding lanes of the last tile included (they must stay untouched: zeros)
            assert np.array_equal(s.cpu().numpy(), to_tiled(x.cpu().numpy(), tile)), (n, K, tile, shift)
            yb = torch.full((n * K + shift + 8,), -7.0, dtype=torch.float64, device="cuda")
            y = yb[shift:shift + n * K].view(n, K)
            ws.soa_to_aos(n, K, tile, s, y)
            assert torch.equal(x, y), (n, K, tile, shift)
            assert bool((yb[shift + n * K:] == -7.0).all()) and bool((yb[:shift] == -7.0).all())


@pytest.mark.parametrize("name,n,tile", [("talos_reduced", 1000, 32), ("solo12", 77, 32), ("random_3", 333, 100),
                                         ("panda7", 31, 32)])
def test_no_out_of_bounds_writes(torch_mod, api, name, n, tile):
    """compute-sanitizer is closed on this pool, so: every device buffer sits between guard bands of a sentinel value
    inside one arena; after eval + apply + applyJ + applyJT the bands must be intact (the tail lanes of the last warp
    recompute the last instance and may only touch the padding INSIDE the last tile)."""
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = [m for m in MODELS if m.name == name][0]
    q, v, a = mdl.random_state(d, n, seed=17)
    m = api.Model(d); ws = api.Workspace(m, n)
    I = m.info
    nt = (n + tile - 1) // tile
    G = 4096                                         # guard doubles
    sizes = {"q": I.nq, "v": I.nv, "a": I.nv, "oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": I.nv,
             "qj": I.nq_joints, "root": 7, "dq": I.nv_joints, "rootv": 6, "pos": 7 * I.n_out, "vel": 6 * I.n_out,
             "w": 6 * I.n_out, "tj": I.nv_joints, "tr": 6,
             "DQ": I.nv * I.nv, "DV": I.nv * I.nv, "SQ": max(1, len(m.drnea_pattern()[1])), "SV": max(1, len(m.drnea_pattern()[1]))}
    total = G + sum(nt * K * tile + G for K in sizes.values())
    SENT = -7.25e300
    arena = torch.full((total,), SENT, dtype=torch.float64, device="cuda")
    views, off = {}, G
    for k, K in sizes.items():
        views[k] = arena[off:off + nt * K * tile].view(nt, K, tile)
        off += nt * K * tile + G
    views["q"].copy_(_dev(torch, to_tiled(q, tile))); views["v"].copy_(_dev(torch, to_tiled(v, tile)))
    views["a"].copy_(_dev(torch, to_tiled(a, tile)))
    ff = bool(I.has_freeflyer_root)
    views["qj"].copy_(_dev(torch, to_tiled(np.ascontiguousarray(q[:, 7:] if ff else q), tile)))
    views["dq"].copy_(_dev(torch, to_tiled(np.ascontiguousarray(v[:, 6:] if ff else v), tile)))
    views["root"].copy_(_dev(torch, to_tiled(np.ascontiguousarray(q[:, :7]), tile)) if ff else torch.zeros_like(views["root"]))
    views["rootv"].copy_(_dev(torch, to_tiled(np.ascontiguousarray(v[:, :6]), tile)) if ff else torch.zeros_like(views["rootv"]))
    views["w"].copy_(torch.rand_like(views["w"])); views["tj"].zero_(); views["tr"].zero_()
    ws.eval_soa(n, tile, views["q"], views["v"], views["a"], views["oMi"], views["J"], views["M"], views["tau"])
    ws.apply_soa(n, tile, views["qj"], views["root"] if ff else None, views["pos"])
    ws.applyJ_soa(n, tile, views["dq"], views["rootv"] if ff else None, views["vel"])
    ws.applyJT_soa(n, tile, views["w"], views["tj"], views["tr"] if ff else None)
    # ... and the later entry points on the same arena: Mass / ForceField products, geometric stiffness, constraint rows
    # and dense frame Jacobians (their row-major outputs get their own guarded arena)
    ws.addMDx_soa(n, tile, views["q"], views["a"], 0.5, views["tau"])
    ws.addForce_soa(n, tile, views["q"], views["v"], views["tau"])
    ws.addForce_soa(n, tile, views["q"], None, views["tau"])
    mdl_ok = not any(int(d.jtype[i]) == mdl.J_FF and int(d.parents[i]) != 0 for i in range(1, d.njoints))
    if mdl_ok:
        ws.applyDJT_soa(n, tile, views["dq"], views["rootv"] if ff else None, views["w"], 1.0, views["tj"],
                        views["tr"] if ff else None)
        # forward dynamics (both kernels) and the inverse mass matrix: outputs land in tau / M of the same arena
        for fused in (True, False):
            ws.set_aba_fused(fused)
            ws.aba_soa(n, tile, views["q"], views["v"], views["a"], views["tau"])
        ws.minv_soa(n, tile, views["q"], views["M"])
        # derivatives of inverse dynamics, dense (with and without M) and support-only
        ws.rnea_derivatives_soa(n, tile, views["q"], views["v"], views["a"], views["DQ"], views["DV"], views["M"])
        ws.rnea_derivatives_soa(n, tile, views["q"], views["v"], views["a"], views["DQ"], views["DV"], None)
        ws.rnea_derivatives_soa_sparse(n, tile, views["q"], views["v"], views["a"], views["SQ"], views["SV"])
    nrows = n + 3
    rp = torch.arange(0, 2 * nrows + 1, 2, dtype=torch.int32, device="cuda")
    ri = (torch.arange(0, nrows, dtype=torch.int32, device="cuda") % n).contiguous()
    col = torch.randint(0, int(I.n_out), (2 * nrows,), dtype=torch.int32, device="cuda")
    val = torch.rand((2 * nrows, 6), dtype=torch.float64, device="cuda")
    arena2 = torch.full((2 * G + nrows * I.nv + G + nrows * 6 * I.nv,), SENT, dtype=torch.float64, device="cuda")
    rows = arena2[G:G + nrows * I.nv].view(nrows, I.nv)
    Jd = arena2[2 * G + nrows * I.nv:2 * G + nrows * I.nv + nrows * 6 * I.nv].view(nrows, I.nv, 6)
    keepf = torch.zeros((nrows,), dtype=torch.int32, device="cuda")
    for cached in (True, False):
        ws.set_rows_cache(cached)
        ws.applyJT_rows_dev(nrows, rp, ri, col, val, rows, keepf)
    ws.frame_jacobians_dev(nrows, ri, col[:nrows].contiguous(), Jd)
    ws.synchronize()
    for lo, hi in ((0, G), (G + nrows * I.nv, 2 * G + nrows * I.nv)):
        assert bool((arena2[lo:hi] == SENT).all()), "guard band of the row / Jacobian arena was written"
    assert not bool((rows == SENT).any()) and not bool((Jd == SENT).any())
    off = 0
    for k, K in [("", 0)] + list(sizes.items()):
        off += nt * K * tile
        band = arena[off:off + G]
        assert bool((band == SENT).all()), f"guard band after '{k}' was written"
        off += G
    # and the live part of every output was written
    for k in ("oMi", "J", "M", "tau", "pos", "vel") + (("DQ", "DV", "SQ", "SV") if mdl_ok and d.nv > 0 else ()):
        live = from_tiled(views[k].cpu().numpy(), n)
        assert not np.any(live == SENT), k


@pytest.mark.parametrize("nmoving,branch_prob,ff", [(40, 0.0, False), (62, 0.0, True), (62, 0.9, False), (50, 0.3, True)])
def test_deep_and_bushy_trees(torch_mod, api, nmoving, branch_prob, ff):
    """Storage planner corner cases: long serial chains (sweep state of ~1000 doubles: few resident warps, Y-stack too
    big for one thread's tensor-memory share) and very bushy trees (many branch records), up to RBD_MAX_JOINTS."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.random_tree(1000 + nmoving, nmoving, free_flyer=ff, all_types=True, branch_prob=branch_prob)
    n = 150
    q, v, a = mdl.random_state(d, n, seed=23)
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for tmem in (True, False):
        try:
            got, launches = _eval_gpu(torch_mod, api, d, q, v, a, 32, tmem=tmem)
        except api.RbdError as e:
            # without tensor memory the 62-joint chain's state exceeds one SM even for the split sweeps: a clean error
            assert e.code == -5 and not tmem and nmoving >= 60 and branch_prob == 0.0, str(e)
            continue
        # 3 = the fused sweep did not fit and ran as FK + Jacobian, CRBA and RNEA sweeps (the single-algorithm
        # instantiations have no oMi / J stores compiled in)
        assert launches in (1, 3)
        for k in ("oMi", "J", "M", "tau"):
            assert rel_err(got[k], ref[k]) <= TOL_FP64, (k, tmem)
    # mapping path on the same model
    torch = torch_mod
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, 32))
    qj, root, dq, rootv = _split(d, q, v)
    nt = (n + 31) // 32
    pos = torch.zeros((nt, 7 * d.n_out, 32), dtype=torch.float64, device="cuda")
    w = np.random.default_rng(1).uniform(-1, 1, (n, 6 * d.n_out))
    tau = torch.zeros((nt, d.nv_joints, 32), dtype=torch.float64, device="cuda")
    rt = torch.zeros((nt, 6, 32), dtype=torch.float64, device="cuda") if ff else None
    ws.apply_soa(n, 32, T(qj), T(root), pos)
    ws.applyJT_soa(n, 32, T(w), tau, rt)
    ws.synchronize()
    tau = from_tiled(tau.cpu().numpy(), n)
    orc = oracle.Oracle(d)
    for i in range(0, n, 29):
        orc.apply(qj[i], None if root is None else root[i])
        rtau, _rroot = orc.applyJT(w[i])
        assert rel_err(tau[i], rtau) <= TOL_FP64


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tmem", [True, False])
def test_eval_f32_mode_within_1e4(torch_mod, api, d, tmem):
    """Optional FP32 mode of the fused evaluation (BASELINE north_star: tolerance 1e-4 relative): float I/O and state,
    same kernel structure, against the FP64 oracle evaluated on the float-rounded inputs."""
    import oracle
    from helpers import TOL_FP32
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    n, tile = 1000, 32
    q, v, a = mdl.random_state(d, n, seed=19)
    q32, v32, a32 = (x.astype(np.float32) for x in (q, v, a))
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_eval_tmem(tmem)
    I = m.info
    nt = (n + tile - 1) // tile
    K = {"oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": I.nv}
    outs = {k: torch.full((nt, K[k], tile), float("nan"), dtype=torch.float32, device="cuda") for k in K}
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q32, v32, a32))
    ws.eval_soa_f32(n, tile, qs, vs, as_, outs["oMi"], outs["J"], outs["M"], outs["tau"])
    ws.synchronize()
    ref, _ = oracle.eval_batch(d, q32.astype(np.float64), v32.astype(np.float64), a32.astype(np.float64), nthreads=0)
    for k in K:
        got = from_tiled(outs[k].cpu().numpy(), n).astype(np.float64)
        assert rel_err(got, ref[k]) <= TOL_FP32, k


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 100])
def test_mass_forcefield_products(torch_mod, api, d, tile):
    """SOFA Mass::addMDx / ForceField::addForce in joint space (rbd_addMDx_soa, rbd_addForce_soa): accumulating,
    matrix-free.  res += f M(q) dx against the oracle's CRBA matrix; f -= rnea(q, v, 0); gravity only: f -= rnea(q, 0, 0),
    which is what the reference's UniformMass + applyJT path evaluates (URDFModelLoader.cpp:369-391)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    n = 500
    q, v, dx = mdl.random_state(d, n, seed=31)
    rng = np.random.default_rng(32)
    r0 = rng.uniform(-1, 1, (n, d.nv))
    zero = np.zeros_like(v)
    ref, _ = oracle.eval_batch(d, q, zero, zero, nthreads=0)
    nv = d.nv
    M = np.zeros((n, nv, nv))
    for c in range(nv):
        for r in range(c + 1):
            M[:, r, c] = M[:, c, r] = ref["M"][:, c * (c + 1) // 2 + r]
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: _dev(torch, to_tiled(x, tile))
    qs = T(q)
    res = T(r0); ws.addMDx_soa(n, tile, qs, T(dx), 0.75, res)
    f = T(r0); ws.addForce_soa(n, tile, qs, T(v), f)
    g = T(r0); ws.addForce_soa(n, tile, qs, None, g)
    ws.synchronize()
    assert rel_err(from_tiled(res.cpu().numpy(), n), r0 + 0.75 * np.einsum("nij,nj->ni", M, dx)) <= TOL_FP64
    bias, _ = oracle.eval_batch(d, q, v, zero, nthreads=0)
    assert rel_err(from_tiled(f.cpu().numpy(), n), r0 - bias["tau"]) <= TOL_FP64
    assert rel_err(from_tiled(g.cpu().numpy(), n), r0 - ref["tau"]) <= TOL_FP64


def test_mass_forcefield_host_entry_points(api):
    """rbd_addMDx_host / rbd_addForce_host: SOFA-side host vectors, accumulated in place through the chunk pipeline
    (n larger than one chunk, not a multiple of the tile)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.solo12()
    n = 20000 + 13
    q, v, dx = mdl.random_state(d, n, seed=41)
    rng = np.random.default_rng(42)
    r0 = rng.uniform(-1, 1, (n, d.nv))
    m = api.Model(d); ws = api.Workspace(m, n)
    res = r0.copy(); ws.addMDx_host(n, q, dx, -0.5, res)
    f = r0.copy(); ws.addForce_host(n, q, v, f)
    g = r0.copy(); ws.addForce_host(n, q, None, g)
    idx = np.r_[0:40, n - 40:n]
    zero = np.zeros_like(v[idx])
    ref, _ = oracle.eval_batch(d, q[idx], zero, zero, nthreads=0)
    nv = d.nv
    M = np.zeros((len(idx), nv, nv))
    for c in range(nv):
        for r in range(c + 1):
            M[:, r, c] = M[:, c, r] = ref["M"][:, c * (c + 1) // 2 + r]
    assert rel_err(res[idx], r0[idx] - 0.5 * np.einsum("nij,nj->ni", M, dx[idx])) <= TOL_FP64
    bias, _ = oracle.eval_batch(d, q[idx], v[idx], zero, nthreads=0)
    assert rel_err(f[idx], r0[idx] - bias["tau"]) <= TOL_FP64
    assert rel_err(g[idx], r0[idx] - ref["tau"]) <= TOL_FP64


@pytest.mark.parametrize("name", ["talos_reduced", "solo12", "panda7", "random_2", "random_3"] + [x.name for x in MODELS if x.name.startswith("schunk")])
def test_applyJT_ring_equals_register_path(torch_mod, api, name):
    """The three sweeps of the vector applyJT (ring kernel: depth stack or forward projection, chosen by the planner;
    register path) accumulate the same generalized forces: many tiles per warp so that the ring wraps across tile
    boundaries, a partial last tile, and a wrench array that is only 8-byte aligned (the ring needs 16: fallback)."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    cands = [m for m in MODELS if m.name == name]
    if not cands:
        pytest.skip(f"no fixture named {name}")
    d = cands[0]
    n, tile = 148 * 12 * 32 * 2 + 32 * 5 + 7, 32
    q, v, _ = mdl.random_state(d, n, seed=51)
    qj, root, _dq, _rv = _split(d, q, v)
    rng = np.random.default_rng(52)
    w = rng.uniform(-1, 1, (n, d.n_out * 6))
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")
    ws.apply_soa(n, tile, T(qj), T(root), pos)
    wt = T(w)
    # the same wrenches at an address that is 8 (mod 16)
    shifted = torch.empty(wt.numel() + 1, dtype=torch.float64, device="cuda")
    shifted[1:] = wt.reshape(-1)
    res = []
    for ring, src in ((True, wt), (False, wt), (True, shifted[1:])):
        stages = ws.set_applyJT_ring(ring)
        tau = torch.zeros((nt, d.nv_joints, tile), dtype=torch.float64, device="cuda")
        rt = torch.zeros((nt, 6, tile), dtype=torch.float64, device="cuda") if root is not None else None
        ws.applyJT_soa(n, tile, src, tau, rt)
        ws.synchronize()
        res.append((from_tiled(tau.cpu().numpy(), n), None if rt is None else from_tiled(rt.cpu().numpy(), n)))
    assert stages >= 0
    for tau_x, rt_x in res[1:]:
        assert rel_err(tau_x, res[0][0]) <= 1e-12
        if rt_x is not None:
            assert rel_err(rt_x, res[0][1]) <= 1e-12


def test_joint_space_mass_mirror(api):
    """mass.JointSpaceMass (SOFA Mass / ForceField names on the joint dofs) against the oracle, and against the
    reference's own route for the gravity term: applyJT of m_b g on the body-CoM outputs (URDFModelLoader.cpp:369-391)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200.mass import JointSpaceMass
    from sofa.rigidbodydynamics_b200.mapping import KinematicChainMapping
    d = mdl.panda7()
    n = 64
    q, v, dx = mdl.random_state(d, n, seed=61)
    mass = JointSpaceMass(d, n)
    M = mass.getMassMatrix(q)
    ref, _ = oracle.eval_batch(d, q, np.zeros_like(v), np.zeros_like(v), nthreads=0)
    res = np.zeros((n, d.nv)); mass.addMDx(res, q, dx, 2.0)
    assert rel_err(res, 2.0 * np.einsum("nij,nj->ni", M, dx)) <= TOL_FP64
    assert np.all(np.linalg.eigvalsh(M) > 0)                                   # SPD
    f = np.zeros((n, d.nv)); mass.addForce(f, q)                               # gravity only
    assert rel_err(f, -ref["tau"]) <= TOL_FP64
    fb = np.zeros((n, d.nv)); mass.addForce(fb, q, v)
    bias, _ = oracle.eval_batch(d, q, v, np.zeros_like(v), nthreads=0)
    assert rel_err(fb, -bias["tau"]) <= TOL_FP64
    # the reference's route: gravity wrench m_b g at every body-CoM output, mapped back by applyJT
    kcm = KinematicChainMapping(d, n)
    out = np.zeros((n, d.n_out, 7)); kcm.apply([out], [q], [])
    w = np.zeros((n, d.n_out, 6))
    for b in range(d.njoints):
        w[:, b, :3] = np.asarray(d.inertia).reshape(d.njoints, 10)[b, 0] * np.asarray(d.gravity)[:3]
    tau = np.zeros((n, d.nv)); kcm.applyJT([tau], [], [w])
    assert kcm.d_componentState == "Valid"
    assert rel_err(tau, f) <= TOL_FP64
    with pytest.raises(ValueError):
        mass.addMDx(res, q, dx[:, :3], 1.0)


@pytest.mark.parametrize("name", ["talos_reduced", "solo12", "random_3", "talos_reduced_subset", "schunk_svh_hand_right_ff_subset"])
def test_constraint_rows_cached_jacobians(torch_mod, api, name):
    """rbd_applyJT_rows_dev on the Jacobian cache (default) equals the per-row sweep, follows a new apply (the cache is
    refilled lazily after every apply, as the reference's data.J is by computeJointJacobians), and matches the oracle."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    d = [m for m in MODELS_MAP if m.name == name][0]
    n, tile = 300, 32
    rng = np.random.default_rng(71)
    nrows = 2 * n + 5
    row_inst = np.sort(rng.integers(0, n, nrows)).astype(np.int32)
    nnz_per = rng.integers(0, 4, nrows)
    row_ptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.int32)
    nnz = int(row_ptr[-1])
    col = rng.integers(0, d.n_out, nnz).astype(np.int32)
    val = rng.uniform(-1, 1, (nnz, 6))
    m = api.Model(d); ws = api.Workspace(m, n)
    D = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    d_rp, d_ri, d_col, d_val = D(row_ptr), D(row_inst), D(col), D(val)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")
    results = {}
    for seed in (81, 82):                                   # two configurations: the second apply must invalidate
        q, v, _ = mdl.random_state(d, n, seed=seed)
        qj, root, _dq, _rv = _split(d, q, v)
        ws.apply_soa(n, tile, T(qj), T(root), pos)
        for cached in (True, False):
            ws.set_rows_cache(cached)
            rows = torch.full((nrows, d.nv), float("nan"), dtype=torch.float64, device="cuda")
            keep = torch.zeros((nrows,), dtype=torch.int32, device="cuda")
            ws.applyJT_rows_dev(nrows, d_rp, d_ri, d_col, d_val, rows, keep)
            ws.synchronize()
            results[(seed, cached)] = (rows.cpu().numpy(), keep.cpu().numpy())
        a, b = results[(seed, True)], results[(seed, False)]
        assert rel_err(a[0], b[0]) <= 1e-12 and np.array_equal(a[1], b[1])
        orc = oracle.Oracle(d)
        for r in list(range(0, nrows, 37)) + [nrows - 1]:
            i = int(row_inst[r])
            orc.apply(qj[i], None if root is None else root[i])
            sl = slice(int(row_ptr[r]), int(row_ptr[r + 1]))
            ref_row, ref_keep = orc.applyJT_row([int(c) for c in col[sl]], val[sl])
            assert rel_err(a[0][r], ref_row) <= TOL_FP64 and int(a[1][r]) == int(ref_keep)
    assert rel_err(results[(81, True)][0], results[(82, True)][0]) > 1e-3   # different configurations, different rows


@pytest.mark.parametrize("name", ["talos_reduced", "solo12", "panda7", "solo12_subset", "panda7_subset"])
def test_frame_jacobians_from_cache(torch_mod, api, name):
    """rbd_getFrameJacobian_* = pinocchio::getFrameJacobian(LOCAL_WORLD_ALIGNED) as the reference calls it per output
    (KCM.inl:434-490): against the oracle's dense Jacobian, and J_k dq == applyJ, J_k^T w == applyJT (one wrench)."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    d = [m for m in MODELS_MAP if m.name == name][0]
    n, tile = 100, 32
    q, v, _ = mdl.random_state(d, n, seed=91)
    qj, root, dq, rootv = _split(d, q, v)
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")
    vel = torch.zeros((nt, 6 * d.n_out, tile), dtype=torch.float64, device="cuda")
    with pytest.raises(api.RbdError):
        ws.frame_jacobians_host([0], [0])                     # before apply: RBD_ERR_NO_APPLY
    ws.apply_soa(n, tile, T(qj), T(root), pos)
    ws.applyJ_soa(n, tile, T(dq), T(rootv), vel)
    ws.synchronize()
    vel = from_tiled(vel.cpu().numpy(), n).reshape(n, d.n_out, 6)
    rng = np.random.default_rng(92)
    inst = rng.integers(0, n, 40).astype(np.int32); outk = rng.integers(0, d.n_out, 40).astype(np.int32)
    outk[0] = 0                                               # universe output: zero Jacobian
    J = ws.frame_jacobians_host(inst, outk)                   # [40, 6, nv]
    Jdev = torch.full((40, d.nv, 6), float("nan"), dtype=torch.float64, device="cuda")
    ws.frame_jacobians_dev(40, torch.from_numpy(inst).cuda(), torch.from_numpy(outk).cuda(), Jdev)
    ws.synchronize()
    assert np.array_equal(Jdev.cpu().numpy().transpose(0, 2, 1), J)
    if int(d.frame_parent[int(d.out_frames[0])]) == 0:
        assert np.all(J[0] == 0.0)                        # an output attached to the universe has an empty support
    orc = oracle.Oracle(d)
    for t in range(40):
        i, k = int(inst[t]), int(outk[t])
        orc.apply(qj[i], None if root is None else root[i])
        assert rel_err(J[t], orc.frame_jacobian_lwa(int(d.out_frames[k]))) <= TOL_FP64
        full_dq = dq[i] if rootv is None else np.concatenate([rootv[i], dq[i]])
        assert rel_err(J[t] @ full_dq, vel[i, k]) <= TOL_FP64


@pytest.mark.parametrize("name", ["talos_reduced", "solo12", "panda7", "talos_reduced_subset", "schunk_svh_hand_right_ff_subset"])
def test_assembled_getJ(torch_mod, api, name):
    """rbd_getJ_host / _dev: the assembled mapping Jacobian the reference's getJ() does not provide
    (KinematicChainMapping.h:82-83 returns nullptr) — per selected output the non-zero columns of
    getFrameJacobian(LOCAL_WORLD_ALIGNED) (KCM.inl:434-490), against the oracle's dense Jacobians; the assembled matrix
    times dq equals applyJ for every instance."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    d = [m for m in MODELS_MAP if m.name == name][0]
    n, tile = 130, 32
    q, v, _ = mdl.random_state(d, n, seed=93)
    qj, root, dq, rootv = _split(d, q, v)
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")
    vel = torch.zeros((nt, 6 * d.n_out, tile), dtype=torch.float64, device="cuda")
    with pytest.raises(api.RbdError) as e:
        ws.getJ_host(1)
    assert e.value.code == -4                                 # RBD_ERR_NO_APPLY
    ws.apply_soa(n, tile, T(qj), T(root), pos)
    ws.applyJ_soa(n, tile, T(dq), T(rootv), vel)
    ws.synchronize()
    vel = from_tiled(vel.cpu().numpy(), n).reshape(n, d.n_out, 6)
    full_dq = dq if rootv is None else np.concatenate([rootv, dq], axis=1)
    # all outputs: J dq == applyJ, instance by instance
    blk, cols = m.getJ_pattern()
    val = ws.getJ_host(n)                                      # [n, nnz, 6]
    for k in range(d.n_out):
        cs = cols[blk[k]:blk[k + 1]]
        got = np.einsum("nce,nc->ne", val[:, blk[k]:blk[k + 1]], full_dq[:, cs])
        assert rel_err(got, vel[:, k]) <= TOL_FP64
    # a permuted selection, device variant, against the oracle
    rng = np.random.default_rng(94)
    sel = rng.permutation(d.n_out)[:max(1, d.n_out // 3)].astype(np.int32)
    blk, cols = m.getJ_pattern(sel)
    dval = torch.full((n, max(1, len(cols)), 6), float("nan"), dtype=torch.float64, device="cuda")
    ws.getJ_dev(n, sel, dval)
    ws.synchronize()
    dval = dval.cpu().numpy()
    assert np.array_equal(dval[:, :len(cols)], ws.getJ_host(n, sel))
    orc = oracle.Oracle(d)
    for i in (0, 77, n - 1):
        orc.apply(qj[i], None if root is None else root[i])
        for s_, k in enumerate(sel):
            cs = cols[blk[s_]:blk[s_ + 1]]
            Jref = orc.frame_jacobian_lwa(int(d.out_frames[int(k)]))
            assert rel_err(dval[i, blk[s_]:blk[s_ + 1]].T, Jref[:, cs]) <= TOL_FP64
            rest = np.ones(d.nv, dtype=bool); rest[cs] = False
            assert np.all(Jref[:, rest] == 0.0)
    with pytest.raises(api.RbdError):
        m.getJ_pattern(np.array([d.n_out], dtype=np.int32))   # output index out of range


@pytest.mark.parametrize("name,tile", [("talos_reduced", 32), ("solo12", 50), ("random_3", 32), ("panda7", 32)])
def test_applyDJT_geometric_stiffness(torch_mod, api, name, tile):
    """rbd_applyDJT_soa / _host (optional extension; the reference's applyDJT is a no-op and the mirror keeps it one):
    equals the central finite difference of applyJT along dq, accumulates with kfactor, host == device."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    torch = torch_mod
    d = [m for m in MODELS if m.name == name][0]
    n = 70
    q, v, _ = mdl.random_state(d, n, seed=101)
    rng = np.random.default_rng(102)
    w = rng.uniform(-1, 1, (n, 6 * d.n_out))
    ff = bool(d.has_freeflyer_root)
    m = api.Model(d); ws = api.Workspace(m, n)
    T = lambda x: None if x is None else _dev(torch, to_tiled(x, tile))
    nt = (n + tile - 1) // tile
    pos = torch.zeros((nt, 7 * d.n_out, tile), dtype=torch.float64, device="cuda")

    def jt(qx):
        qj, root, _, _ = _split(d, qx, v)
        ws.apply_soa(n, tile, T(qj), T(root), pos)
        tau = torch.zeros((nt, d.nv_joints, tile), dtype=torch.float64, device="cuda")
        rt = torch.zeros((nt, 6, tile), dtype=torch.float64, device="cuda") if ff else None
        ws.applyJT_soa(n, tile, T(w), tau, rt)
        ws.synchronize()
        tj = from_tiled(tau.cpu().numpy(), n)
        return np.concatenate([from_tiled(rt.cpu().numpy(), n), tj], axis=1) if ff else tj

    eps = 1e-6
    fd = (jt(helpers.integrate(d, q, v, eps)) - jt(helpers.integrate(d, q, v, -eps))) / (2 * eps)
    jt(q)                                                     # the cached configuration is q again
    _, _, dq, rootv = _split(d, q, v)
    out = T(np.ones((n, d.nv_joints))); ort = T(np.ones((n, 6))) if ff else None
    ws.applyDJT_soa(n, tile, T(dq), T(rootv), T(w), -2.0, out, ort)
    ws.synchronize()
    got = from_tiled(out.cpu().numpy(), n)
    if ff:
        got = np.concatenate([from_tiled(ort.cpu().numpy(), n), got], axis=1)
    assert rel_err(got, 1.0 - 2.0 * fd) <= 1e-6
    # host entry point (instance-major, accumulated in place) gives the same numbers
    h_out = np.ones((n, d.nv_joints)); h_rt = np.ones((n, 6)) if ff else None
    ws.applyDJT_host(n, dq, rootv, w, -2.0, h_out, h_rt)
    h = np.concatenate([h_rt, h_out], axis=1) if ff else h_out
    assert rel_err(h, got) <= 1e-13


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 1000])
@pytest.mark.parametrize("fused", [1, 0])
def test_aba_matches_oracle(torch_mod, api, d, tile, fused):
    """rbd_aba_soa (forward dynamics) against the oracle's articulated-body algorithm
    (pinocchio's LOCAL-convention restatement), the mpmath forward-dynamics golden vectors (M^-1 (tau - nle) solved in
    50 digits by the independent implementation) and the round trip through the GPU's own RNEA.  Both kernels: the fused
    sweep on on-chip level records (aba2_kernel, the default) and the first version (aba_kernel)."""
    import oracle
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    with open(os.path.join(helpers.GOLDEN, "golden_aba.json")) as f:
        gold = json.load(f)[d.name]["states"]
    n = 333
    q, v, a = mdl.random_state(d, n, seed=61)
    rng = np.random.default_rng(62)
    tau = rng.uniform(-2, 2, (n, d.nv))
    for s_, st in enumerate(gold):
        q[s_], v[s_], tau[s_] = np.array(st["q"]), np.array(st["v"]), np.array(st["tau"])
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_aba_fused(bool(fused))
    nt = (n + tile - 1) // tile
    new = lambda: torch.full((nt, max(1, d.nv), tile), float("nan"), dtype=torch.float64, device="cuda")
    qs, vs, ts, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, tau, a))
    ddq = new()
    ws.aba_soa(n, tile, qs, vs, ts, ddq)
    ws.synchronize()
    assert ("aba2_kernel" if fused else "aba_kernel") in api.last_kernel_name()
    got = from_tiled(ddq.cpu().numpy(), n)[:, :d.nv]
    orc = oracle.Oracle(d)
    idx = list(range(len(gold))) + list(range(len(gold), n, 17)) + [n - 1]
    ref = np.stack([orc.aba(q[i], v[i], tau[i]) for i in idx])
    assert rel_err(got[idx], ref) <= 1e-9                     # conditioning of M(q) enters: not 1e-10
    for s_, st in enumerate(gold):
        assert rel_err(got[s_], np.array(st["qdd"])) <= 1e-9
    # round trip through the GPU's RNEA for every instance
    t2 = new()
    ws.rnea_soa(n, tile, qs, vs, as_, t2)
    d2 = new()
    ws.aba_soa(n, tile, qs, vs, t2, d2)
    ws.synchronize()
    assert rel_err(from_tiled(d2.cpu().numpy(), n)[:, :d.nv], a) <= 1e-8
    # host entry point
    hd = np.zeros((n, d.nv))
    ws.aba_host(n, q, v, tau, hd)
    assert rel_err(hd, got) <= 1e-13


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 1000])
def test_minv_matches_oracle(torch_mod, api, d, tile):
    """rbd_minv_soa / rbd_minv_host (pinocchio::computeMinverse): the packed upper triangle of
    M(q)^-1 against the inverse of the oracle's CRBA mass matrix, against M^-1 M == I with the GPU's own CRBA, and as the
    solution operator of the GPU's forward dynamics (aba(q, 0, tau) - aba(q, 0, 0) == M^-1 tau)."""
    import oracle
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    if d.nv == 0:
        pytest.skip("no dofs")
    n = 333
    q, _v, _a = mdl.random_state(d, n, seed=71)
    m = api.Model(d); ws = api.Workspace(m, n)
    nt = (n + tile - 1) // tile
    mf = d.nv * (d.nv + 1) // 2
    qs = _dev(torch, to_tiled(q, tile))
    Mi = torch.full((nt, mf, tile), float("nan"), dtype=torch.float64, device="cuda")
    ws.minv_soa(n, tile, qs, Mi)
    ws.synchronize()
    assert "minv_kernel" in api.last_kernel_name()
    got = from_tiled(Mi.cpu().numpy(), n)[:, :mf]
    assert np.isfinite(got).all()
    iu = np.triu_indices(d.nv)

    def full(packed):
        G = np.zeros((d.nv, d.nv))
        for c in range(d.nv):
            G[:c + 1, c] = packed[c * (c + 1) // 2: c * (c + 1) // 2 + c + 1]
        return np.triu(G) + np.triu(G, 1).T

    orc = oracle.Oracle(d)
    for i in list(range(0, n, 23)) + [n - 1]:
        ref = orc.minv(q[i])
        assert np.max(np.abs(full(got[i]) - ref)) <= 1e-9 * max(1.0, float(np.max(np.abs(ref))))
    # M^-1 M == I with the GPU's CRBA, every instance
    Mp = torch.full((nt, mf, tile), float("nan"), dtype=torch.float64, device="cuda")
    ws.crba_soa(n, tile, qs, Mp)
    ws.synchronize()
    Mg = from_tiled(Mp.cpu().numpy(), n)[:, :mf]
    worst = 0.0
    for i in range(n):
        worst = max(worst, float(np.max(np.abs(full(got[i]) @ full(Mg[i]) - np.eye(d.nv)))))
    assert worst <= 1e-8
    # solution operator of the GPU's forward dynamics (gravity cancels in the difference)
    rng = np.random.default_rng(72)
    tau = rng.uniform(-1, 1, (n, d.nv)); zero = np.zeros((n, d.nv))
    new = lambda: torch.full((nt, max(1, d.nv), tile), float("nan"), dtype=torch.float64, device="cuda")
    zs, ts = _dev(torch, to_tiled(zero, tile)), _dev(torch, to_tiled(tau, tile))
    a1, a0 = new(), new()
    ws.aba_soa(n, tile, qs, zs, ts, a1); ws.aba_soa(n, tile, qs, zs, zs, a0)
    ws.synchronize()
    da = from_tiled((a1 - a0).cpu().numpy(), n)[:, :d.nv]
    sol = np.stack([full(got[i]) @ tau[i] for i in range(n)])
    assert rel_err(da, sol) <= 1e-8
    # host entry point (instance-major packed upper triangle)
    hm = np.zeros((n, mf))
    ws.minv_host(n, q, hm)
    assert rel_err(hm, got) <= 1e-13


@pytest.mark.parametrize("d", MODELS, ids=IDS)
@pytest.mark.parametrize("tile", [32, 1000])
def test_rnea_derivatives_match_oracle(torch_mod, api, d, tile):
    """rbd_rnea_derivatives_soa / _host (pinocchio::computeRNEADerivatives): dtau_dq, dtau_dv and
    the optional M against the oracle's dense restatement, the 50-digit finite-difference golden vectors, exact structural
    zeros, M against the GPU's own CRBA, and a directional finite difference of the GPU's own rnea for every instance."""
    import json
    import os
    import helpers
    import oracle
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    if d.nv == 0:
        pytest.skip("no dofs")
    with open(os.path.join(helpers.GOLDEN, "golden_drnea.json")) as f:
        gold = json.load(f)[d.name]
    n = 333
    nv = d.nv
    q, v, a = mdl.random_state(d, n, seed=73)
    q[0], v[0], a[0] = np.array(gold["q"]), np.array(gold["v"]), np.array(gold["a"])
    m = api.Model(d); ws = api.Workspace(m, n)
    nt = (n + tile - 1) // tile
    mf = nv * (nv + 1) // 2
    new = lambda K: torch.full((nt, K, tile), float("nan"), dtype=torch.float64, device="cuda")
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    DQ, DV, Mp = new(nv * nv), new(nv * nv), new(mf)
    ws.rnea_derivatives_soa(n, tile, qs, vs, as_, DQ, DV, Mp)
    ws.synchronize()
    assert "drnea_kernel" in api.last_kernel_name()
    gq = from_tiled(DQ.cpu().numpy(), n)[:, :nv * nv].reshape(n, nv, nv)
    gv = from_tiled(DV.cpu().numpy(), n)[:, :nv * nv].reshape(n, nv, nv)
    gM = from_tiled(Mp.cpu().numpy(), n)[:, :mf]
    assert np.isfinite(gq).all() and np.isfinite(gv).all() and np.isfinite(gM).all()
    related = helpers.support_relation(d)
    assert np.all(gq[:, ~related] == 0.0) and np.all(gv[:, ~related] == 0.0)
    orc = oracle.Oracle(d)
    for i in list(range(0, n, 23)) + [n - 1]:
        rq, rv, _ = orc.rnea_derivatives(q[i], v[i], a[i])
        assert rel_err(gq[i], rq) <= TOL_FP64 and rel_err(gv[i], rv) <= TOL_FP64, i
    assert rel_err(gq[0], np.array(gold["dtau_dq"])) <= TOL_FP64 and rel_err(gv[0], np.array(gold["dtau_dv"])) <= TOL_FP64
    # dtau_da == the CRBA kernel's M (same packed format)
    Mc = new(mf)
    ws.crba_soa(n, tile, qs, Mc); ws.synchronize()
    assert rel_err(gM, from_tiled(Mc.cpu().numpy(), n)[:, :mf]) <= 1e-12
    # without M: same matrices
    DQ2, DV2 = new(nv * nv), new(nv * nv)
    ws.rnea_derivatives_soa(n, tile, qs, vs, as_, DQ2, DV2, None); ws.synchronize()
    assert np.array_equal(from_tiled(DQ2.cpu().numpy(), n), from_tiled(DQ.cpu().numpy(), n))
    assert np.array_equal(from_tiled(DV2.cpu().numpy(), n), from_tiled(DV.cpu().numpy(), n))
    # directional central difference of the GPU's own rnea along (dq, dv), every instance
    rng = np.random.default_rng(74)
    dq = rng.uniform(-1, 1, (n, nv)); dv = rng.uniform(-1, 1, (n, nv))
    eps = 1e-6
    taus = []
    for sgn in (1.0, -1.0):
        qe = helpers.integrate(d, q, dq, sgn * eps); ve = v + sgn * eps * dv
        t = new(max(1, nv))
        ws.rnea_soa(n, tile, _dev(torch, to_tiled(qe, tile)), _dev(torch, to_tiled(ve, tile)), as_, t)
        ws.synchronize()
        taus.append(from_tiled(t.cpu().numpy(), n)[:, :nv])
    fd = (taus[0] - taus[1]) / (2 * eps)
    lin = np.einsum("nrc,nc->nr", gq, dq) + np.einsum("nrc,nc->nr", gv, dv)
    assert np.max(np.abs(fd - lin)) <= 1e-6 * max(1.0, float(np.max(np.abs(lin))))
    # support-only format (rbd_model_drnea_pattern + *_sparse): the pattern is the support relation, the values are the
    # dense entries at the pattern's positions bit for bit (same sweep, other addresses), device and host entry points
    row_ptr, col_idx = m.drnea_pattern()
    nnz = len(col_idx)
    assert nnz == int(related.sum()) == row_ptr[nv]
    rows = np.repeat(np.arange(nv), np.diff(row_ptr))
    assert np.array_equal(np.stack([rows, col_idx], 1), np.argwhere(related))
    SQ, SV = new(max(1, nnz)), new(max(1, nnz))
    ws.rnea_derivatives_soa_sparse(n, tile, qs, vs, as_, SQ, SV); ws.synchronize()
    assert "drnea_kernel" in api.last_kernel_name()
    assert np.array_equal(from_tiled(SQ.cpu().numpy(), n)[:, :nnz], gq[:, rows, col_idx])
    assert np.array_equal(from_tiled(SV.cpu().numpy(), n)[:, :nnz], gv[:, rows, col_idx])
    hsq, hsv = np.zeros((n, nnz)), np.zeros((n, nnz))
    ws.rnea_derivatives_host_sparse(n, q, v, a, hsq, hsv)
    assert rel_err(hsq, gq[:, rows, col_idx]) <= 1e-13 and rel_err(hsv, gv[:, rows, col_idx]) <= 1e-13
    # derivative of the CRBA mass matrix, contracted with a: d(M(q) a)/dq = dtau_dq(q, 0, a) - dtau_dq(q, 0, 0) (the second
    # term is pinocchio::computeGeneralizedGravityDerivatives) against a directional central difference of the CRBA kernel
    zs = _dev(torch, to_tiled(np.zeros((n, nv)), tile))
    Da, D0, Dv_ = new(nv * nv), new(nv * nv), new(nv * nv)
    ws.rnea_derivatives_soa(n, tile, qs, zs, as_, Da, Dv_, None)
    ws.rnea_derivatives_soa(n, tile, qs, zs, zs, D0, Dv_, None)
    ws.synchronize()
    dMa = (from_tiled(Da.cpu().numpy(), n)[:, :nv * nv] - from_tiled(D0.cpu().numpy(), n)[:, :nv * nv]).reshape(n, nv, nv)
    Ms = []
    for sgn in (1.0, -1.0):
        Me = new(mf)
        ws.crba_soa(n, tile, _dev(torch, to_tiled(helpers.integrate(d, q, dq, sgn * eps), tile)), Me); ws.synchronize()
        Ms.append(np.stack([oracle.unpack_upper(x, nv) for x in from_tiled(Me.cpu().numpy(), n)[:, :mf]]))
    fdM = np.einsum("nrc,nc->nr", (Ms[0] - Ms[1]) / (2 * eps), a)
    linM = np.einsum("nrc,nc->nr", dMa, dq)
    assert np.max(np.abs(fdM - linM)) <= 1e-6 * max(1.0, float(np.max(np.abs(linM))))
    # host entry point (instance-major [n][nv][nv])
    hq, hv, hM = np.zeros((n, nv, nv)), np.zeros((n, nv, nv)), np.zeros((n, mf))
    ws.rnea_derivatives_host(n, q, v, a, hq, hv, hM)
    assert rel_err(hq, gq) <= 1e-13 and rel_err(hv, gv) <= 1e-13 and rel_err(hM, gM) <= 1e-13
    hq2, hv2 = np.zeros((n, nv, nv)), np.zeros((n, nv, nv))
    ws.rnea_derivatives_host(n, q, v, a, hq2, hv2, None)
    assert np.array_equal(hq, hq2) and np.array_equal(hv, hv2)


@pytest.mark.parametrize("nmoving,free_flyer", [(40, False), (45, True), (63, False)])
def test_rnea_derivatives_deep_chains(torch_mod, api, nmoving, free_flyer):
    """Long serial chains through the C ABI: the hot level records force blocks of fewer than 8 warps (more shared and
    tensor memory per thread); a 63-joint chain fits no block at all: RBD_ERR_UNSUPPORTED, not a wrong result."""
    import oracle
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.random_tree(100 + nmoving, nmoving, free_flyer=free_flyer, branch_prob=0.0, all_types=False)
    n, tile, nv = 100, 32, d.nv
    q, v, a = mdl.random_state(d, n, seed=3)
    m = api.Model(d); ws = api.Workspace(m, n)
    nt = (n + tile - 1) // tile
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    DQ = torch.full((nt, nv * nv, tile), float("nan"), dtype=torch.float64, device="cuda")
    DV = torch.full((nt, nv * nv, tile), float("nan"), dtype=torch.float64, device="cuda")
    if nmoving == 63:
        with pytest.raises(api.RbdError) as e:
            ws.rnea_derivatives_soa(n, tile, qs, vs, as_, DQ, DV, None)
        assert e.value.code == -5
        return
    ws.rnea_derivatives_soa(n, tile, qs, vs, as_, DQ, DV, None)
    ws.synchronize()
    gq = from_tiled(DQ.cpu().numpy(), n)[:, :nv * nv].reshape(n, nv, nv)
    gv = from_tiled(DV.cpu().numpy(), n)[:, :nv * nv].reshape(n, nv, nv)
    assert np.isfinite(gq).all() and np.isfinite(gv).all()
    orc = oracle.Oracle(d)
    for i in (0, 31, 32, n - 1):
        rq, rv, _ = orc.rnea_derivatives(q[i], v[i], a[i])
        assert rel_err(gq[i], rq) <= 1e-9 and rel_err(gv[i], rv) <= 1e-9, i


def test_dynamics_entry_points_refuse_a_free_flyer_below_the_root(torch_mod, api):
    """aba / minv / rnea_derivatives support a free-flyer as joint 1 on the universe only: anything else is
    RBD_ERR_UNSUPPORTED with a message, not a wrong result (the fused eval and the mapping path do support it)."""
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    rng = np.random.default_rng(0)
    b = mdl._Builder("ff_below_root")
    b.add_frame("root_joint", 0, mdl._se3(), mdl.F_FIXED_JOINT)
    j1 = b.add_joint(0, mdl.J_RZ, mdl._se3(), "j1")
    b.append_body(j1, mdl._synthetic_inertia(rng), mdl._se3())
    j2 = b.add_joint(j1, mdl.J_FF, mdl._se3(), "j2")
    b.append_body(j2, mdl._synthetic_inertia(rng), mdl._se3())
    d = b.finish(gravity=(0.0, 0.0, -9.81), note="test")
    n, tile = 32, 32
    q, v, a = mdl.random_state(d, n, seed=1)
    m = api.Model(d); ws = api.Workspace(m, n)
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    new = lambda K: torch.zeros((1, K, tile), dtype=torch.float64, device="cuda")
    for call in (lambda: ws.aba_soa(n, tile, qs, vs, as_, new(d.nv)),
                 lambda: ws.minv_soa(n, tile, qs, new(d.nv * (d.nv + 1) // 2)),
                 lambda: ws.rnea_derivatives_soa(n, tile, qs, vs, as_, new(d.nv * d.nv), new(d.nv * d.nv), None)):
        with pytest.raises(api.RbdError) as e:
            call()
        assert e.value.code == -5 and "free-flyer" in str(e.value)   # RBD_ERR_UNSUPPORTED
    tau = new(d.nv)
    ws.rnea_soa(n, tile, qs, vs, as_, tau); ws.synchronize()           # the eval path takes the model
    assert torch.isfinite(tau).all()


@pytest.mark.parametrize("tile", [32, 1000])
def test_arm_kernel_equals_generic_chain_sweep(torch_mod, api, tile):
    """Fused eval of the 7-joint Panda chain: the unrolled arm kernel (default for 6- / 7-joint serial chains) against the
    generic chain instantiation (rbd_workspace_set_eval_arm(0)) and the oracle, all outputs, M in both formats, oMi / J
    optional; partial last tile."""
    import oracle
    torch = torch_mod
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.panda7()
    n = 1000 + 7
    q, v, a = mdl.random_state(d, n, seed=91)
    m = api.Model(d); ws = api.Workspace(m, n)
    I = m.info
    nt = (n + tile - 1) // tile
    new = lambda K: torch.full((nt, K, tile), float("nan"), dtype=torch.float64, device="cuda")
    qs, vs, as_ = (_dev(torch, to_tiled(x, tile)) for x in (q, v, a))
    res = {}
    for arm in (1, 0):
        ws.set_eval_arm(bool(arm))
        o = {"oMi": new(I.omi_fields), "J": new(I.jac_fields), "M": new(I.mass_fields), "tau": new(I.nv)}
        ws.eval_soa(n, tile, qs, vs, as_, o["oMi"], o["J"], o["M"], o["tau"])
        ws.synchronize()
        assert ("eval_arm_kernel<7" in api.last_kernel_name()) == bool(arm), api.last_kernel_name()
        res[arm] = {k: from_tiled(t.cpu().numpy(), n) for k, t in o.items()}
    ref, _ = oracle.eval_batch(d, q, v, a, nthreads=0)
    for k in ref:
        assert rel_err(res[1][k][:, :ref[k].shape[1]], ref[k]) <= TOL_FP64, k
        assert rel_err(res[1][k], res[0][k]) <= 1e-13, k
    ws.set_eval_arm(True)
    Mv, t2 = new(I.mass_support_fields), new(I.nv)
    assert I.mass_support_fields == I.mass_fields                      # a chain's M has no structural zeros
    ws.eval_soa_sparse(n, tile, qs, vs, as_, None, None, Mv, t2)
    ws.synchronize()
    assert "eval_arm_kernel<7" in api.last_kernel_name()
    assert np.array_equal(from_tiled(Mv.cpu().numpy(), n), res[1]["M"]) and np.array_equal(from_tiled(t2.cpu().numpy(), n), res[1]["tau"])
    # M alone / tau alone are not the arm kernel's job
    ws.crba_soa(n, tile, qs, Mv); ws.synchronize()
    assert "eval_arm_kernel" not in api.last_kernel_name()
    # the vector applyJT of the chain: unrolled register sweep (applyJT_arm_kernel, full tiles of a tile-32 batch) against
    # the generic ring sweep and the oracle, accumulating into a non-zero tau
    if tile == 32:
        rng = np.random.default_rng(5)
        w = rng.uniform(-1, 1, (n, 6 * I.n_out)); tau0 = rng.uniform(-1, 1, (n, I.nv))
        pos = new(7 * I.n_out)
        ws.apply_soa(n, tile, qs, None, pos)
        ws_t = {}
        for arm in (1, 0):
            ws.set_eval_arm(bool(arm))
            t = _dev(torch, to_tiled(tau0, tile))
            ws.applyJT_soa(n, tile, _dev(torch, to_tiled(w, tile)), t, None)
            ws.synchronize()
            ws_t[arm] = from_tiled(t.cpu().numpy(), n)
        ws.set_eval_arm(True)
        assert rel_err(ws_t[1], ws_t[0]) <= 1e-12
        orc = oracle.Oracle(d)
        for i in (0, 517, n - 1):
            orc.apply(q[i])
            ref, _ = orc.applyJT(w[i].reshape(I.n_out, 6), out_tau=tau0[i])
            assert rel_err(ws_t[1][i], ref) <= TOL_FP64

"""Property test over random kinematic trees (CPU): for many random shapes (chains, bushes, free-flyer roots, all joint
types) the kernel bodies — both storage plans, FP64 and FP32 — agree with the oracle, and the planner's output obeys
the B200 budgets it promises."""
import ctypes as C

import numpy as np
import pytest

from helpers import P, TOL_FP32, TOL_FP64, from_tiled, rel_err, to_tiled


def _cdesc(d):
    from sofa.rigidbodydynamics_b200._capi import make_desc
    return make_desc(d)


@pytest.mark.parametrize("seed", range(24))
def test_random_tree_eval_and_mapping(emul, seed):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    rng = np.random.default_rng(1000 + seed)
    nmoving = int(rng.integers(1, 45))
    ff = bool(rng.integers(0, 2))
    d = mdl.random_tree(5000 + seed, nmoving, free_flyer=ff, all_types=bool(rng.integers(0, 2)),
                        branch_prob=float(rng.choice([0.0, 0.15, 0.4, 0.8])), fixed_frames=int(rng.integers(0, 6)))
    n, tile = 40, 32
    q, v, a = mdl.random_state(d, n, seed=seed)
    cd, keep = _cdesc(d)
    info = (C.c_int * 8)()
    rc = emul.emul_model_info(C.addressof(cd), info)
    ref, _ = oracle.eval_batch(d, q, v, a, 1)
    nt = (n + tile - 1) // tile
    K = {"oMi": 12 * (d.njoints - 1), "J": 6 * d.nv, "M": d.nv * (d.nv + 1) // 2, "tau": d.nv}
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    ran = 0
    for use_tmem in (0, 1):
        outs = {k: np.full((nt, K[k], tile), np.nan) for k in K}
        rc = emul.emul_eval(C.addressof(cd), n, tile, 0, use_tmem, P(qs), P(vs), P(as_), P(outs["oMi"]), P(outs["J"]),
                            P(outs["M"]), P(outs["tau"]))
        if rc == -100:
            # the fused state does not fit one SM with this storage: the library then runs the two halves
            o2 = {k: np.full((nt, K[k], tile), np.nan) for k in K}
            r1 = emul.emul_eval(C.addressof(cd), n, tile, 0, use_tmem, P(qs), P(vs), P(as_), P(o2["oMi"]), P(o2["J"]),
                                P(o2["M"]), None)
            r2 = emul.emul_eval(C.addressof(cd), n, tile, 0, use_tmem, P(qs), P(vs), P(as_), None, None, None, P(o2["tau"]))
            if r1 != 0 or r2 != 0:
                continue            # does not fit at all (very deep chain without tensor memory): a clean refusal
            outs = o2
        else:
            assert rc == 0, emul.emul_last_error()
        ran += 1
        for k in K:
            assert rel_err(from_tiled(outs[k], n), ref[k]) <= TOL_FP64, (k, use_tmem, d.njoints)
    assert ran >= 1
    # FP32 mode
    o32 = {k: np.full((nt, K[k], tile), np.nan, dtype=np.float32) for k in K}
    q32, v32, a32 = (x.astype(np.float32) for x in (q, v, a))
    t32 = [to_tiled(x, tile) for x in (q32, v32, a32)]
    rc = emul.emul_eval_f32(C.addressof(cd), n, tile, 0, 1, P(t32[0]), P(t32[1]), P(t32[2]), P(o32["oMi"]), P(o32["J"]),
                            P(o32["M"]), P(o32["tau"]))
    if rc == 0:
        ref32, _ = oracle.eval_batch(d, q32.astype(np.float64), v32.astype(np.float64), a32.astype(np.float64), 1)
        for k in K:
            assert rel_err(from_tiled(o32[k], n).astype(np.float64), ref32[k]) <= 5 * TOL_FP32, k   # long chains: 5e-4
    # applyJT through the storage plan == dense oracle product
    if d.has_freeflyer_root:
        qj, root = np.ascontiguousarray(q[:, 7:]), np.ascontiguousarray(q[:, :7])
    else:
        qj, root = q, None
    qcache = to_tiled(q, 32)
    w = rng.uniform(-1, 1, (n, 6 * d.n_out))
    tw = to_tiled(w, tile)
    ttau = to_tiled(np.zeros((n, d.nv_joints)), tile)
    trt = to_tiled(np.zeros((n, 6)), tile) if root is not None else None
    assert emul.emul_applyJT_plan(C.addressof(cd), n, tile, 1, P(qcache), 32, P(tw), P(ttau), P(trt)) == 0
    orc = oracle.Oracle(d)
    tau = from_tiled(ttau, n)
    for i in (0, n - 1):
        orc.apply(qj[i], None if root is None else root[i])
        rtau, _ = orc.applyJT(w[i])
        assert rel_err(tau[i], rtau) <= TOL_FP64


def test_planner_budgets(emul):
    """Every plan the planner returns fits the B200 per-SM budgets it was given."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    info = (C.c_int * 8)()
    for seed in range(40):
        rng = np.random.default_rng(seed)
        d = mdl.random_tree(7000 + seed, int(rng.integers(1, 63)), free_flyer=bool(rng.integers(0, 2)),
                            branch_prob=float(rng.choice([0.0, 0.2, 0.6])))
        cd, keep = _cdesc(d)
        rc = emul.emul_model_info(C.addressof(cd), info)
        if rc != 0:
            continue
        smem_d, tmem_d, warps = info[2], info[3], info[4]
        assert warps >= 1 and smem_d >= 0 and 0 <= tmem_d <= 256
        assert smem_d * 8 * 32 <= 227 * 1024

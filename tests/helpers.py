"""Shared test helpers: tiled-SoA packing, error metric, the CPU emulation shim of the kernel bodies."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")

TOL_FP64 = 1e-10   # BASELINE.json north_star: <= 1e-10 relative error in FP64
TOL_FP32 = 1e-4    # ... and the optional FP32 mode: <= 1e-4


def to_tiled(x: np.ndarray, tile: int) -> np.ndarray:
    """[n, K] instance-major -> [ntiles, K, tile] tiled SoA (include/rbd_b200.h 'Memory layouts')."""
    n, K = x.shape
    nt = (n + tile - 1) // tile
    o = np.zeros((nt * tile, K), dtype=x.dtype)
    o[:n] = x
    return np.ascontiguousarray(o.reshape(nt, tile, K).transpose(0, 2, 1))


def from_tiled(o: np.ndarray, n: int) -> np.ndarray:
    nt, K, tile = o.shape
    return np.ascontiguousarray(o.transpose(0, 2, 1).reshape(nt * tile, K)[:n])


def rel_err(x, ref) -> float:
    """SURVEY.md §8(c) metric: max|x - ref| / max(1, max|ref|) over one output block."""
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.max(np.abs(x - ref)) / max(1.0, float(np.max(np.abs(ref)))))


def quat_err(qa, qb) -> float:
    """Quaternion mismatch up to sign (q and -q are the same rotation)."""
    qa = np.asarray(qa).reshape(-1, 4); qb = np.asarray(qb).reshape(-1, 4)
    return float(np.max(np.minimum(np.abs(qa - qb).max(axis=1), np.abs(qa + qb).max(axis=1)))) if qa.size else 0.0


def load_emul() -> C.CDLL:
    """tests/emul/libemul.so: the kernel bodies of csrc/rbd_sweeps.h compiled by g++ (test-only)."""
    src = os.path.join(HERE, "emul", "emul.cpp")
    out = os.path.join(HERE, "emul", "libemul.so")
    csrc = os.path.join(ROOT, "sofa", "rigidbodydynamics_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("rbd_sweeps.h", "rbd_sweeps_body.h", "rbd_dyn_body.h", "rbd_deriv_body.h", "rbd_arm_body.h", "rbd_model_build.h", "rbd_dev_model.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-march=native", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", out, src])
    lib = C.CDLL(out)
    vp, ll, i = C.c_void_p, C.c_longlong, C.c_int
    lib.emul_last_error.restype = C.c_char_p
    lib.emul_model_info.argtypes = [vp, vp]
    lib.emul_sincos.argtypes = [ll, vp, vp, vp]
    lib.emul_sincos.restype = None
    lib.emul_eval.argtypes = [vp, ll, ll, i, i] + [vp] * 7
    lib.emul_eval_ex.argtypes = [vp, ll, ll, i, i] + [vp] * 7 + [i, i]
    lib.emul_eval_variant.argtypes = [vp, i, i, vp]
    lib.emul_mass_pattern.argtypes = [vp, vp, vp]
    lib.emul_eval_f32.argtypes = [vp, ll, ll, i, i] + [vp] * 7
    lib.emul_apply.argtypes = [vp, ll, ll, i, vp, vp, vp, ll, vp]
    lib.emul_applyJ.argtypes = [vp, ll, ll, i, vp, ll, vp, vp, vp]
    lib.emul_applyJT.argtypes = [vp, ll, ll, i, vp, ll, vp, vp, vp]
    lib.emul_applyJT_plan.argtypes = [vp, ll, ll, i, vp, ll, vp, vp, vp]
    lib.emul_applyJT_ring.argtypes = [vp, ll, ll, i, vp, ll, vp, vp, vp, vp, i, vp]
    lib.emul_mass_force.argtypes = [vp, ll, ll, vp, vp, vp, vp, C.c_double, i]
    lib.emul_jt_ring_plan.argtypes = [vp, i, i, vp]
    lib.emul_applyJT_rows_cached.argtypes = [vp, ll, vp, ll, ll, vp, vp, vp, vp, vp, vp]
    lib.emul_frame_jacobians.argtypes = [vp, ll, vp, ll, ll, vp, vp, vp]
    lib.emul_getJ.argtypes = [vp, ll, vp, ll, i, vp, vp, vp, vp]
    lib.emul_aba.argtypes = [vp, ll, ll, vp, vp, vp, vp]
    lib.emul_aba2.argtypes = [vp, ll, ll, i, i, vp, vp, vp, vp, vp]
    lib.emul_minv.argtypes = [vp, ll, ll, i, i, vp, vp, vp]
    lib.emul_drnea.argtypes = [vp, ll, ll, i, i, vp, vp, vp, vp, vp, vp, vp, i]
    lib.emul_drnea_pattern.argtypes = [vp, vp, vp, vp]
    lib.emul_pick_wpb.argtypes = [i, ll, i]
    lib.emul_eval_arm.argtypes = [vp, ll, ll] + [vp] * 7
    lib.emul_eval_plan.argtypes = [vp, i, i, vp]
    lib.emul_applyDJT.argtypes = [vp, ll, ll, vp, ll, vp, vp, vp, C.c_double, vp, vp]
    lib.emul_applyDJT_ring.argtypes = [vp, ll, ll, vp, ll, vp, vp, vp, C.c_double, vp, vp]
    lib.emul_applyJT_rows.argtypes = [vp, ll, i, vp, ll, vp, vp, vp, vp, vp, vp]
    return lib


def mass_pattern_reference(d):
    """Support pattern of M's upper triangle straight from the definition (include/rbd_b200.h, rbd_model_mass_pattern):
    (r, c), r <= c, is kept iff the joint owning dof r is the joint owning dof c or one of its ancestors.  Returns
    (col_ptr, row_idx, packed index c (c + 1) / 2 + r of every kept entry), column by column, rows ascending."""
    owner = np.zeros(d.nv, dtype=np.int64)
    nvj = [0] * d.njoints
    for j in range(1, d.njoints):
        nxt = int(d.idx_v[j + 1]) if j + 1 < d.njoints else d.nv
        nvj[j] = nxt - int(d.idx_v[j])
        owner[int(d.idx_v[j]):nxt] = j
    anc = []
    for j in range(d.njoints):
        s, k = set(), j
        while k != 0:
            s.add(k); k = int(d.parents[k])
        anc.append(s)
    col_ptr, rows, packed = [0], [], []
    for c in range(d.nv):
        for r in range(c + 1):
            if owner[r] in anc[owner[c]]:
                rows.append(r); packed.append(c * (c + 1) // 2 + r)
        col_ptr.append(len(rows))
    return np.array(col_ptr, dtype=np.int32), np.array(rows, dtype=np.int32), np.array(packed, dtype=np.int64)


def P(a):
    return None if a is None else a.ctypes.data


def support_relation(d) -> np.ndarray:
    """[nv, nv] bool: dofs r and c belong to joints on one root-to-leaf path (one is the other's ancestor, or the same
    joint) — the structural non-zero pattern of dtau_dq / dtau_dv (and, upper triangle, of M)."""
    nj, nv = d.njoints, d.nv
    owner = np.zeros(nv, dtype=np.int64)
    for j in range(1, nj):
        nxt = int(d.idx_v[j + 1]) if j + 1 < nj else nv
        owner[int(d.idx_v[j]):nxt] = j
    anc = np.zeros((nj, nj), dtype=bool)                   # anc[a, i]: a is i or an ancestor of i
    for i in range(1, nj):
        k = i
        while k != 0:
            anc[k, i] = True; k = int(d.parents[k])
    rel = anc | anc.T
    return rel[np.ix_(owner, owner)]


def all_models():
    from sofa.rigidbodydynamics_b200 import model as mdl
    ms = [mdl.panda7(), mdl.solo12(), mdl.talos_reduced(),
          mdl.random_tree(1, 10), mdl.random_tree(2, 12, free_flyer=True),
          mdl.random_tree(3, 20, branch_prob=0.6), mdl.random_tree(4, 1), mdl.random_tree(5, 6, all_types=False)]
    for name in ("simple_arm", "schunk_svh_hand_right_regularized_inertia", "schunk_svh_hand_right_ff"):
        p = os.path.join(GOLDEN, name + ".model.json")
        if os.path.exists(p):
            with open(p) as f:
                ms.append(mdl.ModelDescriptor.from_json(f.read()))
    return ms


def subset_models():
    """BASELINE robots and the reference's hand with a PERMUTED STRICT SUBSET of the frames as mapped outputs: the
    reference maps every frame (URDFModelLoader.cpp:159-165, "TODO should be configurable"); rbd_model_desc.out_frames
    expresses any selection and order."""
    import dataclasses
    out = []
    for k, d in enumerate(m for m in all_models() if m.name in ("talos_reduced", "solo12", "panda7", "random_3",
                                                                  "schunk_svh_hand_right_ff")):
        rng = np.random.default_rng(1000 + k)
        keep = max(3, d.nframes // 3)
        sel = rng.permutation(d.nframes)[:keep].astype(np.int32)      # permuted, strict subset, universe frame allowed
        assert keep < d.nframes
        out.append(dataclasses.replace(d, name=d.name + "_subset", out_frames=np.ascontiguousarray(sel)))
    return out


def integrate(d, q: np.ndarray, dq: np.ndarray, eps: float) -> np.ndarray:
    """q (+) eps * dq with pinocchio's conventions ([n, nq], [n, nv]): free-flyer p += R (eps v_b), quat <- quat * exp(eps w_b)
    (joint-frame velocity); unbounded revolute joints rotate their (cos, sin) pair; the rest is q += eps dq."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    out = q.copy()
    for i in range(1, d.njoints):
        t, iq, iv = int(d.jtype[i]), int(d.idx_q[i]), int(d.idx_v[i])
        if t == mdl.J_FF:
            x, y, z, w = (q[:, iq + 3 + k] for k in range(4))
            R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                          np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                          np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
            out[:, iq:iq + 3] = q[:, iq:iq + 3] + eps * np.einsum("nij,nj->ni", R, dq[:, iv:iv + 3])
            th = eps * dq[:, iv + 3:iv + 6]
            ang = np.linalg.norm(th, axis=1)
            s = np.where(ang > 1e-300, np.sin(ang / 2) / np.where(ang > 1e-300, ang, 1.0), 0.5)
            dx, dy, dz, dw = s * th[:, 0], s * th[:, 1], s * th[:, 2], np.cos(ang / 2)
            out[:, iq + 3] = w * dx + x * dw + y * dz - z * dy
            out[:, iq + 4] = w * dy - x * dz + y * dw + z * dx
            out[:, iq + 5] = w * dz + x * dy - y * dx + z * dw
            out[:, iq + 6] = w * dw - x * dx - y * dy - z * dz
        elif mdl.J_RUBX <= t <= mdl.J_RUBU:
            c, s_ = q[:, iq], q[:, iq + 1]
            a = eps * dq[:, iv]
            out[:, iq] = c * np.cos(a) - s_ * np.sin(a)
            out[:, iq + 1] = s_ * np.cos(a) + c * np.sin(a)
        else:
            out[:, iq] = q[:, iq] + eps * dq[:, iv]
    return out


def fd_world_jacobian_error(eval_fn, d, q: np.ndarray, eps: float = 1e-6) -> float:
    """Central finite differences of the joint placements along every dof (q (+) eps e_c, pinocchio's integrate: the
    free-flyer moves by a velocity given in ITS OWN frame) against the world joint Jacobian: column c of data.J is the
    world-frame spatial velocity [v_O; w] (at the world origin) that a unit dq_c gives every body below joint(c), and
    leaves the other bodies at rest.  eval_fn(q [N, nq]) -> (oMi [N, 12 (nj - 1)], J [N, 6 nv]).  Returns the largest
    deviation; pins the free-flyer velocity frame and the [lin; ang] ordering independently of any second implementation
    of the Jacobian."""
    n, nj, nv = q.shape[0], d.njoints, d.nv
    owner = np.zeros(nv, dtype=np.int64)
    for j in range(1, nj):
        nxt = int(d.idx_v[j + 1]) if j + 1 < nj else nv
        owner[int(d.idx_v[j]):nxt] = j
    below = np.zeros((nj, nj), dtype=bool)                 # below[a, i]: joint i is a or a descendant of a
    for i in range(1, nj):
        k = i
        while k != 0:
            below[k, i] = True; k = int(d.parents[k])
    qs = [q]
    for c in range(nv):
        e = np.zeros((n, nv)); e[:, c] = 1.0
        qs.append(integrate(d, q, e, eps)); qs.append(integrate(d, q, e, -eps))
    oMi, J = eval_fn(np.ascontiguousarray(np.concatenate(qs)))
    oMi = oMi.reshape(2 * nv + 1, n, nj - 1, 12); J0 = J[:n].reshape(n, nv, 6)
    R0 = oMi[0][..., :9].reshape(n, nj - 1, 3, 3); p0 = oMi[0][..., 9:]
    worst = 0.0
    for c in range(nv):
        P, M = oMi[1 + 2 * c], oMi[2 + 2 * c]
        dR = (P[..., :9] - M[..., :9]).reshape(n, nj - 1, 3, 3) / (2 * eps)
        dp = (P[..., 9:] - M[..., 9:]) / (2 * eps)
        W = np.einsum("nkij,nklj->nkil", dR, R0)           # dR R^T = skew(w)
        w = np.stack([W[..., 2, 1], W[..., 0, 2], W[..., 1, 0]], -1)
        v = dp - np.cross(w, p0)                           # velocity of the point at the world origin
        fd = np.concatenate([v, w], -1)                    # [n, nj - 1, 6]
        for i in range(1, nj):
            ref = J0[:, c] if below[owner[c], i] else np.zeros((n, 6))
            worst = max(worst, float(np.max(np.abs(fd[:, i - 1] - ref))))
    return worst

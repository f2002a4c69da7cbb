"""CPU oracle loader — TEST INFRASTRUCTURE ONLY (see ``oracle/rbd_oracle.c`` header; parity unpinned).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  Nothing under ``sofa/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbd_oracle.so")


def _host_id() -> str:
    """The oracle is compiled with -march=native: rebuild when the snapshot lands on a different host CPU."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rbd_oracle.c")
    stamp = os.path.join(_HERE, "librbd_oracle.hostid")
    hid = _host_id()
    try:
        with open(stamp) as f:
            same_host = f.read().strip() == hid
    except OSError:
        same_host = False
    if force or not same_host or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librbd_oracle.so", "librbd_oracle_flops.so"], stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(hid)
    return LIB_PATH


FLOPS_LIB_PATH = os.path.join(_HERE, "librbd_oracle_flops.so")
_flops_lib = None


def flop_counts(desc, q=None, v=None, a=None) -> dict:
    """Exact FP64 operation counts of ONE evaluation (computeJointJacobians + crba + rnea, all outputs) of the restated
    reference algorithm, from the ORC_COUNT_FLOPS build (oracle/Makefile: the same source compiled as C++ with a
    counting scalar).  ``flops`` = add + mul + div + sqrt (one each); ``sincos`` counts sine and cosine evaluations
    separately.  The counts depend on the model only (the code path has no data-dependent arithmetic branch)."""
    global _flops_lib
    if _flops_lib is None:
        src = os.path.join(_HERE, "rbd_oracle.c")
        if not os.path.exists(FLOPS_LIB_PATH) or os.path.getmtime(FLOPS_LIB_PATH) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", "librbd_oracle_flops.so"], stdout=subprocess.DEVNULL)
        _flops_lib = C.CDLL(FLOPS_LIB_PATH)
        _flops_lib.orc_create.restype = C.c_void_p; _flops_lib.orc_create.argtypes = [C.c_void_p]
        _flops_lib.orc_destroy.argtypes = [C.c_void_p]
        _flops_lib.orc_eval.argtypes = [C.c_void_p] * 8
        _flops_lib.orc_flop_counts.argtypes = [C.c_void_p]
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200._capi import make_desc
    if q is None:
        q, v, a = (x[0] for x in mdl.random_state(desc, 1, seed=1))
    q, v, a = _f64(q), _f64(v), _f64(a)
    cdesc, _keep = make_desc(desc)
    L = _flops_lib
    w = L.orc_create(C.addressof(cdesc))
    oMi = np.zeros(max(1, 12 * (desc.njoints - 1))); J = np.zeros(max(1, 6 * desc.nv))
    M = np.zeros(max(1, desc.nv * (desc.nv + 1) // 2)); tau = np.zeros(max(1, desc.nv))
    out = {}
    parts = {"fk_jacobian": (_p(oMi), _p(J), None, None), "crba": (None, None, _p(M), None),
             "rnea": (None, None, None, _p(tau)), "eval": (_p(oMi), _p(J), _p(M), _p(tau))}
    for name, (po, pj, pm, pt) in parts.items():
        L.orc_flop_reset()
        L.orc_eval(w, _p(q), _p(v), _p(a), po, pj, pm, pt)
        c = (C.c_longlong * 6)()
        L.orc_flop_counts(c)
        out[name] = {"add": int(c[0]), "mul": int(c[1]), "div": int(c[2]), "sqrt": int(c[3]), "sincos": int(c[4]),
                     "flops": int(c[0] + c[1] + c[2] + c[3])}
    L.orc_destroy(w)
    return out


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        vp, dp = C.c_void_p, C.c_void_p
        l.orc_create.restype = vp; l.orc_create.argtypes = [vp]
        l.orc_destroy.restype = None; l.orc_destroy.argtypes = [vp]
        l.orc_compute_joint_jacobians.restype = None; l.orc_compute_joint_jacobians.argtypes = [vp, dp]
        l.orc_update_frame_placements.restype = None; l.orc_update_frame_placements.argtypes = [vp]
        l.orc_get_frame_jacobian_lwa.restype = None; l.orc_get_frame_jacobian_lwa.argtypes = [vp, C.c_int, dp]
        l.orc_rot_to_quat.restype = None; l.orc_rot_to_quat.argtypes = [dp, dp]
        l.orc_apply.restype = None; l.orc_apply.argtypes = [vp, dp, dp, dp]
        l.orc_apply_loop.restype = None; l.orc_apply_loop.argtypes = [vp, C.c_long, dp, dp, dp]
        l.orc_applyJ.restype = None; l.orc_applyJ.argtypes = [vp, dp, dp, dp]
        l.orc_applyJT.restype = None; l.orc_applyJT.argtypes = [vp, dp, dp, dp]
        l.orc_applyJT_row.restype = C.c_int; l.orc_applyJT_row.argtypes = [vp, C.c_int, vp, dp, dp]
        l.orc_crba.restype = None; l.orc_crba.argtypes = [vp, dp, dp]
        l.orc_rnea.restype = None; l.orc_rnea.argtypes = [vp, dp, dp, dp, dp]
        l.orc_aba.restype = None; l.orc_aba.argtypes = [vp, dp, dp, dp, dp]
        l.orc_rnea_derivatives.restype = None; l.orc_rnea_derivatives.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        l.orc_eval.restype = None; l.orc_eval.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp]
        l.orc_eval_batch.restype = C.c_int
        l.orc_eval_batch.argtypes = [vp, C.c_long, dp, dp, dp, dp, dp, dp, dp, C.c_int]
        l.orc_data_J.restype = C.POINTER(C.c_double); l.orc_data_J.argtypes = [vp]
        l.orc_get_oMi.restype = None; l.orc_get_oMi.argtypes = [vp, C.c_int, dp]
        l.orc_get_oMf.restype = None; l.orc_get_oMf.argtypes = [vp, C.c_int, dp]
        l.orc_max_threads.restype = C.c_int; l.orc_max_threads.argtypes = []
        _lib = l
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Oracle:
    """Single-instance pinocchio-semantics oracle bound to one model descriptor."""

    def __init__(self, desc):
        from sofa.rigidbodydynamics_b200._capi import make_desc  # POD marshalling only
        self.d = desc
        self._cdesc, self._keep = make_desc(desc)
        self._w = lib().orc_create(C.addressof(self._cdesc))

    def __del__(self):
        try:
            if getattr(self, "_w", None):
                lib().orc_destroy(self._w); self._w = None
        except Exception:
            pass

    # -- mapping path ------------------------------------------------------------------------------------
    def apply(self, q_joints, root_pose=None):
        q_joints = _f64(q_joints); root_pose = None if root_pose is None else _f64(root_pose)
        out = np.zeros((self.d.n_out, 7))
        lib().orc_apply(self._w, _p(q_joints), _p(root_pose), _p(out))
        return out

    def applyJ(self, dq_joints, root_vel=None):
        dq_joints = _f64(dq_joints); root_vel = None if root_vel is None else _f64(root_vel)
        out = np.zeros((self.d.n_out, 6))
        lib().orc_applyJ(self._w, _p(dq_joints), _p(root_vel), _p(out))
        return out

    def applyJT(self, wrenches, out_tau=None, out_root=None):
        """Returns (out_tau, out_root) after ``+=`` accumulation into the given arrays (zeros if omitted)."""
        w = _f64(wrenches)
        tau = np.zeros(self.d.nv_joints) if out_tau is None else _f64(out_tau).copy()
        root = None
        if self.d.has_freeflyer_root:
            root = np.zeros(6) if out_root is None else _f64(out_root).copy()
        lib().orc_applyJT(self._w, _p(w), _p(tau), _p(root))
        return tau, root

    def applyJT_row(self, cols, vals):
        cols = np.ascontiguousarray(np.asarray(cols, dtype=np.int32)); vals = _f64(vals)
        row = np.zeros(self.d.nv)
        keep = lib().orc_applyJT_row(self._w, int(cols.shape[0]), cols.ctypes.data, _p(vals), _p(row))
        return row, bool(keep)

    def frame_jacobian_lwa(self, frame: int):
        J = np.zeros((self.d.nv, 6))  # col-major 6 x nv == row-major nv x 6
        lib().orc_get_frame_jacobian_lwa(self._w, int(frame), _p(J))
        return J.T.copy()

    # -- dynamics ----------------------------------------------------------------------------------------
    def crba(self, q):
        q = _f64(q); M = np.zeros((self.d.nv, self.d.nv))
        lib().orc_crba(self._w, _p(q), _p(M))
        return M

    def rnea(self, q, v, a):
        q, v, a = _f64(q), _f64(v), _f64(a); tau = np.zeros(self.d.nv)
        lib().orc_rnea(self._w, _p(q), _p(v), _p(a), _p(tau))
        return tau

    def aba(self, q, v, tau):
        """pinocchio::aba(q, v, tau) (LOCAL convention): joint accelerations."""
        q, v, tau = _f64(q), _f64(v), _f64(tau); ddq = np.zeros(self.d.nv)
        lib().orc_aba(self._w, _p(q), _p(v), _p(tau), _p(ddq))
        return ddq

    def rnea_derivatives(self, q, v, a):
        """pinocchio::computeRNEADerivatives(q, v, a): (dtau_dq, dtau_dv, M) — nv x nv each, entry (r, c) = d tau_r / d x_c
        (q: tangent of integrate); M = dtau_da, upper triangle filled."""
        q, v, a = _f64(q), _f64(v), _f64(a); nv = self.d.nv
        dq = np.zeros((nv, nv)); dv = np.zeros((nv, nv)); M = np.zeros((nv, nv))
        lib().orc_rnea_derivatives(self._w, _p(q), _p(v), _p(a), _p(dq), _p(dv), _p(M))
        return dq, dv, M

    def minv(self, q):
        """pinocchio::computeMinverse(q) restated as what it equals: the inverse of the CRBA mass matrix (dense, symmetric;
        numpy LU).  tests/test_oracle.py ties it to the articulated-body restatement: column j == aba(q, 0, e_j) - aba(q, 0, 0)."""
        M = self.crba(q)
        M = np.triu(M) + np.triu(M, 1).T
        return np.linalg.inv(M)

    def eval(self, q, v, a):
        q, v, a = _f64(q), _f64(v), _f64(a)
        d = self.d
        oMi = np.zeros(12 * (d.njoints - 1)); J = np.zeros(6 * d.nv)
        M = np.zeros(d.nv * (d.nv + 1) // 2); tau = np.zeros(d.nv)
        lib().orc_eval(self._w, _p(q), _p(v), _p(a), _p(oMi), _p(J), _p(M), _p(tau))
        return oMi, J, M, tau


def eval_batch(desc, q, v, a, nthreads: int = 0, want=("oMi", "J", "M", "tau")):
    """Instance-major batched eval on ``nthreads`` host threads.  Returns (dict of outputs, threads used)."""
    from sofa.rigidbodydynamics_b200._capi import make_desc
    cdesc, _keep = make_desc(desc)
    q, v, a = _f64(q), _f64(v), _f64(a)
    n = q.shape[0]
    out = {"oMi": np.zeros((n, 12 * (desc.njoints - 1))) if "oMi" in want else None,
           "J": np.zeros((n, 6 * desc.nv)) if "J" in want else None,
           "M": np.zeros((n, desc.nv * (desc.nv + 1) // 2)) if "M" in want else None,
           "tau": np.zeros((n, desc.nv)) if "tau" in want else None}
    used = lib().orc_eval_batch(C.addressof(cdesc), n, _p(q), _p(v), _p(a), _p(out["oMi"]), _p(out["J"]),
                                _p(out["M"]), _p(out["tau"]), int(nthreads))
    return out, used


def rot_to_quat(R):
    R = _f64(R).reshape(9); out = np.zeros(4)
    lib().orc_rot_to_quat(_p(R), _p(out))
    return out


def unpack_upper(Mp: np.ndarray, nv: int) -> np.ndarray:
    """Packed-by-columns upper triangle -> full symmetric matrix."""
    M = np.zeros((nv, nv))
    for c in range(nv):
        for r in range(c + 1):
            M[r, c] = M[c, r] = Mp[c * (c + 1) // 2 + r]
    return M

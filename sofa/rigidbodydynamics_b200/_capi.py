"""ctypes binding of the C ABI declared in ``include/rbd_b200.h`` (the drop-in boundary).

The shared library is built in-tree by ``build.py`` / ``__graft_entry__.build()`` as
``sofa/rigidbodydynamics_b200/librbd_b200.so``.  There is no fallback: if the library is missing,
:func:`load` raises, and every compute entry point fails with ``RBD_ERR_CUDA`` when no GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .model import ModelDescriptor

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBD_B200_LIB: alternate build of the same library (A/B timing of kernel variants); never a different backend
LIB_PATH = os.environ.get("RBD_B200_LIB") or os.path.join(_HERE, "librbd_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class RbdModelDesc(C.Structure):
    """``rbd_model_desc`` (include/rbd_b200.h)."""
    _fields_ = [("njoints", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32),
                ("parents", c_int32_p), ("jtype", c_int32_p), ("axis", c_double_p), ("placement", c_double_p),
                ("idx_q", c_int32_p), ("idx_v", c_int32_p), ("inertia", c_double_p),
                ("nframes", C.c_int32), ("frame_parent", c_int32_p), ("frame_placement", c_double_p),
                ("gravity", C.c_double * 3), ("n_out", C.c_int32), ("out_frames", c_int32_p)]


class RbdModelInfo(C.Structure):
    """``rbd_model_info`` (include/rbd_b200.h)."""
    _fields_ = [("njoints", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32), ("nframes", C.c_int32),
                ("n_out", C.c_int32), ("has_freeflyer_root", C.c_int32), ("nq_joints", C.c_int32),
                ("nv_joints", C.c_int32), ("omi_fields", C.c_int32), ("jac_fields", C.c_int32),
                ("mass_fields", C.c_int32), ("max_depth", C.c_int32), ("eval_smem_doubles", C.c_int32),
                ("eval_bytes", C.c_int64), ("eval_tmem_doubles", C.c_int32), ("eval_warps_per_sm", C.c_int32),
                ("mass_support_fields", C.c_int32)]


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


def make_desc(d: ModelDescriptor):
    """Marshal a :class:`ModelDescriptor` into a ``rbd_model_desc``.  Returns (struct, keepalive)."""
    d.validate()
    s = RbdModelDesc()
    s.njoints, s.nq, s.nv = d.njoints, d.nq, d.nv
    s.parents, s.jtype = _ip(d.parents), _ip(d.jtype)
    s.axis, s.placement = _dp(d.axis), _dp(d.placement)
    s.idx_q, s.idx_v = _ip(d.idx_q), _ip(d.idx_v)
    s.inertia = _dp(d.inertia)
    s.nframes = d.nframes
    s.frame_parent, s.frame_placement = _ip(d.frame_parent), _dp(d.frame_placement)
    for k in range(3):
        s.gravity[k] = float(d.gravity[k])
    s.n_out = d.n_out
    s.out_frames = _ip(d.out_frames)
    return s, d


# every symbol include/rbd_b200.h declares; tests check the built library exports all of them
SYMBOLS = [
    "rbd_version", "rbd_last_error", "rbd_clear_error", "rbd_device_count", "rbd_last_kernel_name", "rbd_probe_fp64_peak",
    "rbd_model_create", "rbd_model_destroy", "rbd_model_get_info", "rbd_model_mass_pattern",
    "rbd_eval_soa_sparse", "rbd_crba_soa_sparse", "rbd_eval_host_sparse",
    "rbd_model_getJ_pattern", "rbd_getJ_dev", "rbd_getJ_host", "rbd_aba_soa", "rbd_aba_host", "rbd_minv_soa", "rbd_minv_host",
    "rbd_rnea_derivatives_soa", "rbd_rnea_derivatives_host",
    "rbd_model_drnea_pattern", "rbd_rnea_derivatives_soa_sparse", "rbd_rnea_derivatives_host_sparse",
    "rbd_workspace_create", "rbd_workspace_destroy", "rbd_workspace_set_stream", "rbd_workspace_synchronize",
    "rbd_workspace_launch_count", "rbd_workspace_set_eval_block", "rbd_workspace_set_eval_tmem", "rbd_workspace_set_applyJT_ring", "rbd_workspace_set_rows_cache", "rbd_workspace_set_aba_fused", "rbd_workspace_set_eval_arm", "rbd_workspace_get_eval_plan", "rbd_workspace_get_eval_plan_f32",
    "rbd_apply_soa", "rbd_applyJ_soa", "rbd_applyJT_soa", "rbd_applyJT_rows_dev",
    "rbd_getFrameJacobian_dev", "rbd_getFrameJacobian_host", "rbd_applyDJT_soa", "rbd_applyDJT_host",
    "rbd_eval_soa", "rbd_eval_soa_f32", "rbd_crba_soa", "rbd_rnea_soa", "rbd_addMDx_soa", "rbd_addForce_soa",
    "rbd_aos_to_soa_dev", "rbd_soa_to_aos_dev",
    "rbd_apply_host", "rbd_applyJ_host", "rbd_applyJT_host", "rbd_applyJT_rows_host", "rbd_eval_host", "rbd_addMDx_host", "rbd_addForce_host",
    "rbd_batch_create", "rbd_batch_destroy", "rbd_batch_bind_apply", "rbd_batch_bind_applyJ", "rbd_batch_bind_applyJT",
    "rbd_batch_apply", "rbd_batch_applyJ", "rbd_batch_applyJT",
    "rbd_urdf_load_file", "rbd_urdf_load_string", "rbd_urdf_destroy", "rbd_urdf_desc", "rbd_urdf_nframes0",
    "rbd_urdf_robot_name", "rbd_urdf_joint_name", "rbd_urdf_frame_name", "rbd_urdf_frame_type", "rbd_urdf_find_frame",
    "rbd_urdf_find_joint",
]

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load ``librbd_b200.so``; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sofa.rigidbodydynamics_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    sig = {
        "rbd_version": (C.c_int, []),
        "rbd_last_error": (C.c_char_p, []),
        "rbd_clear_error": (None, []),
        "rbd_device_count": (C.c_int, [C.POINTER(C.c_int)]),
        "rbd_last_kernel_name": (C.c_char_p, []),
        "rbd_probe_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
        "rbd_model_mass_pattern": (C.c_int, [vp, vp, vp]),
        "rbd_model_getJ_pattern": (C.c_int, [vp, i32, vp, vp, vp]),
        "rbd_getJ_dev": (C.c_int, [vp, i64, i32, vp, vp]),
        "rbd_getJ_host": (C.c_int, [vp, i64, i32, vp, vp]),
        "rbd_aba_soa": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
        "rbd_aba_host": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "rbd_minv_soa": (C.c_int, [vp, i64, i64, vp, vp]),
        "rbd_minv_host": (C.c_int, [vp, i64, vp, vp]),
        "rbd_rnea_derivatives_soa": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp, vp]),
        "rbd_rnea_derivatives_host": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
        "rbd_model_drnea_pattern": (C.c_int, [vp, vp, vp, vp]),
        "rbd_rnea_derivatives_soa_sparse": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp]),
        "rbd_rnea_derivatives_host_sparse": (C.c_int, [vp, i64, vp, vp, vp, vp, vp]),
        "rbd_eval_soa_sparse": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rbd_crba_soa_sparse": (C.c_int, [vp, i64, i64, vp, vp]),
        "rbd_eval_host_sparse": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rbd_model_create": (C.c_int, [C.POINTER(RbdModelDesc), C.POINTER(vp)]),
        "rbd_model_destroy": (None, [vp]),
        "rbd_model_get_info": (C.c_int, [vp, C.POINTER(RbdModelInfo)]),
        "rbd_workspace_create": (C.c_int, [vp, C.c_int, i64, vp, C.POINTER(vp)]),
        "rbd_workspace_destroy": (None, [vp]),
        "rbd_workspace_set_stream": (C.c_int, [vp, vp]),
        "rbd_workspace_synchronize": (C.c_int, [vp]),
        "rbd_workspace_launch_count": (i64, [vp]),
        "rbd_workspace_set_eval_block": (C.c_int, [vp, C.c_int]),
        "rbd_workspace_set_eval_tmem": (C.c_int, [vp, C.c_int]),
        "rbd_workspace_set_applyJT_ring": (C.c_int, [vp, C.c_int, vp]),
        "rbd_workspace_set_rows_cache": (C.c_int, [vp, C.c_int]),
        "rbd_workspace_set_aba_fused": (C.c_int, [vp, C.c_int]),
        "rbd_workspace_set_eval_arm": (C.c_int, [vp, C.c_int]),
        "rbd_workspace_get_eval_plan": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_int32 * 8)]),
        "rbd_workspace_get_eval_plan_f32": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_int32 * 8)]),
        "rbd_apply_soa": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "rbd_applyJ_soa": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "rbd_applyJT_soa": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "rbd_applyJT_rows_dev": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
        "rbd_eval_soa": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rbd_eval_soa_f32": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rbd_crba_soa": (C.c_int, [vp, i64, i64, vp, vp]),
        "rbd_rnea_soa": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
        "rbd_getFrameJacobian_dev": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_applyDJT_soa": (C.c_int, [vp, i64, i64, vp, vp, vp, C.c_double, vp, vp]),
        "rbd_applyDJT_host": (C.c_int, [vp, i64, vp, vp, vp, C.c_double, vp, vp]),
        "rbd_getFrameJacobian_host": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_addMDx_soa": (C.c_int, [vp, i64, i64, vp, vp, C.c_double, vp]),
        "rbd_addForce_soa": (C.c_int, [vp, i64, i64, vp, vp, vp]),
        "rbd_addMDx_host": (C.c_int, [vp, i64, vp, vp, C.c_double, vp]),
        "rbd_addForce_host": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_aos_to_soa_dev": (C.c_int, [vp, i64, i32, i64, vp, vp]),
        "rbd_soa_to_aos_dev": (C.c_int, [vp, i64, i32, i64, vp, vp]),
        "rbd_apply_host": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_applyJ_host": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_applyJT_host": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_applyJT_rows_host": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
        "rbd_eval_host": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rbd_batch_create": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "rbd_batch_destroy": (None, [vp]),
        "rbd_batch_bind_apply": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_batch_bind_applyJ": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_batch_bind_applyJT": (C.c_int, [vp, i64, vp, vp, vp]),
        "rbd_batch_apply": (C.c_int, [vp, i64, i64]),
        "rbd_batch_applyJ": (C.c_int, [vp, i64, i64]),
        "rbd_batch_applyJT": (C.c_int, [vp, i64, i64]),
        "rbd_urdf_load_file": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "rbd_urdf_load_string": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "rbd_urdf_destroy": (None, [vp]),
        "rbd_urdf_desc": (C.POINTER(RbdModelDesc), [vp]),
        "rbd_urdf_nframes0": (i32, [vp]),
        "rbd_urdf_robot_name": (C.c_char_p, [vp]),
        "rbd_urdf_joint_name": (C.c_char_p, [vp, i32]),
        "rbd_urdf_frame_name": (C.c_char_p, [vp, i32]),
        "rbd_urdf_frame_type": (i32, [vp, i32]),
        "rbd_urdf_find_frame": (i32, [vp, C.c_char_p, i32]),
        "rbd_urdf_find_joint": (i32, [vp, C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)        # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RbdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rbd_b200 error {code}: {msg}")
        self.code = code


def check(code: int) -> None:
    if code != 0:
        msg = load().rbd_last_error()
        raise RbdError(code, msg.decode() if msg else "")

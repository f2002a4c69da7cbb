"""Host-side handles over the C ABI (``include/rbd_b200.h``).

``Model`` / ``Workspace`` are thin RAII wrappers of ``rbd_model*`` / ``rbd_workspace*``.  Device entry points take
anything exposing ``data_ptr()`` (torch CUDA tensors: PyTorch is only the allocator/stream provider here) or a raw
integer device address; host entry points take C-contiguous float64 numpy arrays in the reference's own
instance-major layout.  Every call goes through the shared library; there is no Python/torch arithmetic path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from ._capi import RbdError, check
from .model import ModelDescriptor


def _dev_ptr(x) -> Optional[int]:
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        if hasattr(x, "is_cuda") and not x.is_cuda:
            raise ValueError("device entry point got a CPU tensor")
        if hasattr(x, "is_contiguous") and not x.is_contiguous():
            raise ValueError("device entry point needs contiguous tensors")
        return int(x.data_ptr())
    raise TypeError(f"cannot take a device pointer from {type(x)}")


def _host_ptr(x: Optional[np.ndarray], dtype=np.float64) -> Optional[int]:
    if x is None:
        return None
    if hasattr(x, "data_ptr") and not isinstance(x, np.ndarray):   # pinned torch CPU tensor
        if getattr(x, "is_cuda", False):
            raise ValueError("host entry point got a CUDA tensor")
        return int(x.data_ptr())
    if not isinstance(x, np.ndarray) or x.dtype != dtype or not x.flags.c_contiguous:
        raise ValueError(f"host entry point needs a C-contiguous {np.dtype(dtype).name} numpy array")
    return int(x.ctypes.data)


class Model:
    """``rbd_model``: immutable tree constants (replaces ``KinematicChainMapping::setModel`` + frame tables,
    reference KinematicChainMapping.inl:304-317)."""

    def __init__(self, desc: ModelDescriptor):
        self.desc = desc
        self._lib = _capi.load()
        cdesc, self._keep = _capi.make_desc(desc)
        h = C.c_void_p()
        check(self._lib.rbd_model_create(C.byref(cdesc), C.byref(h)))
        self._h = h
        info = _capi.RbdModelInfo()
        check(self._lib.rbd_model_get_info(self._h, C.byref(info)))
        self.info = info

    def drnea_pattern(self):
        """(row_ptr [nv + 1], col_idx [nnz]) of the structurally non-zero entries of dtau_dq / dtau_dv, compressed by rows:
        entry ``e`` is field ``e`` of the value arrays of ``rnea_derivatives_*_sparse`` (``rbd_model_drnea_pattern``)."""
        row_ptr = np.zeros(self.info.nv + 1, dtype=np.int32)
        nnz = C.c_int32(0)
        check(self._lib.rbd_model_drnea_pattern(self._h, row_ptr.ctypes.data, None, C.addressof(nnz)))
        col_idx = np.zeros(max(1, nnz.value), dtype=np.int32)
        check(self._lib.rbd_model_drnea_pattern(self._h, row_ptr.ctypes.data, col_idx.ctypes.data, None))
        return row_ptr, col_idx[:nnz.value]

    def mass_pattern(self):
        """Support pattern of M's upper triangle by columns: ``(col_ptr [nv + 1], row_idx [nnz])`` int32 arrays; entry
        ``e`` is field ``e`` of the ``Mval`` arrays of the ``*_sparse`` entry points (``rbd_model_mass_pattern``)."""
        col_ptr = np.zeros(self.info.nv + 1, dtype=np.int32)
        row_idx = np.zeros(max(1, self.info.mass_support_fields), dtype=np.int32)
        check(self._lib.rbd_model_mass_pattern(self._h, col_ptr.ctypes.data, row_idx.ctypes.data))
        return col_ptr, row_idx[:self.info.mass_support_fields]

    def getJ_pattern(self, sel_out=None):
        """Pattern of the assembled mapping Jacobian for the selected mapped outputs (None: all): ``(blk_ptr [n_sel + 1],
        cols [nnz])`` — block row ``s`` has the dof columns ``cols[blk_ptr[s]:blk_ptr[s+1]]`` (``rbd_model_getJ_pattern``)."""
        sel = None if sel_out is None else np.ascontiguousarray(sel_out, dtype=np.int32)
        n_sel = self.info.n_out if sel is None else int(sel.shape[0])
        blk = np.zeros(n_sel + 1, dtype=np.int32)
        sp = None if sel is None else sel.ctypes.data
        check(self._lib.rbd_model_getJ_pattern(self._h, n_sel, sp, blk.ctypes.data, None))
        cols = np.zeros(max(1, int(blk[-1])), dtype=np.int32)
        check(self._lib.rbd_model_getJ_pattern(self._h, n_sel, sp, blk.ctypes.data, cols.ctypes.data))
        return blk, cols[:int(blk[-1])]

    def mass_packed_index(self) -> np.ndarray:
        """Index of every support entry in the packed upper triangle (``M(r, c)`` at ``c (c + 1) / 2 + r``)."""
        col_ptr, row_idx = self.mass_pattern()
        cols = np.repeat(np.arange(self.info.nv), np.diff(col_ptr))
        return (cols * (cols + 1) // 2 + row_idx).astype(np.int64)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rbd_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Workspace:
    """``rbd_workspace``: per-device buffers + cached configuration (replaces the mapping's ``pinocchio::Data``,
    reference KinematicChainMapping.h:121)."""

    def __init__(self, model: Model, capacity: int, device: int = 0, stream: Optional[int] = None):
        self.model = model
        self._lib = model._lib
        h = C.c_void_p()
        check(self._lib.rbd_workspace_create(model._h, int(device), int(capacity), C.c_void_p(stream or 0), C.byref(h)))
        self._h = h
        self.capacity = int(capacity)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rbd_workspace_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ------------------------------------------------------------------------------------------
    def set_stream(self, stream: Optional[int]):
        check(self._lib.rbd_workspace_set_stream(self._h, C.c_void_p(stream or 0)))

    def synchronize(self):
        check(self._lib.rbd_workspace_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.rbd_workspace_launch_count(self._h))

    def set_eval_block(self, threads: int):
        check(self._lib.rbd_workspace_set_eval_block(self._h, int(threads)))

    def set_eval_tmem(self, enable: bool):
        check(self._lib.rbd_workspace_set_eval_tmem(self._h, 1 if enable else 0))

    def set_eval_arm(self, enable: bool):
        """A/B knob of the fused eval of 6- / 7-joint serial chains: unrolled arm kernel (default) or the generic chain sweep."""
        check(self._lib.rbd_workspace_set_eval_arm(self._h, 1 if enable else 0))

    def set_aba_fused(self, enable: bool):
        """A/B knob of forward dynamics: the fused sweep on on-chip level records (default) or the first version."""
        check(self._lib.rbd_workspace_set_aba_fused(self._h, 1 if enable else 0))

    def set_rows_cache(self, enable: bool):
        """A/B knob of the constraint-row applyJT: cached Jacobians (default) or one sweep per row."""
        check(self._lib.rbd_workspace_set_rows_cache(self._h, 1 if enable else 0))

    def set_applyJT_ring(self, enable: bool) -> int:
        """A/B knob of the vector applyJT: ring kernel (default) or register path; returns the ring slots per warp."""
        ns = C.c_int32(0)
        check(self._lib.rbd_workspace_set_applyJT_ring(self._h, 1 if enable else 0, C.byref(ns)))
        return int(ns.value)

    def eval_plan(self, with_M: bool = True, with_tau: bool = True, f32: bool = False) -> dict:
        out = (C.c_int32 * 8)()
        fn = self._lib.rbd_workspace_get_eval_plan_f32 if f32 else self._lib.rbd_workspace_get_eval_plan
        check(fn(self._h, int(with_M), int(with_tau), C.byref(out)))
        keys = ("warps_per_block", "blocks_per_sm", "smem_doubles", "tmem_doubles", "tmem_cols", "y_in_tmem",
                "smem_bytes_per_block")
        return {k: int(out[i]) for i, k in enumerate(keys)}

    # -- device tiled-SoA entry points ----------------------------------------------------------------------
    def eval_soa(self, n, tile, q, v=None, a=None, oMi=None, J=None, M=None, tau=None):
        check(self._lib.rbd_eval_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a), _dev_ptr(oMi),
                                     _dev_ptr(J), _dev_ptr(M), _dev_ptr(tau)))

    def eval_soa_f32(self, n, tile, q, v=None, a=None, oMi=None, J=None, M=None, tau=None):
        """Optional FP32 mode (tolerance 1e-4): float32 device tensors in the same tiled-SoA layouts."""
        check(self._lib.rbd_eval_soa_f32(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a), _dev_ptr(oMi),
                                         _dev_ptr(J), _dev_ptr(M), _dev_ptr(tau)))

    def eval_soa_sparse(self, n, tile, q, v=None, a=None, oMi=None, J=None, Mval=None, tau=None):
        """The fused evaluation with M in the support-only format (``Model.mass_pattern``)."""
        check(self._lib.rbd_eval_soa_sparse(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a), _dev_ptr(oMi),
                                            _dev_ptr(J), _dev_ptr(Mval), _dev_ptr(tau)))

    def crba_soa_sparse(self, n, tile, q, Mval):
        check(self._lib.rbd_crba_soa_sparse(self._h, n, tile, _dev_ptr(q), _dev_ptr(Mval)))

    def crba_soa(self, n, tile, q, M):
        check(self._lib.rbd_crba_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(M)))

    def rnea_soa(self, n, tile, q, v, a, tau):
        check(self._lib.rbd_rnea_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a), _dev_ptr(tau)))

    def aba_soa(self, n, tile, q, v, tau, ddq):
        """Forward dynamics ddq = aba(q, v, tau) (``rbd_aba_soa``)."""
        check(self._lib.rbd_aba_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(tau), _dev_ptr(ddq)))

    def rnea_derivatives_soa(self, n, tile, q, v, a, dtau_dq, dtau_dv, M=None):
        """dtau_dq, dtau_dv [nv * nv fields, row-major] (and M packed upper, optional) of tau = rnea(q, v, a)
        (``rbd_rnea_derivatives_soa``, pinocchio::computeRNEADerivatives)."""
        check(self._lib.rbd_rnea_derivatives_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a), _dev_ptr(dtau_dq),
                                                 _dev_ptr(dtau_dv), _dev_ptr(M)))

    def rnea_derivatives_soa_sparse(self, n, tile, q, v, a, dq_val, dv_val):
        """The same derivatives, structurally non-zero entries only (``Model.drnea_pattern``)."""
        check(self._lib.rbd_rnea_derivatives_soa_sparse(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(a),
                                                        _dev_ptr(dq_val), _dev_ptr(dv_val)))

    def minv_soa(self, n, tile, q, Minv):
        """Packed upper triangle of M(q)^-1 (``rbd_minv_soa``, pinocchio::computeMinverse)."""
        check(self._lib.rbd_minv_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(Minv)))

    def addMDx_soa(self, n, tile, q, dx, factor, res):
        """res += factor * M(q) dx  (SOFA Mass::addMDx in joint space, matrix-free)."""
        check(self._lib.rbd_addMDx_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(dx), float(factor), _dev_ptr(res)))

    def addForce_soa(self, n, tile, q, v, f):
        """f -= rnea(q, v, 0)  (SOFA ForceField::addForce in joint space); v None: gravity only, f += sum_b J_b^T m_b g."""
        check(self._lib.rbd_addForce_soa(self._h, n, tile, _dev_ptr(q), _dev_ptr(v), _dev_ptr(f)))

    def apply_soa(self, n, tile, q_joints, root_pose, out):
        check(self._lib.rbd_apply_soa(self._h, n, tile, _dev_ptr(q_joints), _dev_ptr(root_pose), _dev_ptr(out)))

    def applyJ_soa(self, n, tile, dq_joints, root_vel, out):
        check(self._lib.rbd_applyJ_soa(self._h, n, tile, _dev_ptr(dq_joints), _dev_ptr(root_vel), _dev_ptr(out)))

    def applyJT_soa(self, n, tile, wrenches, out_tau, out_root=None):
        check(self._lib.rbd_applyJT_soa(self._h, n, tile, _dev_ptr(wrenches), _dev_ptr(out_tau), _dev_ptr(out_root)))

    def applyDJT_soa(self, n, tile, dq_joints, root_vel, wrenches, kfactor, out_tau, out_root=None):
        """Optional extension (the reference's applyDJT is a no-op): out += kfactor * d/de [J(q + e dq)^T w]."""
        check(self._lib.rbd_applyDJT_soa(self._h, n, tile, _dev_ptr(dq_joints), _dev_ptr(root_vel), _dev_ptr(wrenches),
                                         float(kfactor), _dev_ptr(out_tau), _dev_ptr(out_root)))

    def applyDJT_host(self, n, dq_joints, root_vel, wrenches, kfactor, out_tau, out_root=None):
        check(self._lib.rbd_applyDJT_host(self._h, n, _host_ptr(dq_joints), _host_ptr(root_vel), _host_ptr(wrenches),
                                          float(kfactor), _host_ptr(out_tau), _host_ptr(out_root)))

    def applyJT_rows_dev(self, nrows, row_ptr, row_instance, col, val, out_rows, keep):
        check(self._lib.rbd_applyJT_rows_dev(self._h, nrows, _dev_ptr(row_ptr), _dev_ptr(row_instance), _dev_ptr(col),
                                             _dev_ptr(val), _dev_ptr(out_rows), _dev_ptr(keep)))

    def frame_jacobians_dev(self, npairs, instance, out_index, J):
        """Dense LOCAL_WORLD_ALIGNED frame Jacobians of (instance, output) pairs: J [npairs, nv, 6] (device tensors)."""
        check(self._lib.rbd_getFrameJacobian_dev(self._h, npairs, _dev_ptr(instance), _dev_ptr(out_index), _dev_ptr(J)))

    def frame_jacobians_host(self, instance, out_index) -> np.ndarray:
        """... on host arrays; returns [npairs, 6, nv] (pinocchio's Matrix6x as numpy sees it)."""
        inst = np.ascontiguousarray(instance, dtype=np.int32); out = np.ascontiguousarray(out_index, dtype=np.int32)
        nv = self.model.info.nv
        J = np.zeros((inst.shape[0], nv, 6))
        check(self._lib.rbd_getFrameJacobian_host(self._h, inst.shape[0], _host_ptr(inst, np.int32),
                                                  _host_ptr(out, np.int32), _host_ptr(J)))
        return np.ascontiguousarray(J.transpose(0, 2, 1))

    def getJ_host(self, n, sel_out=None) -> np.ndarray:
        """Assembled mapping Jacobian values of instances 0..n-1 for the selected outputs: ``[n, nnz_cols, 6]``, columns in
        the order of ``Model.getJ_pattern`` (``rbd_getJ_host``)."""
        blk, cols = self.model.getJ_pattern(sel_out)
        sel = None if sel_out is None else np.ascontiguousarray(sel_out, dtype=np.int32)
        n_sel = self.model.info.n_out if sel is None else int(sel.shape[0])
        val = np.zeros((n, max(1, cols.shape[0]), 6))
        check(self._lib.rbd_getJ_host(self._h, n, n_sel, None if sel is None else sel.ctypes.data, val.ctypes.data))
        return val[:, :cols.shape[0]]

    def getJ_dev(self, n, sel_out, val):
        sel = None if sel_out is None else np.ascontiguousarray(sel_out, dtype=np.int32)
        n_sel = self.model.info.n_out if sel is None else int(sel.shape[0])
        check(self._lib.rbd_getJ_dev(self._h, n, n_sel, None if sel is None else sel.ctypes.data, _dev_ptr(val)))

    def aos_to_soa(self, n, K, tile, aos, soa):
        check(self._lib.rbd_aos_to_soa_dev(self._h, n, K, tile, _dev_ptr(aos), _dev_ptr(soa)))

    def soa_to_aos(self, n, K, tile, soa, aos):
        check(self._lib.rbd_soa_to_aos_dev(self._h, n, K, tile, _dev_ptr(soa), _dev_ptr(aos)))

    # -- host (SOFA instance-major) entry points --------------------------------------------------------------
    def eval_host(self, n, q, v=None, a=None, oMi=None, J=None, M=None, tau=None):
        check(self._lib.rbd_eval_host(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(a), _host_ptr(oMi),
                                      _host_ptr(J), _host_ptr(M), _host_ptr(tau)))

    def eval_host_sparse(self, n, q, v=None, a=None, oMi=None, J=None, Mval=None, tau=None):
        check(self._lib.rbd_eval_host_sparse(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(a), _host_ptr(oMi),
                                             _host_ptr(J), _host_ptr(Mval), _host_ptr(tau)))

    def aba_host(self, n, q, v, tau, ddq):
        check(self._lib.rbd_aba_host(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(tau), _host_ptr(ddq)))

    def rnea_derivatives_host(self, n, q, v, a, dtau_dq, dtau_dv, M=None):
        check(self._lib.rbd_rnea_derivatives_host(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(a), _host_ptr(dtau_dq),
                                                  _host_ptr(dtau_dv), _host_ptr(M)))

    def rnea_derivatives_host_sparse(self, n, q, v, a, dq_val, dv_val):
        check(self._lib.rbd_rnea_derivatives_host_sparse(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(a),
                                                         _host_ptr(dq_val), _host_ptr(dv_val)))

    def minv_host(self, n, q, Minv):
        check(self._lib.rbd_minv_host(self._h, n, _host_ptr(q), _host_ptr(Minv)))

    def addMDx_host(self, n, q, dx, factor, res):
        """res += factor * M(q) dx on host vectors (instance-major), in place."""
        check(self._lib.rbd_addMDx_host(self._h, n, _host_ptr(q), _host_ptr(dx), float(factor), _host_ptr(res)))

    def addForce_host(self, n, q, v, f):
        """f -= rnea(q, v, 0) on host vectors, in place; v None: gravity only."""
        check(self._lib.rbd_addForce_host(self._h, n, _host_ptr(q), _host_ptr(v), _host_ptr(f)))

    def apply_host(self, n, q_joints, root_pose, out):
        check(self._lib.rbd_apply_host(self._h, n, _host_ptr(q_joints), _host_ptr(root_pose), _host_ptr(out)))

    def applyJ_host(self, n, dq_joints, root_vel, out):
        check(self._lib.rbd_applyJ_host(self._h, n, _host_ptr(dq_joints), _host_ptr(root_vel), _host_ptr(out)))

    def applyJT_host(self, n, wrenches, out_tau, out_root=None):
        check(self._lib.rbd_applyJT_host(self._h, n, _host_ptr(wrenches), _host_ptr(out_tau), _host_ptr(out_root)))

    def applyJT_rows_host(self, nrows, row_ptr, row_instance, col, val, out_rows, keep):
        check(self._lib.rbd_applyJT_rows_host(self._h, nrows, _host_ptr(row_ptr, np.int32),
                                              _host_ptr(row_instance, np.int32), _host_ptr(col, np.int32),
                                              _host_ptr(val), _host_ptr(out_rows), _host_ptr(keep, np.int32)))


def last_kernel_name() -> str:
    """Demangled name of the kernel this thread launched last through the library (``rbd_last_kernel_name``)."""
    s = _capi.load().rbd_last_kernel_name()
    return s.decode() if s else ""


def probe_fp64_peak(device: int = 0) -> float:
    """Measured FP64 peak of `device` in TFLOP/s (``rbd_probe_fp64_peak``: DFMA chains on every SM)."""
    t = C.c_double(0.0)
    check(_capi.load().rbd_probe_fp64_peak(int(device), C.byref(t)))
    return float(t.value)


def load_urdf(source: str, free_flyer: bool = False, is_text: bool = False) -> ModelDescriptor:
    """URDF file (or text, ``is_text=True``) -> :class:`ModelDescriptor` through the library's own reader
    (``rbd_urdf_load_file`` / ``rbd_urdf_load_string``, csrc/rbd_urdf.cpp) — what a C++ caller without pinocchio uses
    in place of ``pinocchio::urdf::buildModel`` + the loader's frame tables (reference URDFModelLoader.cpp:126-183)."""
    lib = _capi.load()
    h = C.c_void_p()
    fn = lib.rbd_urdf_load_string if is_text else lib.rbd_urdf_load_file
    check(fn(source.encode(), int(bool(free_flyer)), C.byref(h)))
    try:
        cd = lib.rbd_urdf_desc(h).contents
        nj, nf, no = int(cd.njoints), int(cd.nframes), int(cd.n_out)

        def arr(ptr, n, dtype):
            return np.array(np.ctypeslib.as_array(ptr, shape=(n,)), dtype=dtype, copy=True) if n else np.zeros(0, dtype)

        d = ModelDescriptor(
            name=(lib.rbd_urdf_robot_name(h) or b"robot").decode(),
            parents=arr(cd.parents, nj, np.int32), jtype=arr(cd.jtype, nj, np.int32),
            axis=arr(cd.axis, 3 * nj, np.float64).reshape(nj, 3),
            placement=arr(cd.placement, 12 * nj, np.float64).reshape(nj, 12),
            idx_q=arr(cd.idx_q, nj, np.int32), idx_v=arr(cd.idx_v, nj, np.int32),
            inertia=arr(cd.inertia, 10 * nj, np.float64).reshape(nj, 10),
            frame_parent=arr(cd.frame_parent, nf, np.int32),
            frame_placement=arr(cd.frame_placement, 12 * nf, np.float64).reshape(nf, 12),
            gravity=np.array([cd.gravity[0], cd.gravity[1], cd.gravity[2]], dtype=np.float64),
            out_frames=arr(cd.out_frames, no, np.int32), nframes0=int(lib.rbd_urdf_nframes0(h)),
            joint_names=[lib.rbd_urdf_joint_name(h, j).decode() for j in range(nj)],
            frame_names=[lib.rbd_urdf_frame_name(h, f).decode() for f in range(nf)],
            frame_types=[int(lib.rbd_urdf_frame_type(h, f)) for f in range(nf)],
            note="from URDF through rbd_urdf_load (pinocchio 3.3.1 URDF conventions, restated; unpinned)")
    finally:
        lib.rbd_urdf_destroy(h)
    d.validate()
    return d


def device_count() -> int:
    lib = _capi.load()
    n = C.c_int(0)
    rc = lib.rbd_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0


__all__ = ["Model", "Workspace", "RbdError", "device_count", "last_kernel_name", "load_urdf"]

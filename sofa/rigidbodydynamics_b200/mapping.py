"""Host-side mirror of the reference's ``KinematicChainMapping`` for N robot instances at once.

Same entry-point names, argument meaning and error behaviour as
``sofa::component::mapping::nonlinear::KinematicChainMapping<Vec1Types, Rigid3Types, Rigid3Types>``
(reference ``src/sofa/RigidBodyDynamics/KinematicChainMapping.h:37-129``), so the parity tests read like tests
of the reference component:

* public overloads take *lists* of state vectors and mark the component ``Invalid`` on an arity mismatch, after
  which every entry point returns silently (``KinematicChainMapping.inl:105-295`` and the ``d_componentState``
  guards at ``:346,406,458,523``);
* ``apply`` / ``applyJ`` overwrite, both ``applyJT`` accumulate (``:385,394,436,446`` vs ``:498-511,592-612``);
* ``applyJT`` writes the root wrench at index 0 of input2 although ``apply`` reads ``indexInput2`` (``:498`` vs
  ``:356``) — replicated;
* ``applyDJT`` is a no-op and ``getJ`` returns ``None`` (``KinematicChainMapping.h:73-83``).

All arithmetic happens in ``librbd_b200.so`` through the ``rbd_*_host`` entry points; arrays are numpy views of
what would be the SOFA ``VecCoord``/``VecDeriv`` storage of N concatenated robots:

    input1 (Vec1):    [N, nq_joints]           input2 (Rigid3): [N, n_root, 7] / [N, n_root, 6]
    output (Rigid3):  [N, n_out, 7] positions, [N, n_out, 6] velocities / forces
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .api import Model, Workspace
from .model import ModelDescriptor

VALID, INVALID = "Valid", "Invalid"
K_MIN_TORQUE_SQRD = 1e-6 * 1e-6   # kMinTorqueSqrd, KinematicChainMapping.inl:43-45 (applied inside the library)

# MatrixDeriv of N robots: {(instance, row_index): [(col, 6-vector), ...]}
MatrixDerivIn = Dict[Tuple[int, int], List[Tuple[int, Sequence[float]]]]


class KinematicChainMapping:
    def __init__(self, desc: ModelDescriptor, n_instances: int = 1, device: int = 0, index_input2: Optional[int] = None):
        self.m_model = Model(desc)                          # setModel (KinematicChainMapping.inl:304-311)
        self._m_data = None                                 # pinocchio::Data analogue, created on first use (needs a GPU)
        self._device = int(device)
        self.n = int(n_instances)
        self.desc = desc
        self.d_componentState = VALID
        self.d_indexFromRoot = index_input2                 # "indexInput2", default = last index (inl:319-336)
        self.msgs: List[str] = []

    @property
    def m_data(self) -> Workspace:
        if self._m_data is None:
            self._m_data = Workspace(self.m_model, max(1, self.n), self._device)
        return self._m_data

    # ---- helpers ----------------------------------------------------------------------------------------
    def _invalid(self, msg: str) -> None:
        self.msgs.append(msg)                               # msg_error()
        self.d_componentState = INVALID

    def _root_index(self, in_root: np.ndarray) -> int:
        size = in_root.shape[1]
        idx = self.d_indexFromRoot
        if idx is None or idx >= size:                      # checkIndexFromRoot, inl:319-336
            idx = size - 1
            self.d_indexFromRoot = idx
        return idx

    @staticmethod
    def _c(a: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=np.float64)

    # ---- public overloads (arity checks, inl:105-295) -------------------------------------------------------
    def apply(self, out_pos: List[np.ndarray], in1_pos: List[np.ndarray], in2_pos: List[np.ndarray]) -> None:
        if self.d_componentState == INVALID:
            return
        if not out_pos or not in1_pos:                      # inl:117-118: empty -> silent return, state untouched
            return
        if len(out_pos) > 1:
            return self._invalid("KinematicChainMapping only supports output vector of size 1")
        if len(in1_pos) > 1:
            return self._invalid("KinematicChainMapping only supports input vector of size 1")
        if len(in2_pos) > 1:
            return self._invalid("KinematicChainMapping only supports input root vector of size 1")
        self._apply(out_pos[0], in1_pos[0], in2_pos[0] if in2_pos else None)

    def applyJ(self, out_vel: List[np.ndarray], in1_vel: List[np.ndarray], in2_vel: List[np.ndarray]) -> None:
        if self.d_componentState == INVALID:
            return
        if not out_vel or not in1_vel:                      # inl:165-166
            return
        if len(out_vel) > 1 or len(in1_vel) > 1 or len(in2_vel) > 1:
            return self._invalid("applyJ: KinematicChainMapping only supports state vectors of size 1")
        self._applyJ(out_vel[0], in1_vel[0], in2_vel[0] if in2_vel else None)

    def applyJT(self, out1_force: List[np.ndarray], out2_force: List[np.ndarray], in_force: List[np.ndarray]) -> None:
        if self.d_componentState == INVALID:
            return
        if not in_force or not out1_force:                  # inl:212-213
            return
        if len(in_force) != 1 or len(out1_force) > 1 or len(out2_force) > 1:
            return self._invalid("applyJT: KinematicChainMapping only supports state vectors of size 1")
        self._applyJT(out1_force[0], out2_force[0] if out2_force else None, in_force[0])

    def applyJT_matrix(self, in_const: MatrixDerivIn, with_root: Optional[bool] = None):
        """``applyJT(ConstraintParams, MatrixDeriv...)`` (inl:249-295 -> :515-639).  Returns ``(out, out_root)``:
        ``out[(instance,row)]`` = dense joint-torque row (nv_joints entries), ``out_root[(instance,row)]`` = root
        wrench (6) written at column 0; rows whose squared norm is <= 1e-12 are dropped (inl:587,605)."""
        if self.d_componentState == INVALID:
            return {}, {}
        I = self.m_model.info
        ff = bool(I.has_freeflyer_root) if with_root is None else with_root
        keys = list(in_const.keys())
        row_ptr = np.zeros(len(keys) + 1, dtype=np.int32)
        row_inst = np.zeros(len(keys), dtype=np.int32)
        cols: List[int] = []
        vals: List[float] = []
        for r, key in enumerate(keys):
            row_inst[r] = key[0]
            for c, v in in_const[key]:
                cols.append(int(c)); vals.extend(float(x) for x in v)
            row_ptr[r + 1] = len(cols)
        col = np.array(cols, dtype=np.int32)
        val = np.array(vals, dtype=np.float64)
        rows = np.zeros((len(keys), I.nv), dtype=np.float64)
        keep = np.zeros(len(keys), dtype=np.int32)
        if len(keys):
            self.m_data.applyJT_rows_host(len(keys), row_ptr, row_inst, col, val, rows, keep)
        out, out_root = {}, {}
        for r, key in enumerate(keys):
            if not keep[r]:
                continue                                    # "dismiss dofs resulting in null torque"
            if ff:
                out_root[key] = rows[r, :6].copy()          # oRoot.addCol(0, rootWrench), inl:590-592
                out[key] = rows[r, 6:].copy()               # inl:595-599
            else:
                out[key] = rows[r].copy()                   # inl:608-613
        return out, out_root

    def applyDJT(self, *args, **kwargs) -> None:            # KinematicChainMapping.h:73-80
        return None

    def getJ(self):                                         # KinematicChainMapping.h:83
        return None

    # ---- protected workers (inl:338-639) -----------------------------------------------------------------
    def _apply(self, out: np.ndarray, in1: np.ndarray, in_root: Optional[np.ndarray]) -> None:
        if self.d_componentState == INVALID:
            return
        I = self.m_model.info
        n = self.n
        use_root = bool(I.has_freeflyer_root) and in_root is not None   # `m_fromRootModel and inRoot != nullptr`
        assert out.shape == (n, I.n_out, 7), "output vector must hold njoints + nframes0 rigids (inl:379)"
        qj = self._c(in1[:, : (I.nq - 7 if use_root else I.nq)])
        root = self._c(in_root[:, self._root_index(in_root), :]) if use_root else None
        res = out if (out.flags.c_contiguous and out.dtype == np.float64) else np.empty((n, I.n_out, 7))
        self.m_data.apply_host(n, qj, root, res)
        if res is not out:
            out[...] = res

    def _applyJ(self, out: np.ndarray, in1: np.ndarray, in_root: Optional[np.ndarray]) -> None:
        if self.d_componentState == INVALID:
            return
        I = self.m_model.info
        n = self.n
        use_root = bool(I.has_freeflyer_root) and in_root is not None
        assert out.shape == (n, I.n_out, 6)
        dq = self._c(in1[:, : (I.nv - 6 if use_root else I.nv)])
        rv = self._c(in_root[:, self._root_index(in_root), :]) if use_root else None
        res = out if (out.flags.c_contiguous and out.dtype == np.float64) else np.empty((n, I.n_out, 6))
        self.m_data.applyJ_host(n, dq, rv, res)
        if res is not out:
            out[...] = res

    def _applyJT(self, out: np.ndarray, out_root: Optional[np.ndarray], inp: np.ndarray) -> None:
        if self.d_componentState == INVALID:
            return
        I = self.m_model.info
        n = self.n
        use_root = bool(I.has_freeflyer_root) and out_root is not None
        w = self._c(inp.reshape(n, I.n_out * 6))
        kv = I.nv - 6 if use_root else I.nv
        tau = self._c(out[:, :kv])
        root = self._c(out_root[:, 0, :]) if use_root else None        # (*outRoot)[0] +=, inl:498
        self.m_data.applyJT_host(n, w, tau, root)
        out[:, :kv] = tau
        if use_root:
            out_root[:, 0, :] = root


class MappingBatch:
    """N ``KinematicChainMapping`` components of ONE scene — N robots loaded from the same URDF — sharing one workspace
    (C ABI: ``rbd_batch_*``).  Every component keeps the reference's worker signatures; the first
    component whose ``apply`` / ``applyJ`` / ``applyJT`` runs in a propagation pass (``epoch``) triggers ONE batched
    launch for all robots, the others find their output vectors already filled.

    The per-robot state vectors are separate numpy arrays (SOFA allocates every MechanicalObject on its own); they are
    bound by address, so they must be C-contiguous float64 and stay alive while bound."""

    def __init__(self, desc: ModelDescriptor, n_robots: int, device: int = 0):
        import ctypes as C
        from . import _capi
        self._C = C
        self.desc = desc
        self.n = int(n_robots)
        self.model = Model(desc)
        self.ws = Workspace(self.model, self.n, device)
        self._lib = self.model._lib
        h = C.c_void_p()
        _capi.check(self._lib.rbd_batch_create(self.ws._h, self.n, C.byref(h)))
        self._h = h
        self.epoch = 0
        self._keep: Dict[Tuple[str, int], tuple] = {}
        self.components = [_BatchedComponent(self, s) for s in range(self.n)]

    def begin_pass(self) -> None:
        """A new propagation pass of the scene (SOFA: one visitor traversal)."""
        self.epoch += 1

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rbd_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(a: Optional[np.ndarray]):
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
            raise ValueError("batched components bind C-contiguous float64 arrays by address")
        return int(a.ctypes.data)


class _BatchedComponent:
    """One robot's mapping inside a :class:`MappingBatch`.  ``bind_*`` is what the component does once in ``init()``
    (the addresses of its SOFA state vectors); ``apply`` / ``applyJ`` / ``applyJT`` are the protected worker overloads
    of the reference (KinematicChainMapping.h:103-117) for a single robot, served by the shared batched launch: they
    return True when this call ran it, False when an earlier component of the same pass already had."""

    def __init__(self, batch: MappingBatch, slot: int):
        self.batch, self.slot = batch, slot

    def bind_apply(self, out: np.ndarray, in1: np.ndarray, in_root: Optional[np.ndarray]) -> None:
        from . import _capi
        b = self.batch
        b._keep[("a", self.slot)] = (out, in1, in_root)
        _capi.check(b._lib.rbd_batch_bind_apply(b._h, self.slot, b._ptr(in1), b._ptr(in_root), b._ptr(out)))

    def bind_applyJ(self, out: np.ndarray, in1: np.ndarray, in_root: Optional[np.ndarray]) -> None:
        from . import _capi
        b = self.batch
        b._keep[("j", self.slot)] = (out, in1, in_root)
        _capi.check(b._lib.rbd_batch_bind_applyJ(b._h, self.slot, b._ptr(in1), b._ptr(in_root), b._ptr(out)))

    def bind_applyJT(self, out: np.ndarray, out_root: Optional[np.ndarray], inp: np.ndarray) -> None:
        from . import _capi
        b = self.batch
        b._keep[("t", self.slot)] = (out, out_root, inp)
        _capi.check(b._lib.rbd_batch_bind_applyJT(b._h, self.slot, b._ptr(inp), b._ptr(out), b._ptr(out_root)))

    def _call(self, fn) -> bool:
        from . import _capi
        rc = fn(self.batch._h, self.slot, self.batch.epoch)
        if rc < 0:
            _capi.check(rc)
        return rc == 1

    def apply(self) -> bool:
        return self._call(self.batch._lib.rbd_batch_apply)

    def applyJ(self) -> bool:
        return self._call(self.batch._lib.rbd_batch_applyJ)

    def applyJT(self) -> bool:
        return self._call(self.batch._lib.rbd_batch_applyJT)

"""Host-side mirror of the SOFA ``Mass`` / ``ForceField`` interface in JOINT space for N robot instances at once.

In the reference scene the mass of a robot never appears as one component: the loader creates one
``UniformMass<Rigid3Types>`` per body below the mapping (reference ``src/sofa/RigidBodyDynamics/URDFModelLoader.cpp:369-391``)
and SOFA's solvers reach the joint dofs matrix-free through ``KinematicChainMapping::applyJ`` / ``applyJT``:

    Mass::addMDx(res, dx, factor)    ==  res += factor * applyJT(M_b * applyJ(dx))      (CG solvers, every iteration)
    ForceField::addForce(f, x, v)    ==  f   += applyJT(m_b * g)                       (gravity; applyDJT is a no-op,
                                                                                       KinematicChainMapping.h:73-80,
                                                                                       so there is no velocity term)
    Mass::addMToMatrix               ==  not available: getJ() returns nullptr (KinematicChainMapping.h:83)

``JointSpaceMass`` offers the same three operations directly on the joint dofs, with the method names and the
accumulate-in-place semantics of ``sofa::core::behavior::Mass`` so that a scene-side component can forward to it
one-to-one; all arithmetic happens in ``librbd_b200.so`` (``rbd_addMDx_host``, ``rbd_addForce_host``, ``rbd_eval_host``).
The inertias are pinocchio's exact rigid-body inertias (CRBA / RNEA semantics, BASELINE north-star); SOFA's
``UniformMass<Rigid3>`` keeps the body inertia un-rotated, so the two agree exactly only for isotropic bodies
(SURVEY.md §3.5).

Arrays are numpy views of what would be the SOFA ``VecCoord`` / ``VecDeriv`` storage of N concatenated robots:
``x`` [N, nq] (free-flyer first: x y z qx qy qz qw), ``v``, ``dx``, ``f``, ``res`` [N, nv].
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .api import Model, Workspace
from .model import ModelDescriptor


class JointSpaceMass:
    def __init__(self, desc: ModelDescriptor, n_instances: int = 1, device: int = 0):
        self.m_model = Model(desc)
        self.desc = desc
        self.n = int(n_instances)
        self._device = int(device)
        self._ws: Optional[Workspace] = None

    @property
    def workspace(self) -> Workspace:
        if self._ws is None:
            self._ws = Workspace(self.m_model, max(1, self.n), self._device)
        return self._ws

    def _check(self, name: str, a: np.ndarray, k: int) -> np.ndarray:
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != (self.n, k):
            raise ValueError(f"{name}: expected a C-contiguous float64 array of shape ({self.n}, {k}), got "
                             f"{a.dtype} {a.shape}")
        return a

    # ---- sofa::core::behavior::Mass ------------------------------------------------------------------
    def addMDx(self, res: np.ndarray, x: np.ndarray, dx: np.ndarray, factor: float = 1.0) -> None:
        """res += factor * M(x) dx   (matrix-free: one sweep, M is never formed)."""
        I = self.m_model.info
        self.workspace.addMDx_host(self.n, self._check("x", x, I.nq), self._check("dx", dx, I.nv), factor,
                                   self._check("res", res, I.nv))

    def accFromF(self, a, x, f):  # noqa: D401 - SOFA name
        """Not provided: it needs M(x)^-1 (forward dynamics)."""
        raise NotImplementedError("accFromF needs the inverse mass matrix (ABA): not part of the reference path")

    def getMassMatrix(self, x: np.ndarray) -> np.ndarray:
        """Dense M(x) [N, nv, nv] (what addMToMatrix would assemble), from the packed upper triangle of CRBA."""
        I = self.m_model.info
        nv = I.nv
        Mp = np.empty((self.n, I.mass_fields))
        self.workspace.eval_host(self.n, self._check("x", x, I.nq), M=Mp)
        M = np.empty((self.n, nv, nv))
        for c in range(nv):
            for r in range(c + 1):
                M[:, r, c] = M[:, c, r] = Mp[:, c * (c + 1) // 2 + r]
        return M

    # ---- sofa::core::behavior::ForceField ------------------------------------------------------------
    def addForce(self, f: np.ndarray, x: np.ndarray, v: Optional[np.ndarray] = None) -> None:
        """f -= C(x, v) v + g(x)  (= f -= rnea(x, v, 0)); v None: the gravity term alone, f += sum_b J_b^T m_b g,
        which is what the reference's UniformMass + applyJT route adds."""
        I = self.m_model.info
        self.workspace.addForce_host(self.n, self._check("x", x, I.nq), None if v is None else self._check("v", v, I.nv),
                                     self._check("f", f, I.nv))

    def addDForce(self, df, dx, kFactor=0.0, bFactor=0.0):  # noqa: D401 - SOFA name
        """Gravity and bias forces contribute no stiffness through this mapping (applyDJT is a no-op in the
        reference, KinematicChainMapping.h:73-80): nothing to add."""
        return None

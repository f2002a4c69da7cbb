"""Independent second implementation of the hot-path mathematics — TEST INFRASTRUCTURE ONLY.

Purpose: the C oracle (``rbd_oracle.c``) restates pinocchio's *spatial-algebra recursions*.  Since pinocchio
itself is unavailable (parity unpinned), the oracle is pinned against this file, which derives the same
quantities by *different* routes, in double precision or — the same code on ``dtype=object`` arrays of
``mpmath.mpf`` — in 50-digit arithmetic:

* forward kinematics with 4x4 homogeneous matrices,
* Jacobians from the textbook geometric formula (axis x lever arm) instead of spatial ``act``/translate,
* mass matrix as  sum_b J_b^T diag(m_b I3, R_b I_b R_b^T) J_b  over body-CoM Jacobians (what the SOFA graph
  of the reference computes through UniformMass + applyJ/applyJT, URDFModelLoader.cpp:369-391) instead of CRBA,
* inverse dynamics from classical 3-vector Newton-Euler kinematics + Kane projection instead of spatial RNEA.

Nothing here is imported by the product.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import numpy as np

from sofa.rigidbodydynamics_b200 import model as mdl


class Arith:
    """Scalar backend: float64 or mpmath."""

    def __init__(self, mp=None):
        self.mp = mp
        if mp is None:
            self.dtype = np.float64
            self.conv = float
            self.sin, self.cos = math.sin, math.cos
        else:
            self.dtype = object
            self.conv = mp.mpf
            self.sin, self.cos = mp.sin, mp.cos

    def arr(self, x):
        a = np.asarray(x, dtype=object if self.mp else np.float64)
        if self.mp:
            a = np.vectorize(self.mp.mpf, otypes=[object])(a) if a.size else a
        return a

    def zeros(self, *shape):
        z = np.zeros(shape, dtype=self.dtype)
        if self.mp:
            z = z + self.mp.mpf(0)
        return z

    def eye(self, n):
        e = self.zeros(n, n)
        for i in range(n):
            e[i, i] = self.conv(1)
        return e


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=a.dtype)


def _axis_rot(ar: Arith, ax, c, s):
    """Rodrigues: c I + (1-c) a a^T + s [a]x."""
    I = ar.eye(3)
    K = np.array([[0 * c, -ax[2], ax[1]], [ax[2], 0 * c, -ax[0]], [-ax[1], ax[0], 0 * c]], dtype=ar.dtype)
    return c * I + (1 - c) * np.outer(ax, ax) + s * K


class NumpyDynamics:
    def __init__(self, d: mdl.ModelDescriptor, mp=None):
        self.d = d
        self.ar = Arith(mp)
        ar = self.ar
        self.pl = [self._hom(ar.arr(d.placement[i][:9]).reshape(3, 3), ar.arr(d.placement[i][9:])) for i in range(d.njoints)]
        self.fpl = [self._hom(ar.arr(d.frame_placement[f][:9]).reshape(3, 3), ar.arr(d.frame_placement[f][9:]))
                    for f in range(d.nframes)]
        self.axis = [ar.arr(d.axis[i]) for i in range(d.njoints)]
        self.mass = [ar.conv(float(d.inertia[i][0])) for i in range(d.njoints)]
        self.com = [ar.arr(d.inertia[i][1:4]) for i in range(d.njoints)]
        self.Ic = []
        for i in range(d.njoints):
            y = d.inertia[i]
            self.Ic.append(ar.arr([[y[4], y[5], y[7]], [y[5], y[6], y[8]], [y[7], y[8], y[9]]]))
        self.g = ar.arr(d.gravity)

    def _hom(self, R, p):
        T = self.ar.eye(4)
        T[:3, :3] = R
        T[:3, 3] = p
        return T

    # ---- kinematics -----------------------------------------------------------------------------------
    def joint_hom(self, i: int, q) -> np.ndarray:
        d, ar = self.d, self.ar
        t = int(d.jtype[i]); o = int(d.idx_q[i]); ax = self.axis[i]
        if t in (mdl.J_RX, mdl.J_RY, mdl.J_RZ, mdl.J_RU):
            c, s = ar.cos(q[o]), ar.sin(q[o])
            return self._hom(_axis_rot(ar, ax, c, s), ar.zeros(3))
        if t in (mdl.J_RUBX, mdl.J_RUBY, mdl.J_RUBZ, mdl.J_RUBU):
            return self._hom(_axis_rot(ar, ax, q[o], q[o + 1]), ar.zeros(3))
        if t in (mdl.J_PX, mdl.J_PY, mdl.J_PZ, mdl.J_PU):
            return self._hom(ar.eye(3), ax * q[o])
        if t == mdl.J_FF:
            x, y, z, w = q[o + 3], q[o + 4], q[o + 5], q[o + 6]
            # standard quaternion rotation formula, un-normalised exactly like Eigen's toRotationMatrix
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=ar.dtype)
            return self._hom(R, np.array([q[o], q[o + 1], q[o + 2]], dtype=ar.dtype))
        raise ValueError(t)

    def fk(self, q) -> List[np.ndarray]:
        q = self.ar.arr(q)
        T = [self.ar.eye(4)]
        for i in range(1, self.d.njoints):
            T.append(T[int(self.d.parents[i])] @ self.pl[i] @ self.joint_hom(i, q))
        return T

    def frame_poses(self, T) -> List[np.ndarray]:
        return [T[int(self.d.frame_parent[f])] @ self.fpl[f] for f in range(self.d.nframes)]

    def ancestors(self, j: int) -> List[int]:
        out = []
        while j > 0:
            out.append(j); j = int(self.d.parents[j])
        return out[::-1]

    def point_jacobian(self, T, joint: int, point_w) -> np.ndarray:
        """6 x nv LOCAL_WORLD_ALIGNED Jacobian of a point rigidly attached to `joint`: [dv/dqd; dw/dqd]."""
        d, ar = self.d, self.ar
        J = ar.zeros(6, d.nv)
        for k in self.ancestors(joint):
            t = int(d.jtype[k]); c = int(d.idx_v[k])
            Rk, pk = T[k][:3, :3], T[k][:3, 3]
            if t == mdl.J_FF:
                for e in range(3):
                    J[:3, c + e] = Rk[:, e]                               # body-frame linear velocity
                    J[:3, c + 3 + e] = cross(Rk[:, e], point_w - pk)      # body-frame angular velocity
                    J[3:, c + 3 + e] = Rk[:, e]
            else:
                aw = Rk @ self.axis[k]
                if t in (mdl.J_PX, mdl.J_PY, mdl.J_PZ, mdl.J_PU):
                    J[:3, c] = aw
                else:
                    J[:3, c] = cross(aw, point_w - pk)
                    J[3:, c] = aw
        return J

    def world_joint_jacobian(self, T) -> np.ndarray:
        """pinocchio data.J: column = velocity of the body point coincident with the WORLD ORIGIN; [lin; ang]."""
        d, ar = self.d, self.ar
        J = ar.zeros(6, d.nv)
        origin = ar.zeros(3)
        for k in range(1, d.njoints):
            Jk = self.point_jacobian(T, k, origin)
            for c in range(int(d.idx_v[k]), int(d.idx_v[k]) + d.nv_joint(k)):
                J[:, c] = Jk[:, c]
        return J

    # ---- mapping path ---------------------------------------------------------------------------------
    def full_q(self, q_joints, root_pose=None):
        if self.d.has_freeflyer_root and root_pose is not None:
            return np.concatenate([self.ar.arr(root_pose), self.ar.arr(q_joints)])
        return self.ar.arr(q_joints)

    def full_dq(self, dq_joints, root_vel=None):
        if self.d.has_freeflyer_root and root_vel is not None:
            return np.concatenate([self.ar.arr(root_vel), self.ar.arr(dq_joints)])
        return self.ar.arr(dq_joints)

    def apply(self, q_joints, root_pose=None) -> List[np.ndarray]:
        """World 4x4 pose of every mapped output (bodyCoMFrames then extraFrames)."""
        T = self.fk(self.full_q(q_joints, root_pose))
        F = self.frame_poses(T)
        return [F[int(f)] for f in self.d.out_frames]

    def out_jacobians(self, q_full) -> List[np.ndarray]:
        T = self.fk(q_full)
        F = self.frame_poses(T)
        return [self.point_jacobian(T, int(self.d.frame_parent[int(f)]), F[int(f)][:3, 3]) for f in self.d.out_frames]

    def applyJ(self, q_full, dq_full) -> np.ndarray:
        dq = self.ar.arr(dq_full)
        return np.array([J @ dq for J in self.out_jacobians(q_full)], dtype=self.ar.dtype)

    def applyJT(self, q_full, wrenches) -> np.ndarray:
        w = self.ar.arr(wrenches).reshape(-1, 6)
        tau = self.ar.zeros(self.d.nv)
        for k, J in enumerate(self.out_jacobians(q_full)):
            tau = tau + J.T @ w[k]
        return tau

    # ---- dynamics ---------------------------------------------------------------------------------------
    def mass_matrix(self, q_full) -> np.ndarray:
        d, ar = self.d, self.ar
        T = self.fk(q_full)
        M = ar.zeros(d.nv, d.nv)
        for b in range(1, d.njoints):
            Rb = T[b][:3, :3]
            cw = T[b][:3, 3] + Rb @ self.com[b]
            J = self.point_jacobian(T, b, cw)
            Iw = Rb @ self.Ic[b] @ Rb.T
            M = M + self.mass[b] * (J[:3].T @ J[:3]) + J[3:].T @ Iw @ J[3:]
        return M

    def inverse_dynamics(self, q_full, v_full, a_full) -> np.ndarray:
        """Classical Newton-Euler kinematics (3-vectors, world frame) + Kane projection onto CoM Jacobians."""
        d, ar = self.d, self.ar
        q = ar.arr(q_full); v = ar.arr(v_full); a = ar.arr(a_full)
        T = self.fk(q)
        z3 = ar.zeros(3)
        om = [z3] * d.njoints; al = [z3] * d.njoints; pd = [z3] * d.njoints; pdd = [z3] * d.njoints
        for i in range(1, d.njoints):
            p = int(d.parents[i]); t = int(d.jtype[i]); c = int(d.idx_v[i])
            Ri, pi = T[i][:3, :3], T[i][:3, 3]
            r = pi - T[p][:3, 3]
            if t == mdl.J_FF:
                vb, wb, vbd, wbd = v[c:c + 3], v[c + 3:c + 6], a[c:c + 3], a[c + 3:c + 6]
                w_rel = Ri @ wb
                v_rel = Ri @ vb
                a_rel = Ri @ (vbd + cross(wb, vb))
                al_rel = Ri @ wbd
            elif t in (mdl.J_PX, mdl.J_PY, mdl.J_PZ, mdl.J_PU):
                aw = Ri @ self.axis[i]
                w_rel = z3; v_rel = aw * v[c]; a_rel = aw * a[c]; al_rel = z3
            else:
                aw = Ri @ self.axis[i]
                w_rel = aw * v[c]; v_rel = z3; a_rel = z3; al_rel = aw * a[c]
            om[i] = om[p] + w_rel
            # w_rel = (relative rate) * (direction carried by the parent): its time derivative is
            # al_rel + omega_parent x w_rel  (for the free-flyer, omega_i x R_i w_b == omega_parent x R_i w_b)
            al[i] = al[p] + al_rel + cross(om[p], w_rel)
            # origin of joint i seen from the (rotating) parent frame: transport theorem
            pd[i] = pd[p] + cross(om[p], r) + v_rel
            pdd[i] = pdd[p] + cross(al[p], r) + cross(om[p], cross(om[p], r)) + 2 * cross(om[p], v_rel) + a_rel
        tau = ar.zeros(d.nv)
        for b in range(1, d.njoints):
            Rb = T[b][:3, :3]
            rc = Rb @ self.com[b]
            cw = T[b][:3, 3] + rc
            ac = pdd[b] + cross(al[b], rc) + cross(om[b], cross(om[b], rc))
            Iw = Rb @ self.Ic[b] @ Rb.T
            F = self.mass[b] * (ac - self.g)
            N = Iw @ al[b] + cross(om[b], Iw @ om[b])
            J = self.point_jacobian(T, b, cw)
            tau = tau + J[:3].T @ F + J[3:].T @ N
        return tau


def quat_xyzw_to_R(qv) -> np.ndarray:
    x, y, z, w = [float(t) for t in qv]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def to_float(a) -> np.ndarray:
    return np.array(a, dtype=object).astype(np.float64) if np.asarray(a).dtype == object else np.asarray(a, dtype=np.float64)

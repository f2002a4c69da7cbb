"""Generates tests/golden/golden_drnea.json: the partial derivatives of inverse dynamics tau = rnea(q, v, a),
d tau / d q (along the tangent of pinocchio's integrate: q (+) eps e_c) and d tau / d v, for one seeded state per model.

Derivation independent of both the oracle's restatement of pinocchio::computeRNEADerivatives (oracle/rbd_oracle.c:
orc_rnea_derivatives) and the CUDA kernel: CENTRAL FINITE DIFFERENCES of the independent inverse dynamics of
oracle/rbd_numpy.py (classical Newton-Euler kinematics + Kane projection — no spatial algebra, no recursion for the
derivatives) evaluated in 50-digit mpmath arithmetic with a step of 1e-20, so truncation (~1e-40) and rounding (~1e-30)
are both far below double precision; rounded to float64.  NOT pinocchio outputs (parity unpinned).

    python tests/golden/make_golden_drnea.py      (needs the *.model.json fixtures next to it; runs anywhere; minutes)
"""
import json
import os
import sys

import mpmath
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sofa.rigidbodydynamics_b200 import model as mdl  # noqa: E402
from oracle import rbd_numpy  # noqa: E402


def integrate_mp(d, q, c, eps):
    """q (+) eps e_c in mpmath: free-flyer p += R (eps e), quat <- quat * exp(eps e / 2) (velocity in the joint's own frame);
    unbounded revolute joints rotate their (cos, sin) pair; everything else q += eps."""
    mp = mpmath.mp
    out = list(q)
    for i in range(1, d.njoints):
        t, iq, iv = int(d.jtype[i]), int(d.idx_q[i]), int(d.idx_v[i])
        nvi = 6 if t == mdl.J_FF else 1
        if not (iv <= c < iv + nvi):
            continue
        k = c - iv
        if t == mdl.J_FF:
            x, y, z, w = q[iq + 3], q[iq + 4], q[iq + 5], q[iq + 6]
            if k < 3:
                R = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
                for r in range(3):
                    out[iq + r] = q[iq + r] + eps * R[r][k]
            else:
                dv = [mp.mpf(0)] * 3
                dv[k - 3] = mp.sin(eps / 2)
                dx, dy, dz, dw = dv[0], dv[1], dv[2], mp.cos(eps / 2)
                out[iq + 3] = w * dx + x * dw + y * dz - z * dy
                out[iq + 4] = w * dy - x * dz + y * dw + z * dx
                out[iq + 5] = w * dz + x * dy - y * dx + z * dw
                out[iq + 6] = w * dw - x * dx - y * dy - z * dz
        elif mdl.J_RUBX <= t <= mdl.J_RUBU:
            cs, sn = q[iq], q[iq + 1]
            out[iq] = cs * mp.cos(eps) - sn * mp.sin(eps)
            out[iq + 1] = sn * mp.cos(eps) + cs * mp.sin(eps)
        else:
            out[iq] = q[iq] + eps
    return out


def main():
    mpmath.mp.dps = 50
    mp = mpmath.mp
    import helpers
    eps = mp.mpf(10) ** -20
    out = {}
    for d in helpers.all_models():
        q, v, a = mdl.random_state(d, 1, seed=0xD7A0)
        nd = rbd_numpy.NumpyDynamics(d, mp=mp)
        qm = [mp.mpf(float(x)) for x in q[0]]
        vm = [mp.mpf(float(x)) for x in v[0]]
        am = [mp.mpf(float(x)) for x in a[0]]
        nv = d.nv
        dq = np.zeros((nv, nv)); dv = np.zeros((nv, nv))
        for c in range(nv):
            tp = nd.inverse_dynamics(integrate_mp(d, qm, c, eps), vm, am)
            tm = nd.inverse_dynamics(integrate_mp(d, qm, c, -eps), vm, am)
            for r in range(nv):
                dq[r, c] = float((tp[r] - tm[r]) / (2 * eps))
            vp = list(vm); vp[c] = vm[c] + eps
            vn = list(vm); vn[c] = vm[c] - eps
            tp = nd.inverse_dynamics(qm, vp, am)
            tm = nd.inverse_dynamics(qm, vn, am)
            for r in range(nv):
                dv[r, c] = float((tp[r] - tm[r]) / (2 * eps))
        out[d.name] = {"nq": d.nq, "nv": nv, "q": q[0].tolist(), "v": v[0].tolist(), "a": a[0].tolist(),
                       "dtau_dq": dq.tolist(), "dtau_dv": dv.tolist()}
        print(d.name, "done", flush=True)
    with open(os.path.join(HERE, "golden_drnea.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()

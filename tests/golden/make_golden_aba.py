"""Generates tests/golden/golden_aba.json: forward dynamics qdd = M(q)^-1 (tau - rnea(q, v, 0)) for a few seeded states
per model, from the INDEPENDENT implementation oracle/rbd_numpy.py (mass matrix = sum J^T Y J, Newton-Euler + Kane bias
forces) with the linear system solved in 50-digit mpmath arithmetic, rounded to float64.  Pins the oracle's
articulated-body algorithm (oracle/rbd_oracle.c: orc_aba) and the CUDA ABA kernel against a derivation that shares
nothing with either — no articulated inertias, no recursion.  NOT pinocchio outputs (parity unpinned).

    python tests/golden/make_golden_aba.py        (needs the *.model.json fixtures next to it; runs anywhere)
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


def main():
    mpmath.mp.dps = 50
    import helpers
    out = {}
    for d in helpers.all_models():
        nstates = 1 if d.nv > 30 else 2
        q, v, _a = mdl.random_state(d, nstates, seed=0xABA)
        rng = np.random.default_rng(0xF00D)
        tau = rng.uniform(-2, 2, (nstates, d.nv))
        nd = rbd_numpy.NumpyDynamics(d, mp=mpmath.mp)
        recs = []
        for s in range(nstates):
            M = nd.mass_matrix(nd.ar.arr(q[s]))
            b = nd.inverse_dynamics(q[s], v[s], np.zeros(d.nv))
            rhs = mpmath.matrix([mpmath.mpf(float(tau[s][i])) - b[i] for i in range(d.nv)])
            Mm = mpmath.matrix(d.nv, d.nv)
            for i in range(d.nv):
                for j in range(d.nv):
                    Mm[i, j] = M[i, j]
            x = mpmath.lu_solve(Mm, rhs)
            recs.append({"q": q[s].tolist(), "v": v[s].tolist(), "tau": tau[s].tolist(),
                         "qdd": [float(x[i]) for i in range(d.nv)]})
        out[d.name] = {"nq": d.nq, "nv": d.nv, "states": recs}
        print(d.name, "done", flush=True)
    with open(os.path.join(HERE, "golden_aba.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()

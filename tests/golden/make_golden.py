"""Generates the committed fixtures under tests/golden/.  Run ONCE in the build container (it reads the reference's
URDF assets under /root/reference, which do not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs
  *.model.json     model descriptors read from the reference's URDFs by sofa.rigidbodydynamics_b200.model.from_urdf
                   (derived constants, not a copy of the asset)
  golden_vectors.json   for each model a few seeded states with every hot-path output evaluated by the INDEPENDENT
                   implementation oracle/rbd_numpy.py in 50-digit mpmath arithmetic and rounded to float64.

These vectors are NOT pinocchio outputs (pinocchio is not installable here, SURVEY.md §8(c): parity unpinned); they
pin the C oracle and the CUDA kernels against a second derivation of the same mathematics at a precision where
rounding is irrelevant.
"""
import json
import os
import sys

import mpmath
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from sofa.rigidbodydynamics_b200 import model as mdl  # noqa: E402
from oracle import rbd_numpy  # noqa: E402

REF = "/root/reference/examples/sofa/RigidBodyDynamics"
URDFS = {
    "simple_arm": (REF + "/JointActuator/simple_arm.urdf", False),
    "schunk_svh_hand_right_regularized_inertia":
        (REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", False),
    "schunk_svh_hand_right_ff":
        (REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", True),
}


def f64(a):
    return np.array(a, dtype=object).astype(np.float64).tolist()


def main():
    mpmath.mp.dps = 50
    models = {}
    for name, (path, ff) in URDFS.items():
        d = mdl.from_urdf(path, free_flyer=ff)
        d.name = name
        with open(os.path.join(HERE, name + ".model.json"), "w") as f:
            f.write(d.to_json())
        models[name] = d
    for d in (mdl.panda7(), mdl.solo12(), mdl.talos_reduced(), mdl.random_tree(1, 10),
              mdl.random_tree(2, 12, free_flyer=True), mdl.random_tree(3, 20, branch_prob=0.6)):
        models[d.name] = d
    out = {}
    for name, d in models.items():
        nstates = 2 if d.nv > 30 else 3
        q, v, a = mdl.random_state(d, nstates, seed=0xC0FFEE)
        rng = np.random.default_rng(0xBEEF)
        w = rng.uniform(-1, 1, (nstates, d.n_out, 6))
        nd = rbd_numpy.NumpyDynamics(d, mp=mpmath.mp)
        recs = []
        for s in range(nstates):
            T = nd.fk(q[s])
            F = nd.frame_poses(T)
            rec = {"q": q[s].tolist(), "v": v[s].tolist(), "a": a[s].tolist(), "w": w[s].reshape(-1).tolist()}
            rec["oMi"] = [f64(np.concatenate([T[i][:3, :3].reshape(9), T[i][:3, 3]])) for i in range(1, d.njoints)]
            rec["J"] = f64(nd.world_joint_jacobian(T).T.reshape(-1))            # column-major 6 x nv
            rec["out_pose"] = [f64(np.concatenate([F[int(f)][:3, :3].reshape(9), F[int(f)][:3, 3]])) for f in d.out_frames]
            rec["applyJ"] = f64(nd.applyJ(q[s], v[s]).reshape(-1))              # dq = v used verbatim
            rec["applyJT"] = f64(nd.applyJT(q[s], w[s]))
            rec["M"] = f64(nd.mass_matrix(nd.ar.arr(q[s])).reshape(-1))          # full symmetric nv x nv
            rec["tau"] = f64(nd.inverse_dynamics(q[s], v[s], a[s]))
            recs.append(rec)
        out[name] = {"njoints": d.njoints, "nq": d.nq, "nv": d.nv, "n_out": d.n_out, "states": recs}
        print(name, "done", flush=True)
    with open(os.path.join(HERE, "golden_vectors.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()

"""The C ABI from compiled code: include/rbd_b200.h is valid C99, and tests/harness/kcm_harness.cpp — a C++ program that
links librbd_b200.so the way the patched reference component would and replays
KinematicChainMapping's call order on SOFA memory layouts — builds everywhere, fails cleanly without a GPU, and matches
the oracle on one."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from helpers import TOL_FP64, rel_err

HARNESS_DIR = os.path.join(helpers.HERE, "harness")


def _build_harness(rbd_lib):
    src = os.path.join(HARNESS_DIR, "kcm_harness.cpp")
    exe = os.path.join(HARNESS_DIR, "kcm_harness")
    libdir = os.path.join(helpers.ROOT, "sofa", "rigidbodydynamics_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe, src, "-L", libdir,
                           "-l:librbd_b200.so", "-Wl,-rpath," + libdir])
    return exe


def test_header_is_valid_c99(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "rbd_b200.h"\nint main(void) { return rbd_version() == RBD_B200_VERSION ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(helpers.ROOT, "include"), str(c)])


def test_harness_builds_and_fails_cleanly_without_gpu(rbd_lib):
    from sofa.rigidbodydynamics_b200 import api
    exe = _build_harness(rbd_lib)
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present (covered by the gpu test)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and r.stdout.startswith("ERROR -3"), r.stdout   # RBD_ERR_CUDA, no fallback


def _two_link_arm():
    """The model kcm_harness.cpp builds by hand."""
    from sofa.rigidbodydynamics_b200 import model as mdl
    b = mdl._Builder("two_link")
    Y = mdl.inertia_pack(1.0, (0.25, 0.0, 0.0), np.diag([0.01, 0.02, 0.02]))
    j1 = b.add_joint(0, mdl.J_RZ, mdl._se3(), "joint1")
    b.add_frame("joint1", j1, mdl._se3(), mdl.F_JOINT); b.add_frame("link1", j1, mdl._se3(), mdl.F_BODY)
    b.inertia[j1] = Y.copy()
    j2 = b.add_joint(j1, mdl.J_RZ, mdl._se3(None, (0.5, 0, 0)), "joint2")
    b.add_frame("joint2", j2, mdl._se3(), mdl.F_JOINT); b.add_frame("link2", j2, mdl._se3(), mdl.F_BODY)
    b.inertia[j2] = Y.copy()
    b.add_frame("tool", j2, mdl._se3(None, (0.5, 0, 0)), mdl.F_FIXED_JOINT)
    return b.finish()


@pytest.mark.gpu
def test_harness_matches_oracle(rbd_lib):
    import oracle
    exe = _build_harness(rbd_lib)
    q = np.array([0.3, -0.7])
    r = subprocess.run([exe, repr(float(q[0])), repr(float(q[1]))], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    rec = {}
    for line in r.stdout.splitlines():
        t = line.split()
        rec.setdefault(t[0], []).append([float(x) for x in t[1:]])
    d = _two_link_arm()
    n_out = d.n_out
    assert n_out == 9 and len(rec["pos"]) == n_out
    assert rec["noapply"][0][0] == -4                                  # RBD_ERR_NO_APPLY before the first apply
    orc = oracle.Oracle(d)
    pos = np.array([x[1:] for x in rec["pos"]]); vel = np.array([x[1:] for x in rec["vel"]])
    ref = orc.apply(q)
    assert rel_err(pos[:, :3], ref[:, :3]) <= TOL_FP64 and helpers.quat_err(pos[:, 3:], ref[:, 3:]) <= 1e-12
    # analytic tip position of a planar 2R arm
    tool = list(d.out_frames).index(d.frame_names.index("tool"))
    tip = [0.5 * np.cos(q[0]) + 0.5 * np.cos(q[0] + q[1]), 0.5 * np.sin(q[0]) + 0.5 * np.sin(q[0] + q[1]), 0.0]
    assert np.allclose(pos[tool, :3], tip, atol=1e-15)
    dq = np.array([0.4, -1.1])
    assert rel_err(vel, orc.applyJ(dq)) <= TOL_FP64
    w = np.array([[0.1 * (k + 1) * (e + 1) for e in range(3)] + [-0.05 * (k + e) for e in range(3)] for k in range(n_out)])
    tau_ref, _ = orc.applyJT(w.reshape(-1), np.array([1.0, 2.0]))
    assert rel_err(np.array(rec["tauJT"][0]), tau_ref) <= TOL_FP64
    row, keep = orc.applyJT_row([n_out - 1, 1], np.array([[1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0.5]], dtype=float))
    rr = rec["rows"][0]
    assert int(rr[0]) == int(keep) == 1 and rel_err(np.array(rr[1:3]), row) <= TOL_FP64
    assert int(rr[3]) == 0                                             # the all-zero row is dropped (kMinTorque)
    oMi, J, Mp, tau = orc.eval(q, dq, np.array([0.2, 0.3]))
    assert rel_err(np.array(rec["M"][0]), Mp) <= TOL_FP64 and rel_err(np.array(rec["tau"][0]), tau) <= TOL_FP64
    # Mass / ForceField products on host vectors
    Mfull = np.array([[Mp[0], Mp[1]], [Mp[1], Mp[2]]])
    assert rel_err(np.array(rec["mdx"][0]), np.array([10.0, 20.0]) + 2.0 * Mfull @ np.array([0.2, 0.3])) <= TOL_FP64
    _, _, _, bias = orc.eval(q, dq, np.zeros(2))
    assert rel_err(np.array(rec["fbias"][0]), -bias) <= TOL_FP64
    _, _, _, grav = orc.eval(q, np.zeros(2), np.zeros(2))
    assert rel_err(np.array(rec["fgrav"][0]), np.array([1.0, -1.0]) - grav) <= TOL_FP64

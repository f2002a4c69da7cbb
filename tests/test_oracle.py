"""Pins the CPU oracle (oracle/rbd_oracle.c).  Parity is UNPINNED against pinocchio itself (not installable here, the
reference ships no tests or golden vectors — SURVEY.md §8(c)); what pins the oracle instead:
  1. tests/golden/golden_vectors.json — every hot-path output from an independent derivation (oracle/rbd_numpy.py:
     4x4 homogeneous FK, geometric Jacobians, M = sum J^T Y J, Newton-Euler + Kane) in 50-digit mpmath arithmetic;
  2. the same independent implementation in double precision on fresh random states;
  3. physical identities and the analytic 1-dof arm of the reference's JointActuator example."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import TOL_FP64, rel_err

MODELS = helpers.all_models()
BYNAME = {m.name: m for m in MODELS}


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(helpers.GOLDEN, "golden_vectors.json")) as f:
        return json.load(f)


def test_oracle_matches_golden_vectors(golden):
    import oracle
    checked = 0
    for name, g in golden.items():
        d = BYNAME[name]
        orc = oracle.Oracle(d)
        for s in g["states"]:
            q, v, a, w = (np.array(s[k]) for k in ("q", "v", "a", "w"))
            oMi, J, Mp, tau = orc.eval(q, v, a)
            assert rel_err(oMi, np.array(s["oMi"]).reshape(-1)) <= 1e-13
            assert rel_err(J, np.array(s["J"])) <= 1e-13
            assert rel_err(oracle.unpack_upper(Mp, d.nv), np.array(s["M"]).reshape(d.nv, d.nv)) <= 1e-12
            assert rel_err(tau, np.array(s["tau"])) <= 1e-11
            if d.has_freeflyer_root:
                qj, root, dq, rootv = q[7:], q[:7], v[6:], v[:6]
            else:
                qj, root, dq, rootv = q, None, v, None
            pose = orc.apply(qj, root)
            ref_pose = np.array(s["out_pose"])
            assert rel_err(pose[:, :3], ref_pose[:, 9:]) <= 1e-13
            from oracle.rbd_numpy import quat_xyzw_to_R
            for k in range(d.n_out):
                assert np.max(np.abs(quat_xyzw_to_R(pose[k, 3:]) - ref_pose[k, :9].reshape(3, 3))) <= 1e-12
            assert rel_err(orc.applyJ(dq, rootv).reshape(-1), np.array(s["applyJ"])) <= 1e-12
            tj, tr = orc.applyJT(w)
            full = np.concatenate([tr, tj]) if d.has_freeflyer_root else tj
            assert rel_err(full, np.array(s["applyJT"])) <= 1e-12
            checked += 1
    assert checked >= 20


@pytest.mark.parametrize("d", MODELS, ids=[m.name for m in MODELS])
def test_oracle_matches_independent_numpy(d):
    import oracle
    from oracle.rbd_numpy import NumpyDynamics
    from sofa.rigidbodydynamics_b200 import model as mdl
    nd = NumpyDynamics(d)
    orc = oracle.Oracle(d)
    q, v, a = mdl.random_state(d, 3, seed=99)
    for i in range(3):
        oMi, J, Mp, tau = orc.eval(q[i], v[i], a[i])
        T = nd.fk(q[i])
        ref_oMi = np.concatenate([np.concatenate([T[k][:3, :3].reshape(9), T[k][:3, 3]]) for k in range(1, d.njoints)]) \
            if d.njoints > 1 else np.zeros(0)
        assert rel_err(oMi, ref_oMi) <= 1e-12
        assert rel_err(J, nd.world_joint_jacobian(T).T.reshape(-1)) <= 1e-12
        assert rel_err(oracle.unpack_upper(Mp, d.nv), nd.mass_matrix(q[i])) <= 1e-11
        assert rel_err(tau, nd.inverse_dynamics(q[i], v[i], a[i])) <= 1e-10


@pytest.mark.parametrize("name", ["panda7", "solo12", "talos_reduced", "random_3",
                                  "schunk_svh_hand_right_regularized_inertia"])
def test_identities(name):
    """rnea(q,v,a) - rnea(q,v,0) = M a;  rnea(q,0,0) = -sum_b J_b^T [m_b g; 0];  M symmetric positive definite;
    finite-difference check of the frame Jacobian."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = BYNAME[name]
    orc = oracle.Oracle(d)
    q, v, a = mdl.random_state(d, 2, seed=123)
    for i in range(2):
        M = oracle.unpack_upper(orc.eval(q[i], v[i], a[i])[2], d.nv)
        assert rel_err(orc.rnea(q[i], v[i], a[i]) - orc.rnea(q[i], v[i], 0 * a[i]), M @ a[i]) <= 1e-11
        assert np.all(np.linalg.eigvalsh(M) > 0)
        # gravity through the mapping: wrench m_b g on every body-CoM output (outputs 0..njoints-1)
        if d.has_freeflyer_root:
            orc.apply(q[i][7:], q[i][:7])
        else:
            orc.apply(q[i])
        w = np.zeros((d.n_out, 6))
        for b in range(d.njoints):
            w[b, :3] = d.inertia[b, 0] * d.gravity
        tj, tr = orc.applyJT(w.reshape(-1))
        full = np.concatenate([tr, tj]) if d.has_freeflyer_root else tj
        assert rel_err(-full, orc.rnea(q[i], 0 * v[i], 0 * a[i])) <= 1e-11
    # finite differences on a fixed-base model (configuration space is flat)
    if not d.has_freeflyer_root and all(int(t) in (1, 2, 3, 4, 5, 6, 7, 8) for t in d.jtype[1:]):
        q0 = q[0]; eps = 1e-6
        p0 = orc.apply(q0)[:, :3].copy()
        k = d.n_out - 1
        frame = int(d.out_frames[k])
        orc.apply(q0)
        Jf = orc.frame_jacobian_lwa(frame)
        for c in range(d.nv):
            dq = np.zeros(d.nv); dq[c] = eps
            pp = orc.apply(q0 + dq)[:, :3]; pm = orc.apply(q0 - dq)[:, :3]
            assert np.max(np.abs((pp[k] - pm[k]) / (2 * eps) - Jf[:3, c])) <= 1e-7
        assert p0.shape == (d.n_out, 3)


def test_simple_arm_analytic():
    """JointActuator/simple_arm.urdf (reference examples): 1 revolute Z joint; tool0 fixed at x = 0.07
    (JointActuator.py:74-79): tool0 = (0.07 cos q, 0.07 sin q, 0); M = Izz_c + m (cx^2 + cy^2)."""
    import oracle
    d = BYNAME["simple_arm"]
    assert d.nq == 1 and d.nv == 1 and d.njoints == 2
    orc = oracle.Oracle(d)
    tool = d.frame_names.index("tool0")
    k = list(d.out_frames).index(tool)
    m, cx, cy = d.inertia[1, 0], d.inertia[1, 1], d.inertia[1, 2]
    izz = d.inertia[1, 9]
    for q in (0.0, 0.3, -2.0, 3.1):
        pose = orc.apply(np.array([q]))
        assert np.allclose(pose[k, :3], [0.07 * np.cos(q), 0.07 * np.sin(q), 0.0], atol=1e-15)
        M = orc.crba(np.array([q]))
        assert abs(M[0, 0] - (izz + m * (cx * cx + cy * cy))) <= 1e-18 + 1e-14 * abs(M[0, 0])


def test_rot_to_quat_eigen_branches():
    """Eigen 3.4 Quaternion(Matrix3) branch order (Conversions.h:60): w from the trace branch, else largest diagonal."""
    import oracle
    from oracle.rbd_numpy import quat_xyzw_to_R
    I = np.eye(3)
    assert np.allclose(oracle.rot_to_quat(I), [0, 0, 0, 1])
    Rx = np.diag([1.0, -1.0, -1.0]); Ry = np.diag([-1.0, 1.0, -1.0]); Rz = np.diag([-1.0, -1.0, 1.0])
    assert np.allclose(oracle.rot_to_quat(Rx), [1, 0, 0, 0])
    assert np.allclose(oracle.rot_to_quat(Ry), [0, 1, 0, 0])
    assert np.allclose(oracle.rot_to_quat(Rz), [0, 0, 1, 0])
    rng = np.random.default_rng(0)
    for _ in range(200):
        A = rng.normal(size=(3, 3)); Q, _r = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        qv = oracle.rot_to_quat(Q)
        assert abs(np.linalg.norm(qv) - 1) < 1e-12
        assert np.max(np.abs(quat_xyzw_to_R(qv) - Q)) < 1e-12


def test_flop_counting_build_counts_the_restated_algorithm():
    """oracle -DORC_COUNT_FLOPS: the same source with a counting scalar.  The counts are a function of
    the model only, the three algorithms add up to the fused evaluation, and the counting build computes the same
    numbers as the product oracle."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    for d in (mdl.panda7(), mdl.solo12(), mdl.talos_reduced()):
        q, v, a = mdl.random_state(d, 2, seed=9)
        c0 = oracle.flop_counts(d, q[0], v[0], a[0]); c1 = oracle.flop_counts(d, q[1], v[1], a[1])
        assert c0 == c1
        for k in ("add", "mul", "div", "sqrt", "sincos", "flops"):
            assert c0["eval"][k] == c0["fk_jacobian"][k] + c0["crba"][k] + c0["rnea"][k]
        nrev = sum(1 for t in d.jtype[1:] if mdl.J_RX <= int(t) <= mdl.J_RU)
        assert c0["fk_jacobian"]["sincos"] == 2 * nrev          # one sine + one cosine per revolute joint and FK pass
        assert c0["eval"]["flops"] > 100 * d.nv


def test_aba_matches_independent_forward_dynamics():
    """orc_aba (pinocchio's articulated-body algorithm, LOCAL convention) against tests/golden/golden_aba.json: forward
    dynamics solved as M(q)^-1 (tau - nle) in 50-digit arithmetic by the independent implementation (no articulated
    inertias, no recursion; tests/golden/make_golden_aba.py), and the round trip aba(q, v, rnea(q, v, a)) == a."""
    import json
    import os
    import helpers
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    with open(os.path.join(helpers.GOLDEN, "golden_aba.json")) as f:
        gold = json.load(f)
    seen = 0
    for d in helpers.all_models():
        orc = oracle.Oracle(d)
        for st in gold[d.name]["states"]:
            qdd = orc.aba(np.array(st["q"]), np.array(st["v"]), np.array(st["tau"]))
            assert helpers.rel_err(qdd, np.array(st["qdd"])) <= 1e-9, d.name   # conditioning of M enters: not 1e-10
            seen += 1
        q, v, a = mdl.random_state(d, 3, seed=77)
        for s in range(3):
            assert helpers.rel_err(orc.aba(q[s], v[s], orc.rnea(q[s], v[s], a[s])), a[s]) <= 1e-10, d.name
    assert seen >= 11


def test_minv_is_the_solution_operator_of_aba():
    """Oracle.minv (inverse of the CRBA matrix) against the oracle's articulated-body algorithm on unit torques: column j
    == aba(q, 0, e_j) - aba(q, 0, 0) (v = 0: no Coriolis terms, gravity cancels) — two restatements with no shared code."""
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    for d in helpers.all_models():
        if d.nv == 0:
            continue
        orc = oracle.Oracle(d)
        q, _v, _a = mdl.random_state(d, 2, seed=9)
        for s_ in range(2):
            G = orc.minv(q[s_])
            z = np.zeros(d.nv)
            a0 = orc.aba(q[s_], z, z)
            for c in range(d.nv):
                e = np.zeros(d.nv); e[c] = 1.0
                col = orc.aba(q[s_], z, e) - a0
                assert np.max(np.abs(col - G[:, c])) <= 1e-9 * max(1.0, float(np.max(np.abs(G))))


def test_rnea_derivatives_match_independent_finite_differences():
    """orc_rnea_derivatives (pinocchio::computeRNEADerivatives restated with dense world-frame matrices) against
    tests/golden/golden_drnea.json: central differences of the independent Newton-Euler + Kane inverse dynamics in
    50-digit arithmetic (make_golden_drnea.py), d tau / d q taken along the tangent of integrate; dtau_da == crba; the
    structural zeros (dofs on different branches) are exact zeros; and, on fresh states, double-precision central
    differences of the oracle's own rnea."""
    import json
    import os
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    with open(os.path.join(helpers.GOLDEN, "golden_drnea.json")) as f:
        gold = json.load(f)
    for d in helpers.all_models():
        orc = oracle.Oracle(d)
        r = gold[d.name]
        q, v, a = np.array(r["q"]), np.array(r["v"]), np.array(r["a"])
        dq, dv, M = orc.rnea_derivatives(q, v, a)
        assert helpers.rel_err(dq, np.array(r["dtau_dq"])) <= 1e-12, d.name
        assert helpers.rel_err(dv, np.array(r["dtau_dv"])) <= 1e-12, d.name
        assert helpers.rel_err(np.triu(M), np.triu(orc.crba(q))) <= 1e-12, d.name
        related = helpers.support_relation(d)
        assert np.all(dq[~related] == 0.0) and np.all(dv[~related] == 0.0), d.name
        q, v, a = mdl.random_state(d, 1, seed=4242)
        dq, dv, _ = orc.rnea_derivatives(q[0], v[0], a[0])
        eps = 1e-6
        for c in range(d.nv):
            e = np.zeros((1, d.nv)); e[0, c] = 1.0
            fq = (orc.rnea(helpers.integrate(d, q, e, eps)[0], v[0], a[0]) - orc.rnea(helpers.integrate(d, q, e, -eps)[0], v[0], a[0])) / (2 * eps)
            fv = (orc.rnea(q[0], v[0] + eps * e[0], a[0]) - orc.rnea(q[0], v[0] - eps * e[0], a[0])) / (2 * eps)
            assert np.max(np.abs(fq - dq[:, c])) <= 1e-6 * max(1.0, np.max(np.abs(dq))), (d.name, c)
            assert np.max(np.abs(fv - dv[:, c])) <= 1e-6 * max(1.0, np.max(np.abs(dv))), (d.name, c)

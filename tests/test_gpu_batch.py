"""Multi-robot batching context (SURVEY.md §8(f) rank 1, C ABI rbd_batch_*): N robots of one scene, each with its own
separately allocated state vectors, served by one batched launch per propagation pass."""
import numpy as np
import pytest

from helpers import TOL_FP64, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot", ["solo12", "panda7"])
def test_scene_of_robots_matches_per_robot_oracle(rbd_lib, robot):
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200.mapping import MappingBatch
    d = mdl.ROBOTS[robot]()
    N = 37
    ff = d.has_freeflyer_root
    q, v, _ = mdl.random_state(d, N, seed=5)
    rng = np.random.default_rng(6)
    # every robot owns its vectors (separate allocations, as SOFA MechanicalObjects do)
    in1 = [np.ascontiguousarray(q[i, 7:] if ff else q[i]) for i in range(N)]
    root = [np.ascontiguousarray(q[i, :7]) if ff else None for i in range(N)]
    dq = [np.ascontiguousarray(v[i, 6:] if ff else v[i]) for i in range(N)]
    rootv = [np.ascontiguousarray(v[i, :6]) if ff else None for i in range(N)]
    pos = [np.zeros((d.n_out, 7)) for _ in range(N)]
    vel = [np.zeros((d.n_out, 6)) for _ in range(N)]
    w = [rng.uniform(-1, 1, (d.n_out, 6)) for _ in range(N)]
    tau = [rng.uniform(-1, 1, d.nv_joints) for _ in range(N)]
    rt = [rng.uniform(-1, 1, 6) if ff else None for _ in range(N)]
    tau0 = [t.copy() for t in tau]; rt0 = [r.copy() if r is not None else None for r in rt]
    batch = MappingBatch(d, N)
    for i, c in enumerate(batch.components):             # component init(): bind the robot's own vectors
        c.bind_apply(pos[i], in1[i], root[i]); c.bind_applyJ(vel[i], dq[i], rootv[i]); c.bind_applyJT(tau[i], rt[i], w[i])
    batch.begin_pass()
    ran = [c.apply() for c in batch.components]
    assert ran == [True] + [False] * (N - 1)             # the first component's call served the whole scene
    assert [c.applyJ() for c in batch.components] == [True] + [False] * (N - 1)
    assert [c.applyJT() for c in batch.components] == [True] + [False] * (N - 1)
    orc = oracle.Oracle(d)
    for i in range(N):
        rp = orc.apply(in1[i], root[i])
        assert rel_err(pos[i][:, :3], rp[:, :3]) <= TOL_FP64
        assert rel_err(vel[i], orc.applyJ(dq[i], rootv[i])) <= TOL_FP64
        rtau, rroot = orc.applyJT(w[i].reshape(-1), tau0[i], rt0[i])
        assert rel_err(tau[i], rtau) <= TOL_FP64
        if ff:
            assert rel_err(rt[i], rroot) <= TOL_FP64
    # a new pass with new joint values recomputes; within the pass a second apply is a no-op
    batch.begin_pass()
    in1[3][:] += 0.25
    assert batch.components[5].apply() is True
    assert batch.components[3].apply() is False
    assert rel_err(pos[3][:, :3], orc.apply(in1[3], root[3])[:, :3]) <= TOL_FP64


def test_unbound_slot_is_an_error(rbd_lib):
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    from sofa.rigidbodydynamics_b200.mapping import MappingBatch
    d = mdl.panda7()
    batch = MappingBatch(d, 3)
    batch.begin_pass()
    out = np.zeros((d.n_out, 7)); q = np.zeros(d.nq)
    batch.components[0].bind_apply(out, q, None)
    with pytest.raises(api.RbdError) as e:
        batch.components[0].apply()                      # slots 1 and 2 were never bound
    assert e.value.code == -1

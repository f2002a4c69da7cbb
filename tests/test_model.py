"""Model descriptor / URDF reader: the structural pins the reference's example scenes give and the
validation the C ABI applies in rbd_model_create."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from sofa.rigidbodydynamics_b200 import model as mdl

BYNAME = {m.name: m for m in helpers.all_models()}
REF = "/root/reference/examples/sofa/RigidBodyDynamics"


def test_schunk_hand_pins():
    d = BYNAME["schunk_svh_hand_right_regularized_inertia"]
    assert d.nq == 20 and d.nv == 20 and d.njoints == 21          # InverseKineHand.py:14-17: q_init has 20 entries
    # universe, root_joint (FIXED_JOINT added by pinocchio's UrdfVisitor::addRootJoint), base_link, then per URDF joint a
    # JOINT/FIXED_JOINT frame followed by the child BODY frame: 1 + 2 + 2*36 = 75
    assert d.nframes0 == 75
    assert d.n_out == 21 + 75
    # schunkHandWithConstraints.py:73-79: constraintBodyIndex = 33 in Frames/framesDof is "index tip": frame 33 is the
    # fixed joint that carries the fftip link (same placement as the fftip BODY frame that follows it)
    assert d.frame_names[33] == "fftip_joint" and d.frame_names[34] == "fftip"
    assert np.array_equal(d.frame_placement[33], d.frame_placement[34]) and d.frame_parent[33] == d.frame_parent[34]
    # joint 1 has axis 0 0 -1 -> RevoluteUnaligned; everything else is revolute Z
    assert int(d.jtype[1]) == mdl.J_RU and np.allclose(d.axis[1], [0, 0, -1])
    assert all(int(t) == mdl.J_RZ for t in d.jtype[2:])
    assert d.is_preorder()


def test_panda_solo_talos_shapes():
    p, s, t = mdl.panda7(), mdl.solo12(), mdl.talos_reduced()
    assert (p.nq, p.nv, p.njoints) == (7, 7, 8)
    assert p.joint_names[7] == "panda_joint7"                      # robotWithConstraints.py:73-79: jointsDof[7]
    assert (s.nq, s.nv, s.njoints) == (19, 18, 14) and s.has_freeflyer_root
    assert (t.nq, t.nv, t.njoints) == (39, 38, 34) and t.has_freeflyer_root
    assert t.eval_bytes() == 12144 and s.eval_bytes() == 4064 and p.eval_bytes() == 1456
    assert int(t.depth().max()) == 11
    for d in (p, s, t):
        assert d.n_out == d.njoints + d.nframes0
        # the loader's output table: CoM frames of every joint (universe included, kSkipUniverse = 0), then all frames
        assert list(d.out_frames[d.njoints:]) == list(range(d.nframes0))
        for j in range(d.njoints):
            f = int(d.out_frames[j])
            assert int(d.frame_parent[f]) == j and np.allclose(d.frame_placement[f][9:], d.inertia[j][1:4])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference assets only exist in the build container")
def test_urdf_reader_reproduces_committed_descriptors():
    for name, path, ff in (("simple_arm", REF + "/JointActuator/simple_arm.urdf", False),
                           ("schunk_svh_hand_right_regularized_inertia",
                            REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", False),
                           ("schunk_svh_hand_right_ff",
                            REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", True)):
        d = mdl.from_urdf(path, free_flyer=ff)
        g = BYNAME[name]
        for k in ("parents", "jtype", "idx_q", "idx_v", "frame_parent", "out_frames"):
            assert np.array_equal(getattr(d, k), getattr(g, k)), k
        for k in ("axis", "placement", "inertia", "frame_placement"):
            assert np.allclose(getattr(d, k), getattr(g, k), atol=0, rtol=0), k


def test_json_roundtrip():
    d = mdl.talos_reduced()
    e = mdl.ModelDescriptor.from_json(d.to_json())
    for k in ("parents", "jtype", "axis", "placement", "idx_q", "idx_v", "inertia", "frame_parent", "frame_placement",
              "gravity", "out_frames"):
        assert np.array_equal(getattr(d, k), getattr(e, k))


def _create(lib, d, mutate=None):
    from sofa.rigidbodydynamics_b200._capi import make_desc
    cd, keep = make_desc(d)
    arrays = {}
    if mutate:
        mutate(cd, arrays)
    h = C.c_void_p()
    rc = lib.rbd_model_create(C.byref(cd), C.byref(h))
    if rc == 0:
        lib.rbd_model_destroy(h)
    return rc, lib.rbd_last_error().decode()


def test_model_create_validation(rbd_lib):
    d = mdl.random_tree(3, 12, branch_prob=0.6)
    assert _create(rbd_lib, d)[0] == 0

    def bad_parent(cd, keep):
        a = np.array(d.parents, dtype=np.int32); a[3] = 5; keep["p"] = a
        cd.parents = a.ctypes.data_as(C.POINTER(C.c_int32))
    rc, msg = _create(rbd_lib, d, bad_parent)
    assert rc == -2 and "parents" in msg                           # RBD_ERR_MODEL

    def not_preorder(cd, keep):
        # a valid parents[i] < i array that is NOT a DFS pre-order: 1<-0, 2<-0, 3<-1
        a = np.zeros(d.njoints, dtype=np.int32); a[2] = 0; a[3] = 1; a[1] = 0
        for i in range(4, d.njoints):
            a[i] = i - 1
        keep["p"] = a
        cd.parents = a.ctypes.data_as(C.POINTER(C.c_int32))
    rc, msg = _create(rbd_lib, d, not_preorder)
    assert rc == -2 and "pre-order" in msg

    def bad_type(cd, keep):
        a = np.array(d.jtype, dtype=np.int32); a[2] = 77; keep["t"] = a
        cd.jtype = a.ctypes.data_as(C.POINTER(C.c_int32))
    assert _create(rbd_lib, d, bad_type)[0] == -5                  # RBD_ERR_UNSUPPORTED

    def null_arrays(cd, keep):
        cd.inertia = None
    assert _create(rbd_lib, d, null_arrays)[0] == -1               # RBD_ERR_INVALID_ARG

    big = mdl.random_tree(7, 70)
    rc, msg = _create(rbd_lib, big)
    assert rc == -5 and "RBD_MAX_JOINTS" in msg

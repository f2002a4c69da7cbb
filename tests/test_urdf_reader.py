"""The library's own URDF reader (csrc/rbd_urdf.cpp, rbd_urdf_load_*): what a C++ caller without
pinocchio uses in place of pinocchio::urdf::buildModel + the loader's frame tables (URDFModelLoader.cpp:126-183).

Checked against the independent Python reader (model.from_urdf_string) on generated URDF text of random trees (every
joint kind, fixed-joint merging, XML oddities), on the reference's own URDF files (build container only), on the
committed descriptors of those files (tests/golden/*.model.json: runs anywhere), and on malformed input."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import helpers
from sofa.rigidbodydynamics_b200 import _capi, api, model as mdl

REF = "/root/reference/examples/sofa/RigidBodyDynamics"
INT_KEYS = ("parents", "jtype", "idx_q", "idx_v", "frame_parent", "out_frames")
FLT_KEYS = ("axis", "placement", "inertia", "frame_placement", "gravity")


def _same(a: mdl.ModelDescriptor, b: mdl.ModelDescriptor, tol=1e-14):
    assert (a.njoints, a.nq, a.nv, a.nframes, a.nframes0, a.n_out) == (b.njoints, b.nq, b.nv, b.nframes, b.nframes0, b.n_out)
    for k in INT_KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in FLT_KEYS:
        x, y = np.asarray(getattr(a, k)), np.asarray(getattr(b, k))
        assert x.shape == y.shape, k
        # same formulas in the same order; libm vs numpy sin/cos may differ in the last bit
        assert np.max(np.abs(x - y), initial=0.0) <= tol * max(1.0, float(np.max(np.abs(y), initial=0.0))), k
    assert a.joint_names == b.joint_names and a.frame_names == b.frame_names and a.frame_types == b.frame_types


def _fmt(x):
    return " ".join(repr(float(v)) for v in x)


def random_urdf(seed: int, nlinks: int = 14, fancy_xml: bool = True) -> str:
    """URDF text of a random tree: every joint kind the reader supports (aligned / unaligned / negative axes, default
    axis, continuous, prismatic, floating, fixed), links without <inertial>, inertial origins with rotations, joint names
    that sort differently from their order of appearance, and (fancy_xml) comments, a DOCTYPE, CDATA, entities, single
    quotes, <mimic> / <limit> / <visual> children that must be ignored."""
    rng = np.random.default_rng(seed)
    kinds = ["revolute", "continuous", "prismatic", "fixed", "revolute", "revolute", "floating"]
    out = []
    if fancy_xml:
        out.append('\n  <?xml version="1.0" encoding="utf-8"?>\n<!DOCTYPE robot [ <!ELEMENT robot ANY> ]>\n<!-- generated -->')
    out.append(f'<robot name="rnd_{seed}" xmlns:xacro="http://www.ros.org/wiki/xacro">')
    names = [f"link_{i}" for i in range(nlinks)]

    def link(i):
        s = [f'  <link name="{names[i]}">']
        if i == 0 or rng.uniform() < 0.85:
            m = float(rng.uniform(0.1, 3.0))
            A = rng.normal(size=(3, 3)); I = A @ A.T * 0.01 + np.eye(3) * 0.01
            org = ""
            r = rng.uniform()
            if r < 0.4:
                org = f'<origin xyz="{_fmt(rng.uniform(-0.1, 0.1, 3))}" rpy="{_fmt(rng.uniform(-3, 3, 3))}"/>'
            elif r < 0.7:
                org = f"<origin xyz='{_fmt(rng.uniform(-0.1, 0.1, 3))}'/>"
            s.append(f'    <inertial>{org}<mass value="{m!r}"/>')
            if rng.uniform() < 0.2:
                s.append(f'      <inertia ixx="{float(I[0, 0])!r}" iyy="{float(I[1, 1])!r}" izz="{float(I[2, 2])!r}"/></inertial>')
            else:
                s.append(f'      <inertia ixx="{float(I[0, 0])!r}" ixy="{float(I[0, 1])!r}" ixz="{float(I[0, 2])!r}" iyy="{float(I[1, 1])!r}" '
                         f'iyz="{float(I[1, 2])!r}" izz="{float(I[2, 2])!r}"/>\n    </inertial>')
        if fancy_xml and rng.uniform() < 0.5:
            s.append('    <visual><geometry><mesh filename="package://a&amp;b/m.stl" scale="1 1 1"/></geometry>'
                     '<material name="x"><color rgba="0 0 0 1"/></material></visual><!-- </link> not yet -->')
        s.append("  </link>")
        return "\n".join(s)

    links = [link(i) for i in range(nlinks)]
    joints = []
    perm = rng.permutation(nlinks - 1)          # joint names sort in an order unrelated to the link numbering
    for i in range(1, nlinks):
        parent = int(rng.integers(0, i)) if rng.uniform() < 0.4 else i - 1
        kind = kinds[int(rng.integers(0, len(kinds)))]
        s = [f'  <joint name="jnt_{perm[i - 1]:03d}" type="{kind}">']
        if rng.uniform() < 0.9:
            s.append(f'    <origin rpy="{_fmt(rng.uniform(-math.pi, math.pi, 3))}" xyz="{_fmt(rng.uniform(-0.3, 0.3, 3))}"/>')
        s.append(f'    <parent link="{names[parent]}"/><child link=\'{names[i]}\'/>')
        if kind in ("revolute", "continuous", "prismatic"):
            r = rng.uniform()
            if r < 0.45:
                e = np.zeros(3); e[int(rng.integers(0, 3))] = 1.0
                s.append(f'    <axis xyz="{_fmt(e)}"/>')
            elif r < 0.6:
                e = np.zeros(3); e[int(rng.integers(0, 3))] = -1.0   # negative axes are "unaligned"
                s.append(f'    <axis xyz="{_fmt(e)}"/>')
            elif r < 0.9:
                s.append(f'    <axis xyz="{_fmt(rng.normal(size=3) * 2.0)}"/>')   # not normalised
            # else: no <axis> -> URDF default 1 0 0
            s.append('    <limit lower="-1" upper="1" effort="10" velocity="2"/><dynamics damping="0.1"/>')
            if fancy_xml and rng.uniform() < 0.2:
                s.append('    <mimic joint="jnt_000" multiplier="1" offset="0"/>')
        if fancy_xml and rng.uniform() < 0.3:
            s.append("    <![CDATA[ <joint> inside character data </joint> ]]>")
        s.append("  </joint>")
        joints.append("\n".join(s))
    body = links + joints
    if fancy_xml:
        order = rng.permutation(len(body))      # URDF does not order links and joints
        body = [body[k] for k in order]
        body.insert(len(body) // 2, "  <gazebo reference='link_0'><material>Gazebo/Grey</material> text &lt; </gazebo>")
    out.extend(body)
    out.append("</robot>\n<!-- trailing comment -->\n")
    return "\n".join(out)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("ff", [False, True])
def test_cpp_reader_equals_python_reader_on_random_urdf(rbd_lib, seed, ff):
    text = random_urdf(seed, nlinks=6 + 3 * seed, fancy_xml=seed % 3 != 0)
    a = api.load_urdf(text, free_flyer=ff, is_text=True)
    b = mdl.from_urdf_string(text, free_flyer=ff)
    _same(a, b)
    assert a.is_preorder()
    # the descriptor is one rbd_model_create accepts (no device needed for that)
    cd, _keep = _capi.make_desc(a)
    h = C.c_void_p()
    assert rbd_lib.rbd_model_create(C.byref(cd), C.byref(h)) == 0, rbd_lib.rbd_last_error()
    rbd_lib.rbd_model_destroy(h)


@pytest.mark.parametrize("name,ff", [("simple_arm", False), ("schunk_svh_hand_right_regularized_inertia", False),
                                     ("schunk_svh_hand_right_ff", True)])
def test_committed_descriptors_survive_a_urdf_round_trip(rbd_lib, name, ff, tmp_path):
    """Runs anywhere: the committed descriptor of a reference URDF (tests/golden/*.model.json) is written out as URDF
    text (one link per joint, no fixed joints) and read back by the C++ reader through a FILE: tree, joint types, axes,
    placements and inertias must come back (frames differ: the golden files also hold the merged fixed links)."""
    g = {m.name: m for m in helpers.all_models()}[name]
    lines = ['<?xml version="1.0"?>', f'<robot name="{name}">']
    first = 2 if ff else 1

    def inertial(y):
        m, c = y[0], y[1:4]
        return (f'<inertial><origin xyz="{_fmt(c)}"/><mass value="{float(m)!r}"/><inertia ixx="{float(y[4])!r}" '
                f'ixy="{float(y[5])!r}" iyy="{float(y[6])!r}" ixz="{float(y[7])!r}" iyz="{float(y[8])!r}" izz="{float(y[9])!r}"/></inertial>')

    root = 1 if ff else 0
    lines.append(f'<link name="L{root}">{inertial(g.inertia[root])}</link>')
    kind = {mdl.J_RX: "revolute", mdl.J_RY: "revolute", mdl.J_RZ: "revolute", mdl.J_RU: "revolute",
            mdl.J_PX: "prismatic", mdl.J_PY: "prismatic", mdl.J_PZ: "prismatic", mdl.J_PU: "prismatic",
            mdl.J_RUBX: "continuous", mdl.J_RUBY: "continuous", mdl.J_RUBZ: "continuous", mdl.J_RUBU: "continuous"}
    for j in range(first, g.njoints):
        R = g.placement[j][:9].reshape(3, 3)
        # URDF rpy of R = Rz(y) Ry(p) Rx(r)
        p = -math.asin(max(-1.0, min(1.0, R[2, 0])))
        r = math.atan2(R[2, 1], R[2, 2]); yw = math.atan2(R[1, 0], R[0, 0])
        lines.append(f'<link name="L{j}">{inertial(g.inertia[j])}</link>')
        lines.append(f'<joint name="J{j:04d}" type="{kind[int(g.jtype[j])]}"><parent link="L{int(g.parents[j])}"/>'
                     f'<child link="L{j}"/><origin xyz="{_fmt(g.placement[j][9:])}" rpy="{_fmt((r, p, yw))}"/>'
                     f'<axis xyz="{_fmt(g.axis[j])}"/></joint>')
    lines.append("</robot>")
    path = tmp_path / (name + ".urdf")
    path.write_text("\n".join(lines))
    d = api.load_urdf(str(path), free_flyer=ff)
    for k in ("parents", "jtype", "idx_q", "idx_v"):
        assert np.array_equal(getattr(d, k), getattr(g, k)), k
    assert np.allclose(d.axis, g.axis, atol=1e-15)
    assert np.allclose(d.placement, g.placement, atol=1e-12)      # through rpy and back
    assert np.allclose(d.inertia, g.inertia, atol=1e-15, rtol=1e-13)
    assert d.n_out == d.njoints + d.nframes0 and d.nframes0 == 1 + 2 * (g.njoints - (1 if ff else 0)) + (0 if ff else 0)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference assets only exist in the build container")
def test_cpp_reader_on_the_reference_urdfs(rbd_lib):
    golden = {m.name: m for m in helpers.all_models()}
    for name, path, ff in (("simple_arm", REF + "/JointActuator/simple_arm.urdf", False),
                           ("schunk_svh_hand_right_regularized_inertia",
                            REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", False),
                           ("schunk_svh_hand_right_ff",
                            REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf", True),
                           (None, REF + "/robots_description/schunk_hand/schunk_svh_hand_right.urdf", False)):
        d = api.load_urdf(path, free_flyer=ff)
        _same(d, mdl.from_urdf(path, free_flyer=ff))
        if name:
            _same(d, golden[name])
    # the structural pins of the reference's scenes, through the C handle's name lookups
    h = C.c_void_p()
    p = REF + "/robots_description/schunk_hand/schunk_svh_hand_right_regularized_inertia.urdf"
    assert rbd_lib.rbd_urdf_load_file(p.encode(), 0, C.byref(h)) == 0
    try:
        assert rbd_lib.rbd_urdf_desc(h).contents.nq == 20                      # InverseKineHand.py:14-17
        assert rbd_lib.rbd_urdf_frame_name(h, 33) == b"fftip_joint"            # schunkHandWithConstraints.py:73-79
        assert rbd_lib.rbd_urdf_find_frame(h, b"fftip", mdl.F_BODY) == 34
        assert rbd_lib.rbd_urdf_find_frame(h, b"fftip", mdl.F_JOINT) == -1
        assert rbd_lib.rbd_urdf_find_joint(h, b"universe") == 0
        assert rbd_lib.rbd_urdf_frame_type(h, 33) == mdl.F_FIXED_JOINT
        assert rbd_lib.rbd_urdf_robot_name(h) is not None
    finally:
        rbd_lib.rbd_urdf_destroy(h)


BAD = {
    "not xml": ("hello", -2),
    "unterminated": ('<robot name="r"><link name="a">', -2),
    "mismatched close": ('<robot name="r"><link name="a"></robot></link>', -2),
    "wrong root": ('<model><link name="a"/></model>', -2),
    "no links": ('<robot name="r"/>', -2),
    "two roots": ('<robot name="r"><link name="a"/><link name="b"/></robot>', -2),
    "loop": ('<robot name="r"><link name="a"/><link name="b"/>'
             '<joint name="j" type="fixed"><parent link="a"/><child link="b"/></joint>'
             '<joint name="k" type="fixed"><parent link="b"/><child link="a"/></joint></robot>', -2),
    "unknown link": ('<robot name="r"><link name="a"/><joint name="j" type="fixed"><parent link="a"/><child link="zz"/></joint></robot>', -2),
    "two parents": ('<robot name="r"><link name="a"/><link name="b"/><link name="c"/>'
                    '<joint name="j" type="fixed"><parent link="a"/><child link="c"/></joint>'
                    '<joint name="k" type="fixed"><parent link="b"/><child link="c"/></joint></robot>', -2),
    "bad number": ('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="revolute"><parent link="a"/>'
                   '<child link="b"/><origin xyz="0 0 zero"/></joint></robot>', -2),
    "zero axis": ('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="revolute"><parent link="a"/>'
                  '<child link="b"/><axis xyz="0 0 0"/></joint></robot>', -2),
    "unknown type": ('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="screw"><parent link="a"/>'
                     '<child link="b"/></joint></robot>', -2),
    "planar": ('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="planar"><parent link="a"/>'
               '<child link="b"/></joint></robot>', -5),
    "duplicate link": ('<robot name="r"><link name="a"/><link name="a"/></robot>', -2),
    "mass missing": ('<robot name="r"><link name="a"><inertial><inertia ixx="1"/></inertial></link></robot>', -2),
    "bad entity": ('<robot name="&nbsp;"><link name="a"/></robot>', -2),
}


@pytest.mark.parametrize("case", sorted(BAD))
def test_malformed_urdf_is_an_error_code_not_a_crash(rbd_lib, case):
    text, code = BAD[case]
    h = C.c_void_p(1)
    rc = rbd_lib.rbd_urdf_load_string(text.encode(), 0, C.byref(h))
    assert rc == code, (rc, rbd_lib.rbd_last_error())
    assert not h.value                                   # no handle on failure
    assert rbd_lib.rbd_last_error()                      # and a reason
    with pytest.raises(api.RbdError):
        api.load_urdf(text, is_text=True)


def test_null_arguments_and_missing_file(rbd_lib):
    h = C.c_void_p()
    assert rbd_lib.rbd_urdf_load_string(None, 0, C.byref(h)) == -1
    assert rbd_lib.rbd_urdf_load_file(b"/nonexistent/robot.urdf", 0, C.byref(h)) == -1
    assert rbd_lib.rbd_urdf_desc(None) in (None,) or not rbd_lib.rbd_urdf_desc(None)
    assert rbd_lib.rbd_urdf_nframes0(None) == -1 and rbd_lib.rbd_urdf_joint_name(None, 0) is None
    rbd_lib.rbd_urdf_destroy(None)


def test_single_link_robot_and_near_axis(rbd_lib):
    d = api.load_urdf('<robot name="one"><link name="base"><inertial><mass value="2"/><inertia ixx="1" iyy="1" izz="1"/>'
                      '</inertial></link></robot>', is_text=True)
    assert (d.njoints, d.nq, d.nv, d.nframes0) == (1, 0, 0, 3) and d.inertia[0][0] == 2.0
    f = api.load_urdf('<robot name="one"><link name="base"/></robot>', free_flyer=True, is_text=True)
    assert (f.njoints, f.nq, f.nv) == (2, 7, 6) and f.has_freeflyer_root
    # Eigen isApprox(1e-12): an axis within rounding of e_z is the aligned model, 1e-6 away is not
    t = ('<robot name="r"><link name="a"/><link name="b"/><joint name="j" type="revolute"><parent link="a"/><child link="b"/>'
         '<axis xyz="%s"/></joint></robot>')
    assert int(api.load_urdf(t % "0 1e-14 1", is_text=True).jtype[1]) == mdl.J_RZ
    near = api.load_urdf(t % "0 1e-6 1", is_text=True)
    assert int(near.jtype[1]) == mdl.J_RU and abs(np.linalg.norm(near.axis[1]) - 1.0) < 1e-15
    _same(near, mdl.from_urdf_string(t % "0 1e-6 1"))
    _same(api.load_urdf(t % "0 1e-14 1", is_text=True), mdl.from_urdf_string(t % "0 1e-14 1"))

"""Kinematic-tree model descriptor: the constants the B200 back end stages once per robot.

The reference builds a ``pinocchio::Model`` from a URDF in ``URDFModelLoader::load()``
(reference ``src/sofa/RigidBodyDynamics/URDFModelLoader.cpp:126-184``) and hands it to the mapping with
two frame-index tables (``setBodyCoMFrames`` / ``m_extraFrames``, ``URDFModelLoader.cpp:263-268``).
Everything the hot path needs from that model is a handful of flat arrays; :class:`ModelDescriptor`
is exactly that set, in the layout ``include/rbd_b200.h`` (``rbd_model_desc``) declares.

Three ways to get one:

* :func:`from_urdf` — URDF reader following pinocchio 3.3.1's URDF conventions (children visited in
  joint-name order, fixed joints merged into the parent joint's inertia, JOINT/FIXED_JOINT frame followed
  by BODY frame, exact-axis test for RX/RY/RZ, ``<mimic>`` ignored).  Written from the documented behaviour
  of ``pinocchio::urdf::buildModel`` (the call at ``URDFModelLoader.cpp:132,137``); pinocchio itself is not
  available offline, see DESIGN.md "parity unpinned".
* :func:`panda7`, :func:`solo12`, :func:`talos_reduced` — the three BASELINE robots.  Their URDFs
  (``example-robot-data``) are not available offline, so these are *topology-faithful synthetic* models:
  tree shape, joint types/axes and dof counts as in SURVEY.md App. B, link geometry approximate, inertias
  synthetic (seeded, physically valid).  Every report labels them so.
* :func:`random_tree` — random trees for property tests.

No arithmetic of the hot path lives here; this is load-time host logic.
"""
from __future__ import annotations

import dataclasses
import json
import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# Joint type codes: must match include/rbd_b200.h (RBD_JOINT_*).
J_UNIVERSE = 0
J_RX, J_RY, J_RZ, J_RU = 1, 2, 3, 4           # revolute about X/Y/Z / unaligned unit axis (nq=nv=1)
J_PX, J_PY, J_PZ, J_PU = 5, 6, 7, 8           # prismatic (nq=nv=1)
J_RUBX, J_RUBY, J_RUBZ, J_RUBU = 9, 10, 11, 12  # unbounded revolute ("continuous"): q=(cos,sin), nq=2, nv=1
J_FF = 13                                      # free-flyer: q=[p(3), quat xyzw(4)], nq=7, nv=6 (body-frame twist)

JOINT_NQ = {J_UNIVERSE: 0, J_RX: 1, J_RY: 1, J_RZ: 1, J_RU: 1, J_PX: 1, J_PY: 1, J_PZ: 1, J_PU: 1,
            J_RUBX: 2, J_RUBY: 2, J_RUBZ: 2, J_RUBU: 2, J_FF: 7}
JOINT_NV = {J_UNIVERSE: 0, J_RX: 1, J_RY: 1, J_RZ: 1, J_RU: 1, J_PX: 1, J_PY: 1, J_PZ: 1, J_PU: 1,
            J_RUBX: 1, J_RUBY: 1, J_RUBZ: 1, J_RUBU: 1, J_FF: 6}
JOINT_NAMES = {J_UNIVERSE: "universe", J_RX: "RX", J_RY: "RY", J_RZ: "RZ", J_RU: "RU", J_PX: "PX", J_PY: "PY",
               J_PZ: "PZ", J_PU: "PU", J_RUBX: "RUBX", J_RUBY: "RUBY", J_RUBZ: "RUBZ", J_RUBU: "RUBU", J_FF: "FF"}

# frame types (informational only; pinocchio FrameType)
F_OP, F_JOINT, F_FIXED_JOINT, F_BODY = 1, 2, 4, 8


def _se3(R=None, p=None) -> np.ndarray:
    """SE3 as 12 doubles: R row-major (9) then p (3)."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
    p = np.zeros(3) if p is None else np.asarray(p, dtype=np.float64).reshape(3)
    return np.concatenate([R.reshape(9), p])


def se3_R(x: np.ndarray) -> np.ndarray:
    return np.asarray(x[:9]).reshape(3, 3)


def se3_p(x: np.ndarray) -> np.ndarray:
    return np.asarray(x[9:12])


def se3_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    Ra, pa, Rb, pb = se3_R(a), se3_p(a), se3_R(b), se3_p(b)
    return _se3(Ra @ Rb, Ra @ pb + pa)


def rpy_to_R(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return Rz @ Ry @ Rx


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


# inertia record: [mass, cx, cy, cz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz]  (pinocchio Symmetric3 storage order)
def inertia_pack(m: float, c: Sequence[float], I: np.ndarray) -> np.ndarray:
    I = np.asarray(I, dtype=np.float64)
    return np.array([m, c[0], c[1], c[2], I[0, 0], I[0, 1], I[1, 1], I[0, 2], I[1, 2], I[2, 2]], dtype=np.float64)


def inertia_unpack(y: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    m = float(y[0]); c = np.array(y[1:4], dtype=np.float64)
    I = np.array([[y[4], y[5], y[7]], [y[5], y[6], y[8]], [y[7], y[8], y[9]]], dtype=np.float64)
    return m, c, I


def inertia_add(ya: np.ndarray, yb: np.ndarray) -> np.ndarray:
    """pinocchio ``InertiaTpl::__plus__`` (mass floor at machine epsilon)."""
    ma, ca, Ia = inertia_unpack(ya)
    mb, cb, Ib = inertia_unpack(yb)
    mab = ma + mb
    inv = 1.0 / max(mab, np.finfo(np.float64).eps)
    ab = ca - cb
    S = _skew(ab)
    return inertia_pack(mab, (ma * ca + mb * cb) * inv, Ia + Ib - (ma * mb * inv) * (S @ S))


def inertia_se3_act(M: np.ndarray, y: np.ndarray) -> np.ndarray:
    """pinocchio ``InertiaTpl::se3Action_impl``: express inertia given in frame b in frame a (M = aMb)."""
    m, c, I = inertia_unpack(y)
    R, p = se3_R(M), se3_p(M)
    return inertia_pack(m, p + R @ c, R @ I @ R.T)


@dataclasses.dataclass
class ModelDescriptor:
    """Flat tree constants (all float64 / int32, C-contiguous).  Field meanings = ``rbd_model_desc``."""
    name: str
    parents: np.ndarray          # [njoints] int32, parents[0] = 0, parents[i] < i
    jtype: np.ndarray            # [njoints] int32
    axis: np.ndarray             # [njoints,3] float64 (unit axis for *U types; e_k for aligned; 0 otherwise)
    placement: np.ndarray        # [njoints,12] float64 jointPlacements
    idx_q: np.ndarray            # [njoints] int32
    idx_v: np.ndarray            # [njoints] int32
    inertia: np.ndarray          # [njoints,10] float64
    frame_parent: np.ndarray     # [nframes] int32 (parent joint)
    frame_placement: np.ndarray  # [nframes,12] float64
    gravity: np.ndarray          # [3] float64, linear part of model.gravity
    out_frames: np.ndarray       # [n_out] int32 = bodyCoMFrames ++ extraFrames (KinematicChainMapping.inl:382-395)
    nframes0: int                # frame count before the CoM frames were appended (URDFModelLoader.cpp:161-165)
    joint_names: List[str] = dataclasses.field(default_factory=list)
    frame_names: List[str] = dataclasses.field(default_factory=list)
    frame_types: List[int] = dataclasses.field(default_factory=list)
    note: str = ""

    # ---- derived sizes -------------------------------------------------------------------------------
    @property
    def njoints(self) -> int:
        return int(self.parents.shape[0])

    @property
    def nq(self) -> int:
        return int(sum(JOINT_NQ[int(t)] for t in self.jtype))

    @property
    def nv(self) -> int:
        return int(sum(JOINT_NV[int(t)] for t in self.jtype))

    @property
    def nframes(self) -> int:
        return int(self.frame_parent.shape[0])

    @property
    def n_out(self) -> int:
        return int(self.out_frames.shape[0])

    @property
    def has_freeflyer_root(self) -> bool:
        return self.njoints > 1 and int(self.jtype[1]) == J_FF and int(self.parents[1]) == 0

    @property
    def nq_joints(self) -> int:
        """Size of SOFA input1 (joint dofs): nq, or nq-7 with a free-flyer root (URDFModelLoader.cpp:200)."""
        return self.nq - 7 if self.has_freeflyer_root else self.nq

    @property
    def nv_joints(self) -> int:
        return self.nv - 6 if self.has_freeflyer_root else self.nv

    def eval_bytes(self) -> int:
        """Algorithmic bytes per eval, SURVEY.md §8(d)."""
        nq, nv, nj = self.nq, self.nv, self.njoints
        return 8 * ((nq + 2 * nv) + 12 * (nj - 1) + 6 * nv + nv * (nv + 1) // 2 + nv)

    # ---- tree helpers (host logic, also used by tests) --------------------------------------------------
    def nv_joint(self, i: int) -> int:
        return JOINT_NV[int(self.jtype[i])]

    def nq_joint(self, i: int) -> int:
        return JOINT_NQ[int(self.jtype[i])]

    def depth(self) -> np.ndarray:
        d = np.zeros(self.njoints, dtype=np.int32)
        for i in range(1, self.njoints):
            d[i] = d[self.parents[i]] + 1
        return d

    def subtree_size(self) -> np.ndarray:
        s = np.ones(self.njoints, dtype=np.int32)
        for i in range(self.njoints - 1, 0, -1):
            s[self.parents[i]] += s[i]
        return s

    def is_preorder(self) -> bool:
        """True when every subtree occupies a contiguous index range (pinocchio's URDF order is DFS pre-order).
        The CUDA sweeps rely on this; :func:`validate` rejects anything else."""
        for i in range(2, self.njoints):
            p = int(self.parents[i])
            k = i - 1                      # parent must be i-1 or one of its ancestors
            while k != p and k != 0:
                k = int(self.parents[k])
            if k != p:
                return False
        return True

    def validate(self) -> None:
        nj = self.njoints
        assert nj >= 1 and self.parents[0] == 0 and self.jtype[0] == J_UNIVERSE
        iq = iv = 0
        for i in range(1, nj):
            assert 0 <= self.parents[i] < i, "parents[i] < i required"
            assert int(self.jtype[i]) in JOINT_NQ and int(self.jtype[i]) != J_UNIVERSE
            assert self.idx_q[i] == iq and self.idx_v[i] == iv, "idx_q/idx_v must be cumulative"
            iq += self.nq_joint(i); iv += self.nv_joint(i)
        assert self.is_preorder(), "joints must be numbered in DFS pre-order"
        assert self.frame_parent.min(initial=0) >= 0 and self.frame_parent.max(initial=0) < nj
        assert self.out_frames.min(initial=0) >= 0 and self.out_frames.max(initial=0) < self.nframes
        for a in (self.axis, self.placement, self.inertia, self.frame_placement, self.gravity):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        for a in (self.parents, self.jtype, self.idx_q, self.idx_v, self.frame_parent, self.out_frames):
            assert a.dtype == np.int32 and a.flags.c_contiguous

    # ---- (de)serialisation so fixtures travel without the reference tree --------------------------------
    def to_json(self) -> str:
        d = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            d[f.name] = v.tolist() if isinstance(v, np.ndarray) else v
        return json.dumps(d)

    @staticmethod
    def from_json(s: str) -> "ModelDescriptor":
        d = json.loads(s)
        i32 = ("parents", "jtype", "idx_q", "idx_v", "frame_parent", "out_frames")
        f64 = {"axis": (-1, 3), "placement": (-1, 12), "inertia": (-1, 10), "frame_placement": (-1, 12),
               "gravity": (3,)}
        for k in i32:
            d[k] = np.ascontiguousarray(np.array(d[k], dtype=np.int32).reshape(-1))
        for k, shp in f64.items():
            d[k] = np.ascontiguousarray(np.array(d[k], dtype=np.float64).reshape(shp))
        return ModelDescriptor(**d)


# =========================================================================================================
# Builder that mimics pinocchio::Model::addJoint / addFrame / appendBodyToJoint bookkeeping
# =========================================================================================================
class _Builder:
    def __init__(self, name: str):
        self.name = name
        self.parents = [0]; self.jtype = [J_UNIVERSE]; self.axis = [np.zeros(3)]
        self.placement = [_se3()]; self.inertia = [np.zeros(10)]
        self.joint_names = ["universe"]
        # Model() creates frame 0 "universe" (FIXED_JOINT, parent joint 0)
        self.frame_parent = [0]; self.frame_placement = [_se3()]; self.frame_names = ["universe"]
        self.frame_types = [F_FIXED_JOINT]

    def add_joint(self, parent: int, jtype: int, placement: np.ndarray, name: str, axis=None) -> int:
        if axis is None:
            axis = {J_RX: (1, 0, 0), J_RY: (0, 1, 0), J_RZ: (0, 0, 1), J_PX: (1, 0, 0), J_PY: (0, 1, 0),
                    J_PZ: (0, 0, 1), J_RUBX: (1, 0, 0), J_RUBY: (0, 1, 0), J_RUBZ: (0, 0, 1)}.get(jtype, (0, 0, 0))
        self.parents.append(parent); self.jtype.append(jtype); self.axis.append(np.asarray(axis, dtype=np.float64))
        self.placement.append(np.asarray(placement, dtype=np.float64)); self.inertia.append(np.zeros(10))
        self.joint_names.append(name)
        return len(self.parents) - 1

    def add_frame(self, name: str, parent_joint: int, placement: np.ndarray, ftype: int) -> int:
        self.frame_parent.append(parent_joint); self.frame_placement.append(np.asarray(placement, dtype=np.float64))
        self.frame_names.append(name); self.frame_types.append(ftype)
        return len(self.frame_parent) - 1

    def append_body(self, joint: int, y: np.ndarray, placement: np.ndarray) -> None:
        """pinocchio ``Model::appendBodyToJoint``: inertias[joint] += placement.act(Y)."""
        self.inertia[joint] = inertia_add(self.inertia[joint], inertia_se3_act(placement, y))

    def finish(self, gravity=(0.0, 0.0, -9.81), note: str = "") -> ModelDescriptor:
        nj = len(self.parents)
        idx_q, idx_v, iq, iv = [0], [0], 0, 0
        for i in range(1, nj):
            idx_q.append(iq); idx_v.append(iv)
            iq += JOINT_NQ[self.jtype[i]]; iv += JOINT_NV[self.jtype[i]]
        nframes0 = len(self.frame_parent)
        # URDFModelLoader.cpp:161-165: every existing frame becomes an "extra frame" ...
        extra = list(range(nframes0))
        # ... URDFModelLoader.cpp:177-183: then one OP_FRAME per joint (universe included, kSkipUniverse = 0,
        # Types.h:32) at the joint's CoM lever with identity rotation.
        com = []
        for j in range(nj):
            com.append(self.add_frame(f"Body_{j}_CoM", j, _se3(None, self.inertia[j][1:4]), F_OP))
        d = ModelDescriptor(
            name=self.name,
            parents=np.array(self.parents, dtype=np.int32), jtype=np.array(self.jtype, dtype=np.int32),
            axis=np.ascontiguousarray(np.array(self.axis, dtype=np.float64).reshape(nj, 3)),
            placement=np.ascontiguousarray(np.array(self.placement, dtype=np.float64).reshape(nj, 12)),
            idx_q=np.array(idx_q, dtype=np.int32), idx_v=np.array(idx_v, dtype=np.int32),
            inertia=np.ascontiguousarray(np.array(self.inertia, dtype=np.float64).reshape(nj, 10)),
            frame_parent=np.array(self.frame_parent, dtype=np.int32),
            frame_placement=np.ascontiguousarray(np.array(self.frame_placement, dtype=np.float64).reshape(-1, 12)),
            gravity=np.array(gravity, dtype=np.float64),
            out_frames=np.array(com + extra, dtype=np.int32), nframes0=nframes0,
            joint_names=list(self.joint_names), frame_names=list(self.frame_names),
            frame_types=list(self.frame_types), note=note)
        d.validate()
        return d


# =========================================================================================================
# URDF reader (pinocchio::urdf::buildModel conventions)
# =========================================================================================================
def _floats(s: Optional[str], n: int, default: float = 0.0) -> List[float]:
    if s is None:
        return [default] * n
    v = [float(x) for x in s.split()]
    assert len(v) == n, f"expected {n} floats, got {s!r}"
    return v


def _origin(el) -> np.ndarray:
    if el is None:
        return _se3()
    xyz = _floats(el.get("xyz"), 3)
    rpy = _floats(el.get("rpy"), 3)
    return _se3(rpy_to_R(*rpy), xyz)


def _link_inertia(link_el) -> np.ndarray:
    """URDF <inertial> -> pinocchio Inertia(mass, com, R I R^T) (pinocchio ``convertFromUrdf``)."""
    ine = link_el.find("inertial")
    if ine is None:
        return np.zeros(10)
    o = _origin(ine.find("origin"))
    m = float(ine.find("mass").get("value"))
    i = ine.find("inertia")
    g = lambda k: float(i.get(k, "0"))
    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
    R = se3_R(o)
    return inertia_pack(m, se3_p(o), R @ I @ R.T)


def from_urdf_string(text: str, free_flyer: bool = False, name: Optional[str] = None) -> ModelDescriptor:
    """URDF text -> descriptor.  ``free_flyer`` = the loader's ``useFreeFlyerRootJoint`` (URDFModelLoader.cpp:130-139).

    Leading whitespace is stripped (``simple_arm.urdf`` starts with a blank, which strict XML parsers reject)."""
    root = ET.fromstring(text.lstrip())
    assert root.tag == "robot"
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {j.get("name"): j for j in root.findall("joint")}
    child_of: Dict[str, List[str]] = {n: [] for n in links}   # parent link -> [joint names]
    has_parent = set()
    # urdfdom keeps joints in a std::map (sorted by name) and fills child lists in that order
    for jn in sorted(joints):
        j = joints[jn]
        child_of[j.find("parent").get("link")].append(jn)
        has_parent.add(j.find("child").get("link"))
    roots = [n for n in links if n not in has_parent]
    assert len(roots) == 1, f"URDF must have exactly one root link, got {roots}"
    root_link = roots[0]

    b = _Builder(name or root.get("name", "robot"))
    body_frame: Dict[str, int] = {}
    Y = _link_inertia(links[root_link])
    if free_flyer:
        # UrdfVisitorWithRootJoint::addRootJoint: addJoint(0, FF, I, "root_joint"); addJointFrame; appendBody; addBodyFrame
        jid = b.add_joint(0, J_FF, _se3(), "root_joint")
        fid = b.add_frame("root_joint", jid, _se3(), F_JOINT)
        b.append_body(jid, Y, _se3())
        body_frame[root_link] = b.add_frame(root_link, jid, _se3(), F_BODY)
    else:
        # UrdfVisitor::addRootJoint -> addFixedJointAndBody(0, I, "root_joint", Y, body)
        fid = b.add_frame("root_joint", 0, _se3(), F_FIXED_JOINT)
        b.append_body(0, Y, _se3())
        body_frame[root_link] = b.add_frame(root_link, 0, _se3(), F_BODY)

    def classify(kind: str, axis: np.ndarray) -> Tuple[int, np.ndarray]:
        # pinocchio's extractCartesianAxis: axis.isApprox(Unit{X,Y,Z}) at Eigen's default precision 1e-12
        n2 = float(axis @ axis)
        ex = tuple(float((axis - e) @ (axis - e)) <= 1e-24 * min(n2, 1.0) for e in np.eye(3))
        table = {"revolute": (J_RX, J_RY, J_RZ, J_RU), "continuous": (J_RUBX, J_RUBY, J_RUBZ, J_RUBU),
                 "prismatic": (J_PX, J_PY, J_PZ, J_PU)}[kind]
        for k in range(3):
            if ex[k]:
                return table[k], np.eye(3)[k].copy()
        return table[3], axis / math.sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2])

    def visit(link: str) -> None:
        for jn in child_of[link]:
            j = joints[jn]
            child = j.find("child").get("link")
            kind = j.get("type")
            pl = _origin(j.find("origin"))
            Yc = _link_inertia(links[child])
            pf = body_frame[link]
            pj, ppl = b.frame_parent[pf], b.frame_placement[pf]
            if kind in ("revolute", "continuous", "prismatic"):
                ax_el = j.find("axis")
                axis = np.array(_floats(ax_el.get("xyz") if ax_el is not None else "1 0 0", 3))
                jt, ax = classify(kind, axis)
                jid = b.add_joint(pj, jt, se3_mul(ppl, pl), jn, ax)
                b.add_frame(jn, jid, _se3(), F_JOINT)
                b.append_body(jid, Yc, _se3())
                body_frame[child] = b.add_frame(child, jid, _se3(), F_BODY)
            elif kind == "floating":
                jid = b.add_joint(pj, J_FF, se3_mul(ppl, pl), jn)
                b.add_frame(jn, jid, _se3(), F_JOINT)
                b.append_body(jid, Yc, _se3())
                body_frame[child] = b.add_frame(child, jid, _se3(), F_BODY)
            elif kind == "fixed":
                fpl = se3_mul(ppl, pl)
                b.add_frame(jn, pj, fpl, F_FIXED_JOINT)
                b.append_body(pj, Yc, fpl)
                body_frame[child] = b.add_frame(child, pj, fpl, F_BODY)
            else:
                raise ValueError(f"unsupported URDF joint type {kind!r} ({jn})")
            visit(child)

    visit(root_link)
    return b.finish(note="from URDF, pinocchio 3.3.1 URDF conventions (restated; unpinned)")


def from_urdf(path: str, free_flyer: bool = False) -> ModelDescriptor:
    with open(path, "r") as f:
        return from_urdf_string(f.read(), free_flyer=free_flyer)


# =========================================================================================================
# Synthetic BASELINE robots (topology-faithful, geometry approximate, inertias synthetic)
# =========================================================================================================
def _synthetic_inertia(rng: np.random.Generator, mass_scale: float = 1.0) -> np.ndarray:
    m = float(rng.uniform(0.3, 4.0) * mass_scale)
    c = rng.uniform(-0.06, 0.06, 3)
    A = rng.uniform(-1.0, 1.0, (3, 3))
    I = (A @ A.T + 0.5 * np.eye(3)) * (m * 4e-3)
    return inertia_pack(m, c, I)


def _chain(b: _Builder, rng, parent_joint: int, parent_pl: np.ndarray, specs, prefix: str = "",
           mass_scale: float = 1.0) -> int:
    """Append a serial chain.  specs = [(name, jtype, xyz, rpy[, axis])]; every joint gets JOINT + BODY frames."""
    pj = parent_joint
    pl_acc = parent_pl
    for s in specs:
        name, jt, xyz, rpy = s[0], s[1], s[2], s[3]
        axis = s[4] if len(s) > 4 else None
        jid = b.add_joint(pj, jt, se3_mul(pl_acc, _se3(rpy_to_R(*rpy), xyz)), prefix + name, axis)
        b.add_frame(prefix + name, jid, _se3(), F_JOINT)
        b.append_body(jid, _synthetic_inertia(rng, mass_scale), _se3())
        b.add_frame(prefix + name.replace("_joint", "") + "_link", jid, _se3(), F_BODY)
        pj = jid
        pl_acc = _se3()
    return pj


_NOTE = ("synthetic: topology/joint types per SURVEY.md App. B (URDF not available offline); "
         "link geometry approximate; inertias synthetic (seeded)")


def panda7() -> ModelDescriptor:
    """Franka Panda, 7 revolute Z joints, fixed base (BASELINE config 'Panda 7-dof')."""
    rng = np.random.default_rng(0x7A0DA)
    b = _Builder("panda7")
    b.add_frame("root_joint", 0, _se3(), F_FIXED_JOINT)
    b.append_body(0, _synthetic_inertia(rng), _se3())
    b.add_frame("panda_link0", 0, _se3(), F_BODY)
    h = math.pi / 2
    specs = [("panda_joint1", J_RZ, (0, 0, 0.333), (0, 0, 0)),
             ("panda_joint2", J_RZ, (0, 0, 0), (-h, 0, 0)),
             ("panda_joint3", J_RZ, (0, -0.316, 0), (h, 0, 0)),
             ("panda_joint4", J_RZ, (0.0825, 0, 0), (h, 0, 0)),
             ("panda_joint5", J_RZ, (-0.0825, 0.384, 0), (-h, 0, 0)),
             ("panda_joint6", J_RZ, (0, 0, 0), (h, 0, 0)),
             ("panda_joint7", J_RZ, (0.088, 0, 0), (h, 0, 0))]
    last = _chain(b, rng, 0, _se3(), specs)
    fl = _se3(None, (0, 0, 0.107))
    b.add_frame("panda_joint8", last, fl, F_FIXED_JOINT)
    b.add_frame("panda_link8", last, fl, F_BODY)
    return b.finish(note=_NOTE)


def solo12() -> ModelDescriptor:
    """Solo-12 quadruped: free-flyer + 4 legs x (HAA about X, HFE about Y, KFE about Y)."""
    rng = np.random.default_rng(0x50102)
    b = _Builder("solo12")
    ff = b.add_joint(0, J_FF, _se3(), "root_joint")
    b.add_frame("root_joint", ff, _se3(), F_JOINT)
    b.append_body(ff, _synthetic_inertia(rng, 0.5), _se3())
    b.add_frame("base_link", ff, _se3(), F_BODY)
    for leg, sx, sy in (("FL", 1, 1), ("FR", 1, -1), ("HL", -1, 1), ("HR", -1, -1)):
        specs = [(f"{leg}_HAA", J_RX, (sx * 0.1946, sy * 0.0875, 0), (0, 0, 0)),
                 (f"{leg}_HFE", J_RY, (0, sy * 0.014, 0), (0, 0, 0)),
                 (f"{leg}_KFE", J_RY, (0, sy * 0.03745, -0.16), (0, 0, 0))]
        last = _chain(b, rng, ff, _se3(), specs, mass_scale=0.1)
        foot = _se3(None, (0, sy * 0.008, -0.16))
        b.add_frame(f"{leg}_ANKLE", last, foot, F_FIXED_JOINT)
        b.add_frame(f"{leg}_FOOT", last, foot, F_BODY)
    return b.finish(note=_NOTE)


def talos_reduced() -> ModelDescriptor:
    """Talos reduced humanoid: free-flyer + 2x6 legs + 2 torso + 2x(7 arm + gripper) + 2 head = 34 joints,
    nq=39, nv=38.  Children visited in joint-name order as pinocchio's URDF parser does."""
    rng = np.random.default_rng(0x7A105)
    b = _Builder("talos_reduced")
    ff = b.add_joint(0, J_FF, _se3(), "root_joint")
    b.add_frame("root_joint", ff, _se3(), F_JOINT)
    b.append_body(ff, _synthetic_inertia(rng, 4.0), _se3())
    b.add_frame("base_link", ff, _se3(), F_BODY)
    for side, sy in (("left", 1), ("right", -1)):
        specs = [(f"leg_{side}_1_joint", J_RZ, (-0.02, sy * 0.085, -0.27105), (0, 0, 0)),
                 (f"leg_{side}_2_joint", J_RX, (0, 0, 0), (0, 0, 0)),
                 (f"leg_{side}_3_joint", J_RY, (0, 0, 0), (0, 0, 0)),
                 (f"leg_{side}_4_joint", J_RY, (0, 0, -0.38), (0, 0, 0)),
                 (f"leg_{side}_5_joint", J_RY, (0, 0, -0.325), (0, 0, 0)),
                 (f"leg_{side}_6_joint", J_RX, (0, 0, 0), (0, 0, 0))]
        last = _chain(b, rng, ff, _se3(), specs, mass_scale=1.5)
        sole = _se3(None, (0, 0, -0.107))
        b.add_frame(f"leg_{side}_sole_fix_joint", last, sole, F_FIXED_JOINT)
        b.add_frame(f"{side}_sole_link", last, sole, F_BODY)
    torso = _chain(b, rng, ff, _se3(), [("torso_1_joint", J_RZ, (0, 0, 0.0722), (0, 0, 0)),
                                        ("torso_2_joint", J_RY, (0, 0, 0), (0, 0, 0))], mass_scale=3.0)
    for side, sy in (("left", 1), ("right", -1)):
        specs = [(f"arm_{side}_1_joint", J_RZ, (0, sy * 0.1575, 0.232), (0, 0, 0)),
                 (f"arm_{side}_2_joint", J_RX, (0.00493, sy * 0.1365, 0.04673), (0, 0, 0)),
                 (f"arm_{side}_3_joint", J_RZ, (0, 0, 0), (0, 0, 0)),
                 (f"arm_{side}_4_joint", J_RY, (0.02, 0, -0.273), (0, 0, 0)),
                 (f"arm_{side}_5_joint", J_RZ, (-0.02, 0, -0.2643), (0, 0, 0)),
                 (f"arm_{side}_6_joint", J_RX, (0, 0, 0), (0, 0, 0)),
                 (f"arm_{side}_7_joint", J_RY, (0, 0, 0), (0, 0, 0)),
                 (f"gripper_{side}_joint", J_RY, (0, 0, -0.12), (0, 0, sy * 0.3))]
        _chain(b, rng, torso, _se3(), specs, mass_scale=0.6)
    _chain(b, rng, torso, _se3(), [("head_1_joint", J_RY, (0, 0, 0.316), (0, 0, 0)),
                                   ("head_2_joint", J_RZ, (0.039, 0, 0.048), (0, 0, 0))], mass_scale=0.4)
    d = b.finish(note=_NOTE)
    assert d.njoints == 34 and d.nq == 39 and d.nv == 38
    return d


def random_tree(seed: int, nmoving: int = 8, free_flyer: bool = False, all_types: bool = True,
                branch_prob: float = 0.35, fixed_frames: int = 3) -> ModelDescriptor:
    """Random pre-order tree for property tests: random placements/axes/inertias and joint types."""
    rng = np.random.default_rng(seed)
    b = _Builder(f"random_{seed}")
    b.add_frame("root_joint", 0, _se3(), F_FIXED_JOINT)
    b.append_body(0, _synthetic_inertia(rng), _se3())
    types = ([J_RX, J_RY, J_RZ, J_RU, J_PX, J_PY, J_PZ, J_PU, J_RUBX, J_RUBY, J_RUBZ, J_RUBU] if all_types
             else [J_RX, J_RY, J_RZ])

    def rand_pl():
        return _se3(rpy_to_R(*rng.uniform(-math.pi, math.pi, 3)), rng.uniform(-0.3, 0.3, 3))

    # build a random parent array, then renumber in DFS pre-order
    par = [0]
    if free_flyer:
        par.append(0)
    while len(par) < nmoving + 1:
        n = len(par)
        if n == 1:
            par.append(0)
        elif rng.uniform() < branch_prob:
            par.append(int(rng.integers(0 if not free_flyer else 1, n)))
        else:
            par.append(n - 1)
    children: Dict[int, List[int]] = {i: [] for i in range(len(par))}
    for i in range(1, len(par)):
        children[par[i]].append(i)
    order: List[Tuple[int, int]] = []   # (old index, old parent)

    def dfs(i):
        for c in children[i]:
            order.append((c, i)); dfs(c)
    dfs(0)
    new_of = {0: 0}
    for old, oldp in order:
        is_ff = free_flyer and old == 1
        jt = J_FF if is_ff else int(rng.choice(types))
        ax = None
        if jt in (J_RU, J_PU, J_RUBU):
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        jid = b.add_joint(new_of[oldp], jt, rand_pl(), f"j{old}", ax)
        new_of[old] = jid
        b.add_frame(f"j{old}", jid, _se3(), F_JOINT)
        b.append_body(jid, _synthetic_inertia(rng), _se3())
        b.add_frame(f"b{old}", jid, _se3(), F_BODY)
    for k in range(fixed_frames):
        pj = int(rng.integers(0, len(b.parents)))
        b.add_frame(f"fix{k}", pj, rand_pl(), F_FIXED_JOINT)
    return b.finish(gravity=(0.0, 0.0, -9.81), note="random tree for property tests")


ROBOTS = {"panda7": panda7, "solo12": solo12, "talos_reduced": talos_reduced}


def neutral_q(d: ModelDescriptor) -> np.ndarray:
    """pinocchio::neutral: zeros, (1,0) for unbounded revolute, unit quaternion for free-flyer."""
    q = np.zeros(d.nq)
    for i in range(1, d.njoints):
        t = int(d.jtype[i]); o = int(d.idx_q[i])
        if t in (J_RUBX, J_RUBY, J_RUBZ, J_RUBU):
            q[o] = 1.0
        elif t == J_FF:
            q[o + 6] = 1.0
    return q


def random_state(d: ModelDescriptor, n: int, seed: int = 0x5EED):
    """Synthetic inputs of SURVEY.md §8(d): q ~ U(-pi,pi) (FF: p ~ U(-1,1)^3, quat = normalised N(0,1)^4 xyzw;
    unbounded revolute: (cos,sin) of U(-pi,pi)); v,a ~ U(-1,1).  Returns instance-major arrays [n,nq],[n,nv],[n,nv]."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-math.pi, math.pi, (n, d.nq))
    for i in range(1, d.njoints):
        t = int(d.jtype[i]); o = int(d.idx_q[i])
        if t == J_FF:
            q[:, o:o + 3] = rng.uniform(-1, 1, (n, 3))
            w = rng.normal(size=(n, 4)); w /= np.linalg.norm(w, axis=1, keepdims=True)
            q[:, o + 3:o + 7] = w
        elif t in (J_RUBX, J_RUBY, J_RUBZ, J_RUBU):
            th = rng.uniform(-math.pi, math.pi, n)
            q[:, o] = np.cos(th); q[:, o + 1] = np.sin(th)
        elif t in (J_PX, J_PY, J_PZ, J_PU):
            q[:, o] = rng.uniform(-0.5, 0.5, n)
    v = rng.uniform(-1, 1, (n, d.nv))
    a = rng.uniform(-1, 1, (n, d.nv))
    return np.ascontiguousarray(q), np.ascontiguousarray(v), np.ascontiguousarray(a)

"""Host-side mirror of KinematicChainMapping: the component-level behaviour that needs no arithmetic
(reference KinematicChainMapping.inl:105-295 arity checks -> ComponentState::Invalid -> silent early return;
applyDJT no-op and getJ() == nullptr, KinematicChainMapping.h:73-83)."""
import numpy as np

from sofa.rigidbodydynamics_b200 import model as mdl
from sofa.rigidbodydynamics_b200.mapping import INVALID, VALID, KinematicChainMapping


def test_arity_mismatch_marks_component_invalid(rbd_lib):
    d = mdl.panda7()
    kcm = KinematicChainMapping(d, n_instances=2)
    assert kcm.d_componentState == VALID
    out = np.zeros((2, d.n_out, 7)); in1 = np.zeros((2, d.nq))
    kcm.apply([out, out], [in1], [])                    # two output vectors: msg_error + Invalid (inl:114-119)
    assert kcm.d_componentState == INVALID and kcm.msgs
    # every later entry point returns silently without touching its outputs (inl:346,406,458,523)
    out[...] = 7.0
    kcm.apply([out], [in1], [])
    kcm.applyJ([out[..., :6]], [in1], [])
    kcm.applyJT([in1], [], [out[..., :6]])
    assert np.all(out == 7.0)
    assert kcm.applyJT_matrix({(0, 0): [(1, np.ones(6))]}) == ({}, {})


def test_empty_vectors_return_silently_without_invalidating(rbd_lib):
    """The reference returns early on EMPTY vector lists and leaves the component state alone
    (KinematicChainMapping.inl:117-118,165-166,212-213); only size > 1 invalidates."""
    d = mdl.panda7()
    kcm = KinematicChainMapping(d, n_instances=1)
    out = np.zeros((1, d.n_out, 7)); in1 = np.zeros((1, d.nq))
    kcm.apply([], [in1], [])
    kcm.apply([out], [], [])
    kcm.applyJ([], [in1], [])
    kcm.applyJT([], [], [out[..., :6]])
    kcm.applyJT([in1], [], [])
    assert kcm.d_componentState == VALID and not kcm.msgs
    kcm.applyJT([in1], [], [out[..., :6], out[..., :6]])   # inForce.size() != 1 -> Invalid (inl:215-220)
    assert kcm.d_componentState == INVALID


def test_applyDJT_and_getJ(rbd_lib):
    kcm = KinematicChainMapping(mdl.solo12(), n_instances=1)
    assert kcm.applyDJT(None, None, None) is None       # explicit no-op, KinematicChainMapping.h:73-80
    assert kcm.getJ() is None                           # KinematicChainMapping.h:83


def test_root_index_defaults_to_last(rbd_lib):
    kcm = KinematicChainMapping(mdl.solo12(), n_instances=1)
    roots = np.zeros((1, 3, 7))
    assert kcm._root_index(roots) == 2                  # checkIndexFromRoot, inl:319-336
    kcm2 = KinematicChainMapping(mdl.solo12(), n_instances=1, index_input2=1)
    assert kcm2._root_index(roots) == 1
    kcm3 = KinematicChainMapping(mdl.solo12(), n_instances=1, index_input2=9)
    assert kcm3._root_index(roots) == 2

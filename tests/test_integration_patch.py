"""integration/: the reference-side binding as FILES.
  * apply_backend_patch.py splices KinematicChainMapping.backend.inl into a scratch copy of the reference's
    KinematicChainMapping.{h,inl}: every edit must match exactly once, the pinocchio algorithm calls must be gone, the
    C-ABI calls must be there, braces must balance.  Needs /root/reference (this container); skipped elsewhere.
  * rbd_backend_glue.h — everything SOFA-free the patched component uses — compiles standalone with g++ against
    include/rbd_b200.h, flattens a stand-in model through the same template the component instantiates with
    pinocchio::Model, and creates it through the C ABI."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
INTEG = os.path.join(ROOT, "integration")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference tree is not present on this box")
def test_patch_applies_to_the_reference(tmp_path):
    out = tmp_path / "patched"
    r = subprocess.run([sys.executable, os.path.join(INTEG, "apply_backend_patch.py"), REF, str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    inl = (out / "KinematicChainMapping.inl").read_text()
    hdr = (out / "KinematicChainMapping.h").read_text()
    ref_inl = open(os.path.join(REF, "src/sofa/RigidBodyDynamics/KinematicChainMapping.inl")).read()
    # the pinocchio algorithm calls of the hot path are gone (reference: inl:373,375,434,444,480,490,579) ...
    for call in ("computeJointJacobians", "updateFramePlacements", "getFrameJacobian", "make_shared<pinocchio::Data>"):
        assert call in ref_inl and call not in inl, call
    # ... replaced by one C-ABI call per worker, the staging at init and the helpers the snippets of round 1 lacked
    for call in ("rbd_apply_host(", "rbd_applyJ_host(", "rbd_applyJT_host(", "m_rbdRows.run(", "stageBackend();",
                 "::stageBackend()", "::fail(const char", "rbd_glue::fillDescriptor(", "d_robots(initData("):
        assert inl.count(call) == 1, call
    assert "rbd_glue::Backend m_rbd;" in hdr and "rbd_backend_glue.h" in hdr
    # untouched: the public overloads' arity checks and everything outside the hot path
    for keep in ("only supports output vector of size 1", "checkIndexFromRoot()", "kMinTorqueSqrd"):
        assert keep in inl
    # structure: braces and parentheses balance after the splice (comments and literals stripped)
    code = re.sub(r"//[^\n]*|/\*.*?\*/|\"(?:\\.|[^\"\\])*\"", "", inl, flags=re.S)
    assert code.count("{") == code.count("}") and code.count("(") == code.count(")")
    assert os.path.exists(out / "rbd_backend_glue.h") and os.path.exists(out / "rbd_b200.h")
    # a second application on the already patched tree must fail loudly (no double patching)
    fake = tmp_path / "ref2" / "src" / "sofa" / "RigidBodyDynamics"
    fake.mkdir(parents=True)
    (fake / "KinematicChainMapping.h").write_text(hdr); (fake / "KinematicChainMapping.inl").write_text(inl)
    r2 = subprocess.run([sys.executable, os.path.join(INTEG, "apply_backend_patch.py"), str(tmp_path / "ref2"),
                         str(tmp_path / "again")], capture_output=True, text=True)
    assert r2.returncode != 0


def test_glue_header_compiles_and_flattens_a_model(rbd_lib, tmp_path):
    exe = str(tmp_path / "glue_check")
    lib_dir = os.path.join(ROOT, "sofa", "rigidbodydynamics_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "harness", "glue_check.cpp"), "-o", exe, "-L", lib_dir, "-lrbd_b200",
           "-Wl,-rpath," + lib_dir, "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    fixed = [l for l in lines if l.startswith("DESC ff=0")][0]
    assert "njoints=3 nq=2 nv=2 n_out=5 jtype=0,3,4," in fixed and "axis_last=0.6,0.0,0.8" in fixed
    assert "out=0,1,2,4,3," in fixed and "pl_last_px=0.50" in fixed and "inertia1=1.500,0.030" in fixed and "grav_z=-9.81" in fixed
    floating = [l for l in lines if l.startswith("DESC ff=1")][0]
    assert "njoints=4 nq=9 nv=8 n_out=6 jtype=0,13,3,4," in floating
    models = [l for l in lines if l.startswith("MODEL")]
    assert models[0].startswith("MODEL rc=0 has_ff=0 nq_joints=2 nv_joints=2 n_out=5 mass_support=3")
    assert models[1].startswith("MODEL rc=0 has_ff=1 nq_joints=2 nv_joints=2 n_out=6")
    be = [l for l in lines if l.startswith("BACKEND")]
    import torch
    if torch.cuda.is_available():
        assert all("rc=0 ws=set model=set" in l for l in be)
        assert all(l.split()[1] == "rc=0" and l.split()[2] == "rows_rc=0" for l in lines if l.startswith("APPLY"))
    else:
        assert all("rc=-3 ws=null model=null" in l for l in be)     # RBD_ERR_CUDA, handles released: no fallback
    assert [l for l in lines if l.startswith("UNSUPPORTED")][0].startswith("UNSUPPORTED ok=0 why=joint 2 (JointModelSpherical)")
    assert [l for l in lines if l.startswith("CODES")][0] == "CODES 1 13 -1"

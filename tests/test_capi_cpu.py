"""The C-ABI library loads, exports every symbol include/rbd_b200.h declares, and fails loudly (RBD_ERR_CUDA) when no
device is present: there is no CPU fallback behind the boundary."""
import ctypes as C
import os
import re
import subprocess

import pytest

import helpers


def _header_symbols():
    with open(os.path.join(helpers.ROOT, "include", "rbd_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"RBD_API\s+[A-Za-z_0-9\s\*]*?\b(rbd_[a-zA-Z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(rbd_lib):
    from sofa.rigidbodydynamics_b200 import _capi
    declared = _header_symbols()
    assert len(declared) >= 28
    assert sorted(_capi.SYMBOLS) == declared, "python binding list and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(rbd_[a-zA-Z0-9_]+)", out))
    assert set(declared) <= exported, set(declared) - exported
    # nothing but the C ABI is exported as a strong symbol (-fvisibility=hidden; weak libstdc++ template
    # instantiations keep their default visibility and are harmless)
    leaked = [s for s in re.findall(r"\sT\s+(\S+)", out) if not s.startswith("rbd_")]
    assert not leaked, leaked[:5]


def test_product_does_not_link_the_oracle(rbd_lib):
    from sofa.rigidbodydynamics_b200 import _capi
    out = subprocess.run(["nm", "-D", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out
    src = os.path.join(helpers.ROOT, "sofa", "rigidbodydynamics_b200")
    for root, _dirs, files in os.walk(src):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                with open(os.path.join(root, fn)) as f:
                    text = f.read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("//", "#", "*", "/*")))
                assert "import oracle" not in code and "from oracle" not in code and "rbd_oracle" not in code \
                    and "orc_" not in code, fn


def test_version_and_error_string(rbd_lib):
    assert rbd_lib.rbd_version() == 100
    rbd_lib.rbd_clear_error()
    assert rbd_lib.rbd_last_error() == b""
    assert rbd_lib.rbd_model_create(None, None) == -1
    assert b"null" in rbd_lib.rbd_last_error()


def test_no_device_is_an_error_not_a_fallback(rbd_lib):
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = api.Model(mdl.panda7())
    assert m.info.nq == 7 and m.info.eval_bytes == 1456 and m.info.mass_fields == 28
    with pytest.raises(api.RbdError) as e:
        api.Workspace(m, 16)
    assert e.value.code == -3                                      # RBD_ERR_CUDA


def test_model_info_reports_the_storage_plan(rbd_lib):
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    t = api.Model(mdl.talos_reduced()).info
    assert (t.njoints, t.nq, t.nv, t.n_out) == (34, 39, 38, mdl.talos_reduced().n_out)
    assert t.has_freeflyer_root == 1 and t.nq_joints == 32 and t.nv_joints == 32
    assert t.omi_fields == 396 and t.jac_fields == 228 and t.mass_fields == 741 and t.eval_bytes == 12144
    assert t.eval_warps_per_sm >= 6 and t.eval_tmem_doubles > 0
    assert (t.eval_smem_doubles + 0) * 8 * 32 * t.eval_warps_per_sm <= 227 * 1024

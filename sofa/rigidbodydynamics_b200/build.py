"""In-tree build of ``librbd_b200.so`` (hand-written CUDA for sm_100a + the C ABI), with plain nvcc.

    python -m sofa.rigidbodydynamics_b200.build [--force] [-v] [-DNAME=VALUE ...]

(-D...: an A/B build with extra preprocessor definitions; combine with RBD_B200_LIB_OUT=<path> to keep both libraries)

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("RBD_B200_LIB_OUT") or os.path.join(HERE, "librbd_b200.so")
SOURCES = ["rbd_kernels.cu", "rbd_kernels_dyn.cu", "rbd_kernels_arm.cu", "rbd_capi.cu", "rbd_probe.cu", "rbd_urdf.cpp"]
HEADERS = ["rbd_dev_model.h", "rbd_sweeps.h", "rbd_sweeps_body.h", "rbd_dyn_body.h", "rbd_deriv_body.h", "rbd_arm_body.h", "rbd_kernels.h", "rbd_model_build.h"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "static"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


# per translation unit: the headers it includes (include/rbd_b200.h and this file are dependencies of every unit)
UNIT_DEPS = {
    "rbd_kernels.cu": HEADERS,
    "rbd_kernels_dyn.cu": HEADERS,
    "rbd_kernels_arm.cu": HEADERS,
    "rbd_capi.cu": HEADERS,
    "rbd_probe.cu": [],
    "rbd_urdf.cpp": [],
}
OBJDIR = os.path.join(HERE, "..", "..", "build", "obj")


def _unit_deps(src: str):
    return [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in UNIT_DEPS[src]] + \
           [os.path.join(HERE, "..", "..", "include", "rbd_b200.h"), os.path.abspath(__file__)]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for s in SOURCES for d in _unit_deps(s))


def build(force: bool = False, verbose: bool = False, defines=()) -> str:
    """Compile every translation unit (in parallel, objects cached under build/obj by modification time) and link
    the shared library.  `defines` (A/B builds) bypasses the object cache."""
    if not force and not defines and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)   # use the system host compiler nvcc finds on PATH
    objdir = OBJDIR + ("_" + "_".join(d.replace("=", "-") for d in defines) if defines else "")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("--shared", "-cudart", "static")]

    def compile_unit(src: str):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in _unit_deps(src)):
            return obj, None
        cmd = [_nvcc()] + compile_flags + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", obj, os.path.join(CSRC, src)]
        r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_unit, SOURCES))
    if verbose:
        for _obj, err in results:
            if err:
                sys.stderr.write(err)
    cmd = [_nvcc(), "--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + \
          [o for o, _ in results]
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv,
              defines=tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D")))
    print(p)

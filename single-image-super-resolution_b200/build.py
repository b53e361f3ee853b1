"""In-tree nvcc build of the CUDA library (sm_100a only).

    python single-image-super-resolution_b200/build.py            # libsisr_b200.so
    python single-image-super-resolution_b200/build.py harness    # + build/harness_igemm
    python single-image-super-resolution_b200/build.py probes     # + build/probe_{shift,pdl,mma}

The shared library lands next to this file so that it travels to the GPU box with the repo
snapshot (built artefacts are git-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsisr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler",
          "-Wall", "-Xcompiler", "-Wno-unused-function"]
LIB_SOURCES = ["abi.cu", "igemm_tc.cu", "igemm_pm.cu", "wgrad_tc.cu", "conv_simt.cu", "conv_thin.cu", "elementwise.cu", "spectral.cu",
               "linear.cu", "optim.cu", "peer.cu", "metrics.cu", "tmap.cpp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = list(sources) + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "sisr_b200.h"))
    return any(os.path.getmtime(s) > t for s in deps)


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_lib(force=False, verbose_ptxas=False):
    objs = []
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    for src in LIB_SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        if force or _newer(obj, [sp]):
            extra = ["-Xptxas", "-v"] if verbose_ptxas else []
            _run([NVCC, *ARCH, *COMMON, *extra, "-c", sp, "-o", obj])
        objs.append(obj)
    if force or _newer(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


def build_harness():
    out = os.path.join(ROOT, "build", "harness_igemm")
    srcs = [os.path.join(CSRC, f) for f in ("harness_igemm.cu", "igemm_tc.cu", "igemm_pm.cu", "conv_simt.cu", "conv_thin.cu", "tmap.cpp")]
    if _newer(out, srcs):
        _run([NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-o", out, *srcs, "-lcudart"])
    return out


def build_probes():
    """Stand-alone hardware probes (not part of the library): build/probe_{shift,pdl,mma}."""
    outs = []
    for name, extra in (("probe_shift", ["tmap.cpp"]), ("probe_pdl", []), ("probe_mma", [])):
        out = os.path.join(ROOT, "build", name)
        srcs = [os.path.join(CSRC, name + ".cu")] + [os.path.join(CSRC, f) for f in extra]
        if _newer(out, srcs):
            _run([NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-o", out, *srcs, "-lcudart", "-lcuda"])
        outs.append(out)
    return outs


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    if "harness" in sys.argv:
        build_harness()
    if "probes" in sys.argv:
        build_probes()

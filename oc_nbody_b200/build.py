"""In-tree build of liboc_nbody_b200.so (CUDA, sm_100a).

The library is compiled with nvcc straight from ``oc_nbody_b200/csrc`` into ``oc_nbody_b200/lib`` so that the
built ``.so`` travels with the repository snapshot to the GPU box.  nvcc cross-compiles without a GPU.
(The CPU oracle is test infrastructure and builds itself: ``oracle.build()``.)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "oc_nbody_b200", "csrc")
LIBDIR = os.path.join(ROOT, "oc_nbody_b200", "lib")
LIB = os.path.join(LIBDIR, "liboc_nbody_b200.so")
SOURCES = ["api.cu", "direct_sum.cu", "self_gravity.cu", "grid_interp.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into LIB. Returns the path."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "ocg_internal.cuh"), os.path.join(ROOT, "include", "ocg.h")]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building %s" % LIB)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

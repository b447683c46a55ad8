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
# (object name, source, extra flags); the translation units compile concurrently
SOURCES = [
    ("api", "api.cu", []),
    ("direct_sum", "direct_sum.cu", []),
    ("self_gravity", "self_gravity.cu", []),
    ("hermite", "hermite.cu", []),
    ("grid_interp", "grid_interp.cu", []),
    ("rbf_interp", "rbf_interp.cu", []),
    ("cluster_ops", "cluster_ops.cu", []),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
]
OBJDIR = os.path.join(ROOT, "build", "obj")


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
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("ocg_internal.cuh", "direct_kernel.cuh")] + [os.path.join(ROOT, "include", "ocg.h")]
    objs, procs = [], []
    for name, src, extra in SOURCES:
        obj = os.path.join(OBJDIR, name + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
            procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for obj, pr in procs:  # the translation units compile concurrently
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError("nvcc failed building %s" % obj)
        if verbose:
            sys.stderr.write(err)
    if procs or force or _stale(LIB, objs):
        res = subprocess.run([_nvcc(), "--shared", "-o", LIB] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking %s" % LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

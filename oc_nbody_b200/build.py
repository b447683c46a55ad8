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
# the same sources with -DOCG_TUNING: every sweep shape and the timing-only experiments (tools/probe*.py); never loaded
# by the product, tests or bench
LIB_TUNING = os.path.join(LIBDIR, "liboc_nbody_b200_tuning.so")
# (object name, source, extra flags); the translation units compile concurrently
SOURCES = [
    ("api", "api.cu", []),
    ("direct_sum", "direct_sum.cu", []),
    ("self_gravity", "self_gravity.cu", []),
    ("hermite", "hermite.cu", []),
    ("grid_interp", "grid_interp.cu", []),
    ("rbf_interp", "rbf_interp.cu", []),
    ("cluster_ops", "cluster_ops.cu", []),
    ("comm", "comm.cu", []),
    ("assemble", "assemble.cu", []),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # only the extern "C" entry points of include/ocg.h / ocg_debug.h are exported (OCG_API); internals stay hidden
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden",
]
OBJDIR = os.path.join(ROOT, "build", "obj")
OBJDIR_TUNING = os.path.join(ROOT, "build", "obj_tuning")


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


def build_library(force=False, verbose=False, tuning=False):
    """Compile every CUDA source for sm_100a into LIB (tuning=True: LIB_TUNING, with -DOCG_TUNING). Returns the path."""
    lib, objdir = (LIB_TUNING, OBJDIR_TUNING) if tuning else (LIB, OBJDIR)
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")] + [
        os.path.join(ROOT, "include", h) for h in ("ocg.h", "ocg_debug.h")]
    objs, procs = [], []
    for name, src, extra in SOURCES:
        obj = os.path.join(objdir, name + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-DOCG_TUNING"] if tuning else []) + (["-Xptxas", "-v"] if verbose else []) + [
                "-c", "-o", obj, os.path.join(CSRC, src)]
            procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for obj, pr in procs:  # the translation units compile concurrently
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError("nvcc failed building %s" % obj)
        if verbose:
            sys.stderr.write(err)
    if procs or force or _stale(lib, objs):
        res = subprocess.run([_nvcc(), "--shared", "-o", lib] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking %s" % lib)
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))

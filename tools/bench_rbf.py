"""K7 (kNN + RBF-PHS get_gravity_at_point, the reference's own spatial interpolation) on a B200:
ms per kick and stars/s at the reference's N = 1 024 and at N = 16 384, with and without the tidal tensor, next to the
same evaluation by scipy (cKDTree + RBFInterpolator, the reference's CPU mechanism) on a bounded sample.
python tools/bench_rbf.py -> gpurun_out/bench_rbf.json"""
import ctypes
import json
import os
import sys

os.environ.setdefault("OCG_TUNING_LIB", "1")  # sweep shapes / phase counters live in the OCG_TUNING build
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_extra import timeit  # noqa: E402
from oc_nbody_b200 import default_context  # noqa: E402


def main():
    ctx = default_context(0)
    dev = torch.device("cuda", 0)
    n = 16
    ax = np.linspace(-0.05, 0.05, n)
    o = np.array([8.0, 0.0, 0.0])
    g = np.stack(np.meshgrid(ax + o[0], ax + o[1], ax + o[2], indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([g, o[None]])
    rng = np.random.default_rng(3)
    f = rng.normal(0, 1e-3, (4, pts.shape[0]))
    nodes = [torch.from_numpy(ax).to(dev)] * 3
    d_o, d_f = torch.from_numpy(o[None]).to(dev), torch.from_numpy(f).to(dev)
    out = {}
    if "--profile" in sys.argv:
        p = o + rng.normal(0, 0.002, (296, 3))
        sx, sy, sz = (torch.from_numpy(np.ascontiguousarray(p[:, k])).to(dev) for k in range(3))
        res = torch.empty((4, 296), dtype=torch.float64, device=dev)
        for _ in range(3):
            ctx.grid_interp_rbf((n, n, n), nodes, d_o, d_f, sx, sy, sz, None, res)
        torch.cuda.synchronize()
        return
    for ns in (1024, 16384):
        p = o + rng.normal(0, 0.002, (ns, 3))
        sx, sy, sz = (torch.from_numpy(np.ascontiguousarray(p[:, k])).to(dev) for k in range(3))
        res = torch.empty((4, ns), dtype=torch.float64, device=dev)
        ten = torch.empty((3, 4, ns), dtype=torch.float64, device=dev)
        st = torch.empty(ns, dtype=torch.int32, device=dev)
        for name, t in (("values", None), ("values_and_tensor", ten)):
            def k7():
                ctx.grid_interp_rbf((n, n, n), nodes, d_o, d_f, sx, sy, sz, None, res, tensor_out=t, status_out=st)
            med, best = timeit(k7, iters=5, warm=2)
            cyc = (ctypes.c_double * 6)()
            ctx.lib.ocg_debug_rbf_phase_cycles(ctx.h, cyc)
            k7()
            torch.cuda.synchronize()
            ctx.lib.ocg_debug_rbf_phase_cycles(ctx.h, cyc)
            tot = sum(cyc) or 1.0
            out["k7_rbf_%d_stars_%s" % (ns, name)] = dict(ms_median=med, ms_best=best, stars_per_s=ns / med * 1e3,
                                                          us_per_star_per_sm=med * 1e3 * ctx.sm_count / ns,
                                                          status_nonzero=int(((st & 0xff) != 0).sum().item()),
                                                          refinement_iterations_mean=float((st >> 8).double().mean().item()),
                                                          phase_share=dict(zip(("select", "assemble", "factorise", "residual", "solve", "output"),
                                                                               [round(c / tot, 4) for c in cyc])),
                                                          fp32_lu_flop_per_star=2.0 / 3.0 * 206 ** 3)
    # the reference's CPU mechanism on a sample: per star, kNN + 3 RBF interpolants (gizmo_interface.py:661-675, 698-704)
    from scipy.interpolate import RBFInterpolator
    from scipy.spatial import cKDTree
    tree = cKDTree(pts)
    sample = o + rng.normal(0, 0.002, (48, 3))
    t0 = time.perf_counter()
    for q in sample:
        _, ids = tree.query(q, 150)
        for c in range(3):
            RBFInterpolator(pts[ids], f[c][ids], kernel="cubic", degree=5)(q[None])
    dt = time.perf_counter() - t0
    out["cpu_scipy_knn_rbf"] = dict(stars=len(sample), s_total=dt, ms_per_star=dt / len(sample) * 1e3, stars_per_s=len(sample) / dt,
                                    cores=1, note="cKDTree.query(150) + 3 x RBFInterpolator(cubic, degree 5) per star, as the reference loops")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_rbf.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

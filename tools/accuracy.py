"""Accuracy / throughput survey of K1 on a B200: error against the FP64 oracle on configs[1]-shaped data (a sample of
the 64^3 lattice x a synthetic snapshot) for the kernel shapes that differ in FOLD (sources per FP32 accumulation run),
with and without the FP64 precision-radius set, in both error metrics, on the raw field and on the tidal residual the
reference returns (gizmo_interface.py:569-573); plus the kernel time of each shape on the full 64^3 lattice.

python tools/accuracy.py [n_src] [n_targets] -> gpurun_out/accuracy.json   (uses the OCG_TUNING build for the sweep)"""
import json
import os
import sys
import time

os.environ.setdefault("OCG_TUNING_LIB", "1")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from bench import CENTER, G_KPC, make_sources, make_targets  # noqa: E402
from oc_nbody_b200 import default_context  # noqa: E402
from util import rel_err  # noqa: E402

# (shape id, FOLD, wants potential)
SHAPES = [(58, 512, False), (66, 128, False), (67, 64, False), (81, 64, False), (68, 32, False),
          (31, 512, True), (70, 64, True), (76, 512, True), (74, 64, True), (81, 64, True), (75, 64, True), (78, 512, True), (77, 64, True)]


def main():
    n_src = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    ctx = default_context(0)
    g = make_targets(64)
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(g.evolved_grid) - 1, n_t, replace=False))
    tgt = np.concatenate([g.evolved_grid[pick], g.evolved_grid[-1:]])
    pos, mass, eps = make_sources(n_src, seed=1776)
    s32 = oracle.recentre(pos, mass, CENTER)
    t32 = oracle.recentre(tgt, None, CENTER)
    tfull = oracle.recentre(g.evolved_grid, None, CENTER)
    out = {"n_src": n_src, "n_tgt": int(tgt.shape[0]), "timing_targets": int(tfull.shape[0]), "cases": []}
    d_src, d_tgt, d_full = torch.from_numpy(s32).cuda(), torch.from_numpy(t32).cuda(), torch.from_numpy(tfull).cuda()
    nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    for kernel, soft in ((0, eps.astype(np.float32)), (1, (eps * 2.8).astype(np.float32))):
        t0 = time.time()
        ref, pref = oracle.field_direct(s32, soft, t32, kernel, G_KPC, want_pot=True)
        t_or = time.time() - t0
        d_soft = torch.from_numpy(soft).cuda()
        for variant, fold, want_pot in SHAPES:
            if not ctx.variant_built(variant):
                continue
            for precise in (1, 0):
                if precise == 0 and (kernel == 1 or fold not in (512, 64)):
                    continue
                ctx.debug_set("direct_variant", variant)
                ctx.debug_set("precise_near", precise)
                acc = torch.empty((3, tgt.shape[0]), dtype=torch.float64, device="cuda")
                pot = torch.empty(tgt.shape[0], dtype=torch.float64, device="cuda") if want_pot else None
                ctx.field_direct(d_src, d_soft, d_tgt, kernel, G_KPC, acc, pot)
                torch.cuda.synchronize()
                a = acc.cpu().numpy()
                sub_g, sub_r = a - a[:, -1:], ref - ref[:, -1:]
                # kernel time on the full lattice
                accf = torch.empty((3, tfull.shape[0]), dtype=torch.float64, device="cuda")
                potf = torch.empty(tfull.shape[0], dtype=torch.float64, device="cuda") if want_pot else None
                ctx.set_kernel_timing(True)
                best = 1e30
                for _ in range(3):
                    ctx.field_direct(d_src, d_soft, d_full, kernel, G_KPC, accf, potf)
                    torch.cuda.synchronize()
                    best = min(best, ctx.last_direct_kernel_ms())
                ctx.set_kernel_timing(False)
                inter = float(n_src) * tfull.shape[0]
                case = dict(kernel=kernel, variant=variant, name=ctx.variant_name(variant), fold=fold, precise_near=precise,
                            oracle_s=t_or, raw_norm=rel_err(a, ref, floor=1.0), raw_strict=rel_err(a, ref),
                            err_pot=float(np.max(np.abs(pot.cpu().numpy() - pref) / np.abs(pref))) if want_pot else None,
                            residual_norm=rel_err(sub_g[:, :-1], sub_r[:, :-1], floor=1.0),
                            residual_strict=rel_err(sub_g[:, :-1], sub_r[:, :-1]),
                            kernel_ms_full_lattice=best, pct_fp32_peak=100 * 20 * inter / best / 1e9 / nominal)
                out["cases"].append(case)
                print(json.dumps(case), flush=True)
    ctx.debug_set("direct_variant", -1)
    ctx.debug_set("precise_near", 1)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/accuracy.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

"""Accuracy survey on a B200: error of K1 against the FP64 oracle on config-2-shaped data
(a subsample of the 64^3 lattice x a synthetic snapshot), with and without the FP64 near-field path,
in both error metrics.  python tools/accuracy.py [n_src] [n_targets] -> gpurun_out/accuracy.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from bench import CENTER, G_KPC, make_sources, make_targets  # noqa: E402
from oc_nbody_b200 import default_context  # noqa: E402
from util import rel_err, rel_err_strict  # noqa: E402


def main():
    n_src = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    ctx = default_context(0)
    g = make_targets(64)
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(g) - 1, n_t, replace=False))
    tgt = np.concatenate([g.evolved_grid[pick], g.evolved_grid[-1:]])
    pos, mass, eps = make_sources(n_src, seed=1776)
    s32 = oracle.recentre(pos, mass, CENTER)
    t32 = oracle.recentre(tgt, None, CENTER)
    out = {"n_src": n_src, "n_tgt": int(tgt.shape[0]), "cases": []}
    d_src, d_tgt = torch.from_numpy(s32).cuda(), torch.from_numpy(t32).cuda()
    for kernel, soft in ((0, eps.astype(np.float32)), (1, (eps * 2.8).astype(np.float32))):
        t0 = time.time()
        ref, pref = oracle.field_direct(s32, soft, t32, kernel, G_KPC, want_pot=True)
        t_or = time.time() - t0
        d_soft = torch.from_numpy(soft).cuda()
        # 31 = plain target-paired kernel (K4's; K1 when the potential is wanted); 58 = production mass-folded kernel
        for variant, precise in ((31, 1), (31, 0), (58, 1), (58, 0)):
            ctx.lib.ocg_debug_set_variant(variant)
            ctx.lib.ocg_debug_set_precise_near(precise)
            want_pot = variant not in (46, 58)
            acc = torch.empty((3, tgt.shape[0]), dtype=torch.float64, device="cuda")
            pot = torch.empty(tgt.shape[0], dtype=torch.float64, device="cuda") if want_pot else None
            ctx.field_direct(d_src, d_soft, d_tgt, kernel, G_KPC, acc, pot)
            torch.cuda.synchronize()
            a = acc.cpu().numpy()
            sub_g, sub_r = a - a[:, -1:], ref - ref[:, -1:]
            case = dict(kernel=kernel, variant=variant, precise_near=precise, oracle_s=t_or, err_gate=rel_err(a, ref),
                        err_strict=rel_err_strict(a, ref),
                        err_pot=float(np.max(np.abs(pot.cpu().numpy() - pref) / np.abs(pref))) if want_pot else None,
                        err_tidal_residual_gate=rel_err(sub_g[:, :-1], sub_r[:, :-1]),
                        err_tidal_residual_strict=rel_err_strict(sub_g[:, :-1], sub_r[:, :-1]))
            out["cases"].append(case)
            print(json.dumps(case), flush=True)
    ctx.lib.ocg_debug_set_variant(-1)
    ctx.lib.ocg_debug_set_precise_near(1)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/accuracy.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

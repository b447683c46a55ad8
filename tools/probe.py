"""GPU probes: FP32 issue peaks (FFMA, FFMA2, MUFU.RSQ) and every variant of the direct-sum kernel.
Run on a B200:  python tools/probe.py [n_src] [n_grid]  -> gpurun_out/probe.json"""
import json
import os
import sys

os.environ.setdefault("OCG_TUNING_LIB", "1")  # sweep shapes / phase counters live in the OCG_TUNING build

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oc_nbody_b200 import default_context  # noqa: E402


def main():
    n_src = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    n_grid = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ctx = default_context(0)
    out = {"sm_count": ctx.sm_count, "sm_clock_khz": ctx.sm_clock_khz}
    out["ffma_tflops"] = ctx.probe_throughput(0)
    out["ffma2_tflops"] = ctx.probe_throughput(1)
    out["rsq_gops"] = ctx.probe_throughput(2)
    out["nominal_tflops"] = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    print(json.dumps(out), flush=True)

    rng = np.random.default_rng(0)
    src = np.concatenate([rng.normal(0, 5, (n_src, 3)), rng.uniform(1e3, 1e5, (n_src, 1))], axis=1).astype(np.float32)
    soft = rng.uniform(0.004, 0.1, n_src).astype(np.float32)
    ax = np.linspace(-0.6, 0.6, n_grid)
    tg = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    tg = np.concatenate([tg, np.zeros((1, 3))])
    tgt = np.concatenate([tg, np.zeros((tg.shape[0], 1))], 1).astype(np.float32)
    d_src, d_soft, d_tgt = torch.from_numpy(src).cuda(), torch.from_numpy(soft).cuda(), torch.from_numpy(tgt).cuda()
    acc = torch.empty((3, tgt.shape[0]), dtype=torch.float64, device="cuda")
    pot = torch.empty(tgt.shape[0], dtype=torch.float64, device="cuda")
    inter = float(n_src) * tgt.shape[0]
    ctx.set_kernel_timing(True)
    res = []
    nvar = ctx.variant_count()
    only = [int(x) for x in os.environ.get("OCG_PROBE_VARIANTS", "").split(",") if x]
    for kernel, want_pot in ((0, False), (0, True), (1, False)):
        for v in range(nvar):
            if not ctx.variant_built(v) or (only and v not in only):
                continue
            if not only and (want_pot or kernel == 1) and v not in (0, 1, 31, 70, 74) and not (kernel == 1 and v in (46, 58, 67)):
                continue
            ctx.debug_set("direct_variant", v)
            best = 1e30
            try:
                for rep in range(3):
                    ctx.field_direct(d_src, d_soft, d_tgt, kernel, 1.0, acc, pot if want_pot else None)
                    torch.cuda.synchronize()
                    best = min(best, ctx.last_direct_kernel_ms())
            except Exception as exc:  # noqa: BLE001  (a shape without this form)
                print("variant", v, "pot" if want_pot else "", "skipped:", str(exc)[:80], flush=True)
                continue
            r = dict(variant=v, name=ctx.variant_name(v), kernel=kernel, pot=want_pot, ms=best,
                     ginter_s=inter / best / 1e6, pct_peak=100 * 20 * inter / best / 1e9 / out["nominal_tflops"])
            res.append(r)
            print(json.dumps(r), flush=True)
    ctx.debug_set("direct_variant", -1)
    out["variants"] = res
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

"""configs[0]-shaped accuracy probe (16^3 grid, 1e6-particle snapshot, spline kernel, potential on): tidal-residual error
(strict metric) of the field build against the size of the FP64 precision-radius set and the FOLD of the mid-size kernel.
python tools/accuracy_c0.py -> gpurun_out/accuracy_c0.json"""
import json
import os
import sys

os.environ.setdefault("OCG_TUNING_LIB", "1")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oc_nbody_b200 import default_context  # noqa: E402
from oc_nbody_b200.gizmo_field import gizmo_field  # noqa: E402
from oc_nbody_b200.synthetic import make_snapshot  # noqa: E402
from util import rel_err  # noqa: E402


def main():
    ctx = default_context(0)
    snap = make_snapshot(1_000_000, seed=1776)
    center = np.array([8.0, 0.0, 0.0])
    opts = dict(grid_x_size_in_kpc=0.6, grid_y_size_in_kpc=0.6, grid_z_size_in_kpc=0.6, grid_resolution=0.6 / 16, softening_kernel="spline")
    field = gizmo_field(opts, [snap], chosen_positions=center[None], ctx=ctx, build=False)
    from oc_nbody_b200.grid_cartesian import grid
    g = grid(0.6, 0.6, 0.6, 0.6 / 16)
    r, m, soft = field._source_arrays_(snap)
    s32 = oracle.recentre(r, m, center)
    t32 = oracle.recentre(g.init_grid + center, None, center)
    raw = oracle.field_direct(s32, soft.astype(np.float32), t32, oracle.KERNEL_SPLINE, field.G)
    sub = raw - raw[:, g.origin_row:g.origin_row + 1]
    keep = np.arange(len(g)) != g.origin_row
    d_src, d_soft, d_tgt = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (s32, soft.astype(np.float32), t32))
    out = []
    for variant in (80, 77, 79):
        if not ctx.variant_built(variant):
            continue
        for cap in (16384, 65536, 262144):
            ctx.debug_set("direct_variant", variant)
            ctx.debug_set("near_cap", cap)
            acc = torch.empty((3, len(g)), dtype=torch.float64, device="cuda")
            pot = torch.empty(len(g), dtype=torch.float64, device="cuda")
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.field_direct(d_src, d_soft, d_tgt, 1, field.G, acc, pot)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            a = acc.cpu().numpy()
            row = dict(variant=variant, name=ctx.variant_name(variant), near_cap=cap, ms=min(ts), raw_strict=rel_err(a, raw),
                       residual_strict=rel_err((a - a[:, g.origin_row:g.origin_row + 1])[:, keep], sub[:, keep]))
            out.append(row)
            print(json.dumps(row), flush=True)
    ctx.debug_set("direct_variant", -1)
    ctx.debug_set("near_cap", 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/accuracy_c0.json", "w"), indent=1)


if __name__ == "__main__":
    main()

"""K4 (cluster self-gravity) over the plain-tile kernel variants: which tile shape balances best at the BASELINE sizes.
python tools/probe_k4.py -> gpurun_out/probe_k4.json"""
import json
import os
import sys

os.environ.setdefault("OCG_TUNING_LIB", "1")  # sweep shapes / phase counters live in the OCG_TUNING build

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oc_nbody_b200 import default_context  # noqa: E402
from oc_nbody_b200.synthetic import make_plummer_cluster  # noqa: E402
from oc_nbody_b200.units import G_KPC_KMS_MYR  # noqa: E402


def main():
    ctx = default_context(0)
    dev = torch.device("cuda", 0)
    nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    out = []
    # "shardK": the K targets [0, K) of one 65 536-star cluster — what one rank of 65536/K runs in the sharded BRIDGE step
    for name, nseg, npc in (("n65536", 1, 65536), ("256x4096", 256, 4096), ("n16384", 1, 16384), ("16x16384", 16, 16384),
                            ("shard8192", 1, 65536), ("shard16384", 1, 65536), ("shard32768", 1, 65536), ("n20000", 1, 20000)):
        shard = int(name[5:]) if name.startswith("shard") else None
        pos_pc, _, mass = make_plummer_cluster(npc)
        pos = np.concatenate([pos_pc * 1e-3 + np.array([[8.0 + 0.01 * k], [0.0], [0.0]]) for k in range(nseg)], axis=1)
        m = np.tile(mass, nseg)
        seg = np.arange(nseg + 1, dtype=np.int64) * npc
        d_pos, d_m = torch.from_numpy(np.ascontiguousarray(pos)).to(dev), torch.from_numpy(m).to(dev)
        a = torch.empty((3, nseg * npc), dtype=torch.float64, device=dev)
        inter = float(nseg) * npc * npc if shard is None else float(shard) * npc
        for v in [-1] + [x for x in ([26, 27] if shard or name == "n20000" else list(range(21, 37)) + [0, 4, 69, 70, 71, 72, 73]) if ctx.variant_built(x)]:
            ctx.debug_set("direct_variant", v)
            ts = []
            try:
                for rep in range(8):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ctx.self_gravity(d_pos, d_m, (0.01e-3) ** 2, G_KPC_KMS_MYR, a, None, seg_offsets=seg if nseg > 1 else None,
                                     tgt_begin=0, tgt_end=shard)
                    e1.record()
                    torch.cuda.synchronize()
                    if rep >= 2:
                        ts.append(e0.elapsed_time(e1))
            except Exception as exc:  # noqa: BLE001
                print(name, v, "failed:", exc)
                continue
            finally:
                ctx.debug_set("direct_variant", -1)
            ms = float(np.median(ts))
            r = dict(config=name, variant=v, name=ctx.variant_name(v) if v >= 0 else "heuristic", ms=ms,
                     pct_fp32_peak=100 * 20 * inter / ms / 1e9 / nominal)
            out.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_k4.json", "w"), indent=1)


if __name__ == "__main__":
    main()

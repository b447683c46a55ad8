"""K6 (Hermite force loop: acceleration + jerk) on a B200: every kernel variant at configs[2] size (N = 65 536) and on
a configs[3]-shaped batch (256 x 4096), the fused small-cluster kernel at the reference's own N = 1 024, and the
BRIDGE step with the Hermite cluster code (eager and as one CUDA-graph launch).
Roofline: FP32 FMA pipe; 41 flop per interaction (FMA = 2: 6 sub, 6 r^2, 5 d.w, 1 rsqrt, 3 r^-2/m r^-1/m r^-3, 2 alpha,
6 t = alpha d + w, 12 accumulate) executed as 26 FMA-pipe operations, so the ceiling is 41/52 = 78.8 % of peak.
python tools/bench_hermite.py -> gpurun_out/bench_hermite.json"""
import json
import os
import sys

os.environ.setdefault("OCG_TUNING_LIB", "1")  # sweep shapes / phase counters live in the OCG_TUNING build

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_extra import timeit  # noqa: E402
from oc_nbody_b200 import default_context  # noqa: E402
from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR  # noqa: E402
from oc_nbody_b200.units import G_KPC_KMS_MYR  # noqa: E402

FLOP = 41.0
OPS = 26.0


def main():
    ctx = default_context(0)
    dev = torch.device("cuda", 0)
    nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    out = {"fp32_peak_nominal_tflops": nominal, "flop_per_interaction": FLOP, "fma_pipe_ops_per_interaction": OPS}
    from oc_nbody_b200.synthetic import make_plummer_cluster
    eps2 = (0.01e-3) ** 2
    ncl = 256
    ang = np.linspace(0, 2 * np.pi, ncl, endpoint=False)
    origin = np.stack([8 * np.cos(ang), 8 * np.sin(ang), np.zeros(ncl)], 1)
    nv = ctx.variant_count(family=1)
    if "--profile" in sys.argv:
        # the launch an ncu capture targets: production variant, N = 65 536, no potential, 6 calls
        pos_pc, vel1, mass = make_plummer_cluster(65536)
        d_pos, d_vel, d_m = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (pos_pc * 1e-3 + origin[0][:, None], vel1, mass))
        a = torch.empty((3, 65536), dtype=torch.float64, device=dev)
        j = torch.empty((3, 65536), dtype=torch.float64, device=dev)
        for _ in range(6):
            ctx.self_gravity_hermite(d_pos, d_vel, d_m, eps2, G_KPC_KMS_MYR, KMS_TO_KPC_PER_MYR, a, j, None)
        torch.cuda.synchronize()
        return
    for name, nseg, npc in (("k6_hermite_n65536", 1, 65536), ("k6_hermite_256x4096", 256, 4096), ("k6_hermite_n1024_small", 1, 1024),
                            ("k6_hermite_n4096_small", 1, 4096)):
        pos_pc, vel1, mass = make_plummer_cluster(npc)
        pos = np.concatenate([pos_pc * 1e-3 + origin[k % ncl][:, None] for k in range(nseg)], axis=1)
        vel = np.tile(vel1, (1, nseg))
        m = np.tile(mass, nseg)
        seg = np.arange(nseg + 1, dtype=np.int64) * npc
        d_pos, d_vel, d_m = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (pos, vel, m))
        a = torch.empty((3, nseg * npc), dtype=torch.float64, device=dev)
        j = torch.empty((3, nseg * npc), dtype=torch.float64, device=dev)
        ph = torch.empty(nseg * npc, dtype=torch.float64, device=dev)
        inter = float(nseg) * npc * npc
        res = {}
        variants = range(nv) if npc > 4096 or nseg > 1 else [-1]
        for v in variants:
            ctx.debug_set("hermite_variant", v)
            for pot in (None, ph):
                def k6():
                    ctx.self_gravity_hermite(d_pos, d_vel, d_m, eps2, G_KPC_KMS_MYR, KMS_TO_KPC_PER_MYR, a, j, pot,
                                             seg_offsets=seg if nseg > 1 else None)
                med, best = timeit(k6, iters=10, warm=3)
                ctx.lib.ocg_set_kernel_timing(ctx.h, 1)
                k6()
                kms = ctx.lib.ocg_last_direct_kernel_ms(ctx.h)
                ctx.lib.ocg_set_kernel_timing(ctx.h, 0)
                key = ("auto" if v < 0 else ctx.variant_name(v, family=1)) + (" +pot" if pot is not None else "")
                res[key] = dict(ms_median=med, ms_best=best, kernel_ms=kms if kms > 0 else None, ginter_s=inter / med / 1e6,
                                tflops=FLOP * inter / med / 1e9, pct_fp32_peak=100 * FLOP * inter / med / 1e9 / nominal,
                                pct_fma_pipe_slots=100 * 2 * OPS * inter / med / 1e9 / nominal,
                                kernel_pct_fp32_peak=(100 * FLOP * inter / kms / 1e9 / nominal) if kms > 0 else None)
        ctx.debug_set("hermite_variant", -1)
        out[name] = dict(interactions=inter, note="pack + kernel per call, finish fused (small: one fused launch)", variants=res)

    # ---- the same force loop on the host cores (the oracle's FP64 OpenMP restatement; bounded sample of N = 65 536) ----
    if "--no-cpu-baseline" not in sys.argv:
        import time
        import oracle
        pos_pc, vel1, mass = make_plummer_cluster(65536)
        rows = 65536  # the whole configs[2]-size evaluation: a few seconds on 16 cores
        t0 = time.perf_counter()
        oracle.self_gravity_hermite(pos_pc * 1e-3 + origin[0][:, None], vel1, mass, eps2, G_KPC_KMS_MYR, KMS_TO_KPC_PER_MYR,
                                    t0=0, t1=rows)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = dict(value=rows * 65536.0 / dt / 1e9, unit="G interactions/s", cores=oracle.num_threads(), kind="port",
                                   sample="%d targets x 65 536 sources (one full evaluation), FP64 OpenMP acc + jerk (oracle/ocg_oracle.c), %.1f s" % (rows, dt))

    # ---- BRIDGE step with the Hermite cluster code ----
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.units import units
    rng = np.random.default_rng(7)
    center = np.array([8.0, 0.0, 0.0])
    nn = 16

    class _Snap(object):
        snapshot = {"index": 0, "time": 0.0}
    fake = gizmo_field.__new__(gizmo_field)
    gizmo_field.__init__(fake, dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05,
                                    grid_resolution=0.05 / nn), [_Snap(), _Snap()], time_in_Myr=[0.0, 23.0], build=False,
                         ctx=ctx)
    npts = nn ** 3 + 1
    tid = rng.normal(0, 1e-3, (2, 3, npts))
    fake.set_snapshot_fields(tid[:, 0], tid[:, 1], tid[:, 2], pot=rng.normal(0, 1, (2, npts)))
    fake.evolve_grid(center)
    for name, nst in (("bridge_step_hermite_1k_stars", 1024), ("bridge_step_hermite_65k_stars", 65536)):
        pos_pc, vel, mass = make_plummer_cluster(nst)
        for graph in (False, True):
            cl = cluster_code(mass, pos_pc * 1e-3 + center[:, None], vel, softening_pc=0.01, ctx=ctx, integrator="hermite")
            fake.evolve_model(0.0 | units.Myr)
            system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=graph)
            system.add_system(cl, (fake,))
            system.add_system(fake)
            state = {"t": 0.0}

            def step():
                state["t"] += 0.1
                system.evolve_model(state["t"] | units.Myr, timestep=0.1 | units.Myr)
            med, best = timeit(step, iters=20, warm=4)
            out[name + ("_cuda_graph" if graph else "")] = dict(
                ms_median=med, ms_best=best, graph_replays=system.graph_replays, aarseth_dt_myr=float(cl.dt_min.item()),
                note="K(dt/2) D(dt) K(dt/2); D = one shared 4th-order Hermite step: 2 force evaluations (acc + jerk) per step")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_hermite.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

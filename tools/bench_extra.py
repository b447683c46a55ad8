"""Secondary measurements on a B200 (not the headline bench):
  K3  grid_interp     configs[3]: 256 clusters x 4096 stars, 32^3 grids, 2 snapshots  -> GB/s vs HBM peak
  K4  self_gravity    configs[2]: N = 65 536 ; configs[3]: 256 x 4096 batched          -> interactions/s, % FP32 peak
  BRIDGE step         configs[0]: 1k stars + 16^3 grid ; configs[2]: 65k stars + 16^3  -> ms per step
python tools/bench_extra.py -> gpurun_out/bench_extra.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oc_nbody_b200 import default_context  # noqa: E402
from oc_nbody_b200.units import G_KPC_KMS_MYR  # noqa: E402

HBM_PEAK = 6454.0  # GB/s, MEASURED_PEAKS.json (driver-measured copy bandwidth on this pool)
try:
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # evict L2 (buffer larger than the 126 MB L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ctx = default_context(0)
    dev = torch.device("cuda", 0)
    out = {"hbm_peak_gbs": HBM_PEAK}
    rng = np.random.default_rng(7)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12

    # ---- K3, configs[3] ----
    ncl, nstar, n = 256, 4096, 32
    nodes = [torch.from_numpy(np.linspace(-0.05, 0.05, n)).to(dev) for _ in range(3)]
    n_node = n ** 3 + 1
    rec = torch.randn((2, ncl, n_node, 4), dtype=torch.float32, device=dev)
    ang = np.linspace(0, 2 * np.pi, ncl, endpoint=False)
    origin = np.stack([8 * np.cos(ang), 8 * np.sin(ang), np.zeros(ncl)], 1)
    scl = np.repeat(np.arange(ncl, dtype=np.int32), nstar)
    p = origin[scl] + rng.normal(0, 0.004, (ncl * nstar, 3))
    sx, sy, sz = (torch.from_numpy(np.ascontiguousarray(p[:, k])).to(dev) for k in range(3))
    d_or, d_scl = torch.from_numpy(origin).to(dev), torch.from_numpy(scl).to(dev)
    acc = torch.empty((3, ncl * nstar), dtype=torch.float64, device=dev)
    pot = torch.empty(ncl * nstar, dtype=torch.float64, device=dev)

    def k3():
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.37, sx, sy, sz, d_scl, acc, None)

    def k3p():
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.37, sx, sy, sz, d_scl, acc, pot)
    for name, fn, ncomp, var in (("k3_interp_c4", k3, 3, 2), ("k3_interp_c4_with_potential", k3p, 4, 2),
                                 ("k3_interp_c4_minb2", k3, 3, 0), ("k3_interp_c4_minb3", k3, 3, 1)):
        ctx.debug_set("interp_variant", var)
        med, best = timeit(fn, flush=flush)
        ctx.debug_set("interp_variant", 2)
        alg = ncl * (nstar * (24 + 8 * ncomp) + 2 * 3 * 4 * n ** 3)  # 3-component planes
        moved = ncl * (nstar * (24 + 4 + 8 * ncomp) + 2 * 16 * n ** 3)  # records are float4: 16 B per node per snapshot
        out[name] = dict(ms_median=med, ms_best=best, algorithmic_bytes=alg, gbs=alg / med / 1e6,
                         frac_of_hbm_peak=alg / med / 1e6 / HBM_PEAK, bytes_if_every_record_read_once=moved,
                         stars=ncl * nstar, l2="flushed before every launch")

    # same launch with the stars spread uniformly over each grid: every record is really fetched from HBM
    # (with cluster-like concentration above, only the central ~1% of each grid is touched and the planes stay in L2)
    pu = origin[scl] + rng.uniform(-0.05, 0.05, (ncl * nstar, 3))
    ux, uy, uz = (torch.from_numpy(np.ascontiguousarray(pu[:, k])).to(dev) for k in range(3))

    def k3u():
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.37, ux, uy, uz, d_scl, acc, None)
    med, best = timeit(k3u, flush=flush)
    moved = ncl * (nstar * (24 + 4 + 24) + 2 * 16 * n ** 3)
    out["k3_interp_c4_uniform_stars"] = dict(ms_median=med, ms_best=best, bytes_moved_min=moved, gbs_moved=moved / med / 1e6,
                                             frac_of_hbm_peak_moved=moved / med / 1e6 / HBM_PEAK,
                                             algorithmic_bytes=ncl * (nstar * 48 + 2 * 3 * 4 * n ** 3),
                                             gbs=ncl * (nstar * 48 + 2 * 3 * 4 * n ** 3) / med / 1e6)

    # ---- K4 ----
    from oc_nbody_b200.synthetic import make_plummer_cluster
    for name, nseg, npc in (("k4_self_gravity_n65536", 1, 65536), ("k4_self_gravity_256x4096", 256, 4096),
                            ("k4_self_gravity_n1024", 1, 1024)):
        pos_pc, _, mass = make_plummer_cluster(npc)
        pos = np.concatenate([pos_pc * 1e-3 + origin[k % ncl][:, None] for k in range(nseg)], axis=1)
        m = np.tile(mass, nseg)
        seg = np.arange(nseg + 1, dtype=np.int64) * npc
        d_pos, d_m = torch.from_numpy(np.ascontiguousarray(pos)).to(dev), torch.from_numpy(m).to(dev)
        a = torch.empty((3, nseg * npc), dtype=torch.float64, device=dev)
        eps2 = (0.01e-3) ** 2

        def k4():
            ctx.self_gravity(d_pos, d_m, eps2, G_KPC_KMS_MYR, a, None, seg_offsets=seg if nseg > 1 else None)
        med, best = timeit(k4)
        inter = float(nseg) * npc * npc
        out[name] = dict(ms_median=med, ms_best=best, interactions=inter, ginter_s=inter / med / 1e6,
                         pct_fp32_peak=100 * 20 * inter / med / 1e9 / nominal, note="pack + kernel per call (finish fused into the kernel)")

    # ---- configs[3] as a whole: one BRIDGE step of 256 clusters x 4096 stars, each kicked by its own 32^3 grid ----
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR
    pos_pc, vel1, mass1 = make_plummer_cluster(nstar)
    bpos = torch.from_numpy(np.ascontiguousarray(np.concatenate([pos_pc * 1e-3 + origin[k][:, None] for k in range(ncl)], axis=1))).to(dev)
    bvel = torch.from_numpy(np.ascontiguousarray(np.tile(vel1, (1, ncl)))).to(dev)
    bm = torch.from_numpy(np.tile(mass1, ncl)).to(dev)
    bseg = np.arange(ncl + 1, dtype=np.int64) * nstar
    a_t = torch.empty((3, ncl * nstar), dtype=torch.float64, device=dev)
    a_s = torch.empty((3, ncl * nstar), dtype=torch.float64, device=dev)
    ctx.self_gravity(bpos, bm, (0.01e-3) ** 2, G_KPC_KMS_MYR, a_s, None, seg_offsets=bseg)

    def batch_step(dt=0.1):
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.37, bpos[0], bpos[1], bpos[2], d_scl, a_t, None)
        ctx.kick(bvel, a_t, 0.5 * dt)
        ctx.kick(bvel, a_s, 0.5 * dt)
        ctx.drift(bpos, bvel, dt, KMS_TO_KPC_PER_MYR)
        ctx.self_gravity(bpos, bm, (0.01e-3) ** 2, G_KPC_KMS_MYR, a_s, None, seg_offsets=bseg)
        ctx.kick(bvel, a_s, 0.5 * dt)
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.38, bpos[0], bpos[1], bpos[2], d_scl, a_t, None)
        ctx.kick(bvel, a_t, 0.5 * dt)
    med, best = timeit(batch_step, iters=10, warm=2)
    out["bridge_step_c4_256_clusters_x_4096_stars"] = dict(
        ms_median=med, ms_best=best, stars=ncl * nstar,
        note="K3 batched (per-cluster 32^3 grids, 2 snapshots) + K5 + K4 batched (256 segments) + K5 + K3 + K5, device resident")

    # ---- BRIDGE step ----
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.units import units
    center = np.array([8.0, 0.0, 0.0])
    nn = 16
    fld = gizmo_field(dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05,
                           grid_resolution=0.05 / nn), [], build=False, ctx=ctx) if False else None
    del fld

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
    for name, nst in (("bridge_step_c1_1k_stars", 1024), ("bridge_step_c3_65k_stars", 65536)):
        pos_pc, vel, mass = make_plummer_cluster(nst)
        for graph in (False, True):
            cl = cluster_code(mass, pos_pc * 1e-3 + center[:, None], vel, softening_pc=0.01, ctx=ctx)
            fake.evolve_model(0.0 | units.Myr)
            system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=graph)
            system.add_system(cl, (fake,))
            system.add_system(fake)
            state = {"t": 0.0}

            def step():
                state["t"] += 0.1
                system.evolve_model(state["t"] | units.Myr, timestep=0.1 | units.Myr)
            l0 = ctx.launch_count()
            med, best = timeit(step, iters=20, warm=4)
            out[name + ("_cuda_graph" if graph else "")] = dict(
                ms_median=med, ms_best=best, kernel_launches_per_step=(ctx.launch_count() - l0) / 24.0,
                graph_replays=system.graph_replays,
                note="K(dt/2) D(dt) K(dt/2), device-resident, no host copies; self-gravity evaluated once per step" +
                     ("; whole step replayed as one CUDA graph launch (library launch counter only sees the eager steps)" if graph else ""))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_extra.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

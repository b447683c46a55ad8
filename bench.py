#!/usr/bin/env python
"""bench.py — headline benchmark of the oceanic gravity hot path on B200.

Metric (BASELINE.json): pairwise gravitational interactions per second, reported in G/s, of the softened
direct-sum field build (K1), with the fraction of FP32 peak at 20 flop per interaction.

Workload
  N = 1 : BASELINE.json configs[1] — 64^3 (+1 origin) Cartesian grid from a synthetic 10M-particle
          GIZMO-format snapshot, Plummer-softened direct sum, one B200.  A step is one pass of the field build
          over one snapshot (2.62e12 interactions).
  N > 1 : STRONG scaling of that same job (north_star's target): the one 10M-particle snapshot is split over the ranks
          (rank p takes every N-th particle), every rank computes the full-grid partial field of its shard, then one
          NCCL all-reduce (sum, fp64, 3*(64^3+1)) over NVLink and the frame subtraction.  `--workload c5s` is the
          64^3 x 1e8 job north_star names for the 8-GPU target, `--workload c5` configs[4] (128^3 x 1e8);
          `--scaling weak` keeps the particle count per rank fixed instead.

  value : whole-job interactions/s with inputs (FP32 recentred sources/targets) resident in HBM.
  e2e   : same metric through the C-ABI host call ocg_field_build_host (N=1) / the device calls fed from pinned
          host tensors (N>1): FP64 host buffers in, H2D, recentre, K1, [all-reduce], K1b, D2H of the field.
  parity_check : rows of the (all-reduced) raw field against the FP64 oracle over ALL shards, strict metric; the run
          fails when the raw field misses 1e-5.
  roofline_k4, roofline_k3, bridge_step, cpu_baseline.bridge_step (N=1): K4 at N = 65 536 (FP32 bound), K3 at configs[3]
  (HBM bound) and the BRIDGE step, GPU and CPU.
  --impl reference : the reference's CPU path (the FP64 OpenMP port in oracle/, ALL host threads, also under torchrun)
          on a bounded sample of the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 20.0  # GPU-Gems-3 ch.31 convention, acc only
# (grid, sources on this GPU) -> DRAM bytes moved by one launch of the streaming kernel (ncu --set full, profiles/)
NCU_DRAM_BYTES_PER_LAUNCH = {
    # profiles/r02_ncu_dram_bench_full.csv: 308 637 696 B read (200 MB of tiles once — streamed in 32 MB passes that stay
    # in L2: lts__t_bytes 18.3 GB — + targets + the partial slots of the shared rows read back) + 107 409 408 B written
    # (partial slots + the FP64 field): 0.006 % of HBM bandwidth over the 949 ms launch.  History: r01 917 MB (667 MB of
    # chunk partials); stream-K without passes 9.14 GB (every row re-fetched the tiles from HBM,
    # profiles/r02_ncu_dram_before_passes.csv)
    (64, 10000000): 308637696 + 107409408,
}
G_KPC = 4.398600413517813e-09
CENTER = np.array([8.0, 0.0, 0.0])
HALF = 0.6
STAR_SOFT, DARK_SOFT = 11.2e-3, 112.0e-3  # kpc (test_options:25-26)


# ----------------------------------------------------------------------------------- inputs ----
def make_sources(n, seed):
    """FP64 (pos [n,3], mass [n], plummer eps [n]) in the reference's source order star|dark|gas."""
    from oc_nbody_b200.synthetic import make_snapshot
    snap = make_snapshot(n, seed=seed)
    pos = np.concatenate([snap[s]["position"] for s in ("star", "dark", "gas")])
    mass = np.concatenate([snap[s]["mass"] for s in ("star", "dark", "gas")])
    soft = np.concatenate([np.full(len(snap["star"]["mass"]), STAR_SOFT), np.full(len(snap["dark"]["mass"]), DARK_SOFT),
                           2.8e-3 * snap["gas"]["smooth.length"]]) / 2.8  # Plummer-equivalent epsilon of the spline h
    return np.ascontiguousarray(pos), np.ascontiguousarray(mass), np.ascontiguousarray(soft)


def make_targets(n_grid):
    from oc_nbody_b200.grid_cartesian import grid
    g = grid(HALF, HALF, HALF, HALF / n_grid)
    g.gen_evolved_grid(CENTER)
    return g


# ------------------------------------------------------------------------------ clock sampler ----
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU baseline ----
def cpu_sample_rate(tgt_pos, src_pos, src_mass, src_eps, seconds):
    """Time the oracle's vectorised FP64 direct sum (all host threads) on a bounded sample of the workload:
    all the sources handed in x a strided subset of the grid targets sized for ~`seconds` of CPU work.
    Returns (interactions/s, description, threads)."""
    import oracle
    e2 = src_eps * src_eps
    n_s = src_pos.shape[0]
    cal_t = np.ascontiguousarray(tgt_pos[:: max(1, tgt_pos.shape[0] // 2048)][:2048])
    oracle.field_direct_fast(src_pos[:20000], src_mass[:20000], e2[:20000], cal_t[:64], G_KPC)  # spin the threads up
    t0 = time.perf_counter()
    oracle.field_direct_fast(src_pos, src_mass, e2, cal_t, G_KPC)
    rate0 = n_s * cal_t.shape[0] / max(time.perf_counter() - t0, 1e-4)
    n_t = int(min(tgt_pos.shape[0], max(512, rate0 * seconds / n_s)))
    tp = np.ascontiguousarray(tgt_pos[:: max(1, tgt_pos.shape[0] // n_t)][:n_t])
    t0 = time.perf_counter()
    oracle.field_direct_fast(src_pos, src_mass, e2, tp, G_KPC)
    dt = time.perf_counter() - t0
    desc = "%d of the grid targets x %d of the snapshot particles, FP64 OpenMP direct sum (oracle/ocg_oracle.c), %.1f s" % (
        tp.shape[0], n_s, dt)
    return n_s * tp.shape[0] / dt, desc, oracle.num_threads()


def reference_time_interp_cost(n_points=16 ** 3 + 1, n_snap=9):
    """The reference's OWN per-step time interpolation, run as-is with scipy (gizmo_interface.py:587-620): one
    splrep per grid point and component at start-up, one splev per grid point and component on every evolve_model.
    configs[0] size (16^3+1 points, 9 snapshots as test_options:20-22), single process (the reference spreads the
    same calls over multiprocessing.Pool(ncpu))."""
    from scipy import interpolate
    rng = np.random.default_rng(1776)
    times = 23.0 * np.arange(n_snap)
    series = rng.normal(0.0, 1e-2, (3, n_snap, n_points))
    t0 = time.perf_counter()
    tck = [[interpolate.splrep(times, series[c][:, i]) for i in range(n_points)] for c in range(3)]
    setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    for c in range(3):
        for i in range(n_points):
            float(interpolate.splev(57.3, tck[c][i]))
    step = time.perf_counter() - t0
    return {"n_grid_points": n_points, "n_snapshots": n_snap, "splrep_setup_s": setup, "splev_per_evolve_model_s": step,
            "calls_per_step": 3 * n_points,
            "note": "reference mechanism run with scipy as-is, 1 process; here evolve_model is O(1) and the blend is fused into K3"}


def host_threads():
    """Every host core this process may use (torchrun exports OMP_NUM_THREADS=1: not what a CPU baseline wants)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def sample_sources(args):
    """The bounded source sample the CPU legs use: the first 1e6 particles of an equally shaped snapshot."""
    return make_sources(min(args.n_src_total, 1000000), seed=1776)


def run_reference(args, world):
    """--impl reference: the reference's CPU path = the oracle port (the reference itself cannot be imported:
    SURVEY §0.3), ALL host threads (also under torchrun), on the same config as the repo arm at this N; each step is a
    bounded sample of that workload.  Rank 0 alone runs; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = oracle.set_num_threads(host_threads())
    g = make_targets(args.grid)
    pos, mass, eps = sample_sources(args)
    rates, secs, desc = [], [], ""
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r, desc, threads = cpu_sample_rate(g.evolved_grid, pos, mass, eps, per_step)
        if i >= args.warmup:
            rates.append(r)
            secs.append(time.perf_counter() - t0)
    val = float(np.mean(rates)) / 1e9
    inter_step = float(len(g)) * args.n_src_total
    line = {
        "impl": "reference", "metric": "pairwise_grav_interactions_per_sec", "value": val, "unit": "G/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "G/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "G/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "ms_per_full_step_extrapolated": inter_step / (val * 1e9) * 1e3,
        "reference_time_interp": reference_time_interp_cost(),
        "note": "a step here is the bounded sample named in cpu_baseline.sample (ms_per_step is its wall time); the rate of "
                "a direct sum is size independent, ms_per_full_step_extrapolated applies it to the whole workload",
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    name = {"c2": "configs[1]", "c5": "configs[4]", "c5s": "64^3 x 1e8 (north_star strong-scaling target)"}[args.workload]
    return {"workload": "%s: %d^3+1 grid targets x %d synthetic snapshot particles in total, Plummer-softened direct-sum field "
                        "build" % (name, args.grid, args.n_src_total),
            "grid": args.grid, "n_targets": args.grid ** 3 + 1, "n_sources_total": args.n_src_total,
            "softening": "plummer, per-source epsilon", "scaling": args.scaling,
            "parallelism": "source-sharded (every n-th particle) + NCCL all-reduce(fp64) over the ranks; 1 rank = single GPU",
            "l2_policy": ("inputs larger than L2 (%.0f MB of source tiles per GPU vs 126 MB L2)" % (args.n_src_total * 20 / 1e6 / n_gpus)
                          if args.n_src_total * 20 / n_gpus > 2 * 126e6 or (n_gpus == 1 and args.n_src_total * 20 > 126e6) else
                          "L2 flushed before every timed step (a 256 MB buffer is overwritten, outside the step's CUDA events): "
                          "%.0f MB of source tiles per GPU would fit the 126 MB L2" % (args.n_src_total * 20 / 1e6 / n_gpus))}


def cpu_bridge_step_ms(n_stars, steps):
    """BRIDGE step on the host: the CPU pipeline assembled from the oracle (tests/test_gpu_fullsize.py::
    test_config0_full_bridge_run) — trilinear + linear-in-time tidal kick from a 16^3 grid, FP64 OpenMP direct-sum
    self-gravity, KDK leapfrog — all host threads."""
    import oracle
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR
    from oc_nbody_b200.units import G_KPC_KMS_MYR
    nn = 16
    rng = np.random.default_rng(7)
    ax = np.linspace(-0.05, 0.05, nn)
    recs = [oracle.pack_planes(rng.normal(0, 1e-3, (3, nn ** 3 + 1))) for _ in range(2)]
    pos_pc, vel, mass = make_plummer_cluster(n_stars)
    x, v = pos_pc * 1e-3 + CENTER[:, None], vel.copy()
    eps2, dt = (0.01e-3) ** 2, 0.1

    def tidal(xx, w):
        return oracle.grid_interp([ax, ax, ax], CENTER[None], recs[0], recs[1], w, xx[0], xx[1], xx[2])
    ts = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        v = oracle.kick(v, tidal(x, 0.1), 0.5 * dt)
        v = oracle.kick(v, oracle.self_gravity(x, mass, eps2, G_KPC_KMS_MYR), 0.5 * dt)
        x = oracle.drift(x, v, dt, KMS_TO_KPC_PER_MYR)
        v = oracle.kick(v, oracle.self_gravity(x, mass, eps2, G_KPC_KMS_MYR), 0.5 * dt)
        v = oracle.kick(v, tidal(x, 0.2), 0.5 * dt)
        if i:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def bridge_step_times(ctx):
    """BRIDGE step time (the second half of BASELINE.json's metric): K(dt/2) D(dt) K(dt/2) of a Plummer cluster kicked by
    a 16^3 grid field interpolated in space and time, device resident, eager launches and one-launch CUDA-graph replay.
    configs[0] shape (1 024 stars) and configs[2] shape (65 536 stars)."""
    import torch
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import units

    class _Snap(object):
        snapshot = {"index": 0, "time": 0.0}
    nn = 16
    fld = gizmo_field(dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / nn),
                      [_Snap(), _Snap()], time_in_Myr=[0.0, 23.0], build=False, ctx=ctx)
    rng = np.random.default_rng(7)
    tid = rng.normal(0, 1e-3, (2, 3, nn ** 3 + 1))
    fld.set_snapshot_fields(tid[:, 0], tid[:, 1], tid[:, 2])
    fld.evolve_grid(CENTER)
    out = {"unit": "ms", "dt_myr": 0.1, "grid": "16^3, 2 snapshots, linear in time"}
    for nst in (1024, 65536):
        pos_pc, vel, mass = make_plummer_cluster(nst)
        for graph in (False, True):
            cl = cluster_code(mass, pos_pc * 1e-3 + CENTER[:, None], vel, softening_pc=0.01, ctx=ctx)
            fld.evolve_model(0.0 | units.Myr)
            system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=graph)
            system.add_system(cl, (fld,))
            system.add_system(fld)
            t, ts = 0.0, []
            for i in range(24):
                t += 0.1
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                system.evolve_model(t | units.Myr, timestep=0.1 | units.Myr)
                e1.record()
                torch.cuda.synchronize()
                if i >= 4:
                    ts.append(e0.elapsed_time(e1))
            out["%d_stars_%s" % (nst, "cuda_graph" if graph else "eager")] = float(np.median(ts))
    # the reference-algorithm modes: ph4's Hermite scheme as the drift (K6, two acc+jerk evaluations
    # per step) and the kNN(150) + RBF-PHS interpolant as the kick (K7, two kicks per step)
    rbf = gizmo_field(dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / nn,
                           space_interpolation="rbf"), [_Snap(), _Snap()], time_in_Myr=[0.0, 23.0], build=False, ctx=ctx)
    rbf.set_snapshot_fields(tid[:, 0], tid[:, 1], tid[:, 2])
    rbf.evolve_grid(CENTER)
    for key, nst, field, kw, graph in (("1024_stars_hermite_cuda_graph", 1024, fld, dict(integrator="hermite"), True),
                                       ("65536_stars_hermite_cuda_graph", 65536, fld, dict(integrator="hermite"), True),
                                       ("1024_stars_rbf_kick_eager", 1024, rbf, {}, False)):
        pos_pc, vel, mass = make_plummer_cluster(nst)
        cl = cluster_code(mass, pos_pc * 1e-3 + CENTER[:, None], vel, softening_pc=0.01, ctx=ctx, **kw)
        field.evolve_model(0.0 | units.Myr)
        system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=graph)
        system.add_system(cl, (field,))
        system.add_system(field)
        t, ts = 0.0, []
        for i in range(12):
            t += 0.1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            system.evolve_model(t | units.Myr, timestep=0.1 | units.Myr)
            e1.record()
            torch.cuda.synchronize()
            if i >= 4:
                ts.append(e0.elapsed_time(e1))
        out[key] = float(np.median(ts))
    return out


def k3_roofline(ctx):
    """Third roofline entry (driver-run), the HBM-bound kernel of the path: K3 (trilinear gather + time blend) at configs[3]
    — 256 clusters x 4 096 stars, 32^3 grids, two snapshots — against the measured copy bandwidth of MEASURED_PEAKS.json.
    Algorithmic bytes per launch: per star 24 B in + 24 B out, every 3-component plane once.  L2 flushed before
    every launch (a 256 MB buffer is overwritten)."""
    import torch
    dev = torch.device("cuda", ctx.device)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_kind = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool)"
    except Exception:  # noqa: BLE001
        peak, peak_kind = 6454.0, "fallback: the pool's measured copy bandwidth recorded in round 1 (MEASURED_PEAKS.json absent)"
    ncl, nstar, n = 256, 4096, 32
    rng = np.random.default_rng(7)
    nodes = [torch.from_numpy(np.linspace(-0.05, 0.05, n)).to(dev) for _ in range(3)]
    rec = torch.randn((2, ncl, n ** 3 + 1, 4), dtype=torch.float32, device=dev)
    ang = np.linspace(0, 2 * np.pi, ncl, endpoint=False)
    origin = np.stack([8 * np.cos(ang), 8 * np.sin(ang), np.zeros(ncl)], 1)
    scl = np.repeat(np.arange(ncl, dtype=np.int32), nstar)
    p = origin[scl] + rng.normal(0, 0.004, (ncl * nstar, 3))
    sx, sy, sz = (torch.from_numpy(np.ascontiguousarray(p[:, k])).to(dev) for k in range(3))
    d_or, d_scl = torch.from_numpy(origin).to(dev), torch.from_numpy(scl).to(dev)
    acc = torch.empty((3, ncl * nstar), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(23):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.grid_interp((n, n, n), nodes, d_or, rec[0], rec[1], 0.37, sx, sy, sz, d_scl, acc, None)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    alg = ncl * (nstar * 48 + 2 * 3 * 4 * n ** 3)
    ach = alg / ms / 1e6
    return {"bound": "hbm", "kernel": "K3 grid_interp_kernel, configs[3]: 256 clusters x 4096 stars, 32^3 grids, 2 snapshots",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_kind": peak_kind,
            "ms_per_launch": ms, "algorithmic_bytes_per_launch": alg, "l2": "flushed before every launch"}


def k4_roofline(ctx, nominal):
    """Second roofline entry (driver-run): K4 cluster self-gravity at configs[2]'s N = 65 536 (4.29e9 interactions per
    evaluation), whole evaluation (pack + stream-K kernel, two launches) timed with CUDA events; `frac_kernel_alone` from the
    library's own events around the kernel.  The 1.3 MB of source tiles are L2-resident by design (every row streams them)."""
    import torch
    from oc_nbody_b200.synthetic import make_plummer_cluster
    n = 65536
    pos_pc, _, mass = make_plummer_cluster(n)
    dev = torch.device("cuda", ctx.device)
    d_pos = torch.from_numpy(np.ascontiguousarray(pos_pc * 1e-3 + CENTER[:, None])).to(dev)
    d_m = torch.from_numpy(mass).to(dev)
    acc = torch.empty((3, n), dtype=torch.float64, device=dev)
    ts, ks = [], []
    ctx.set_kernel_timing(True)
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.self_gravity(d_pos, d_m, (0.01e-3) ** 2, G_KPC, acc)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
            ks.append(ctx.last_direct_kernel_ms())
    ctx.set_kernel_timing(False)
    ms, kms = float(np.median(ts)), float(np.median(ks))
    inter = float(n) * n
    ach = FLOP_PER_INTERACTION * inter / (ms * 1e-3) / 1e12
    return {"bound": "fp32", "kernel": "K4 ocg_self_gravity, N = 65536 (configs[2]): pack + direct_sum_tp_kernel (wide rows, stream-K, finish fused)",
            "achieved": ach, "peak": nominal, "unit": "TFLOP/s", "frac": ach / nominal, "traffic": None,
            "ms_per_evaluation": ms, "kernel_ms": kms, "frac_kernel_alone": FLOP_PER_INTERACTION * inter / (kms * 1e-3) / 1e12 / nominal,
            "interactions_per_launch": inter, "flop_per_interaction": FLOP_PER_INTERACTION}


# ------------------------------------------------------------------------------------ main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5", "c5s"],
                    help="c2 = configs[1] (64^3 x 1e7); c5 = configs[4] (128^3 x 1e8); c5s = 64^3 x 1e8")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the workload's total particle count is split over the ranks; weak: every rank "
                         "holds that many particles")
    ap.add_argument("--n-src", type=float, default=None, help="override the particle count (debug)")
    ap.add_argument("--grid", type=int, default=None, help="override grid nodes per axis (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the BRIDGE-step and K4 extras (profiling runs)")
    ap.add_argument("--variant", default="auto", help="auto | kernel shape id (debug)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.grid = args.grid or (128 if args.workload == "c5" else 64)
    base = int(args.n_src or (1e7 if args.workload == "c2" else 1e8))
    args.n_src_total = base * world if args.scaling == "weak" else base
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: oc_nbody_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from oc_nbody_b200 import Context
    ctx = Context(local_rank)
    if args.variant != "auto":
        ctx.debug_set("direct_variant", int(args.variant))
    dev = torch.device("cuda", local_rank)

    g = make_targets(args.grid)
    n_tgt = len(g)
    # strong scaling: ONE snapshot (the N = 1 workload), rank p takes every world-th particle starting at p, so the shards
    # are statistically alike (equal near-field work); weak scaling: every rank generates its own snapshot
    if args.scaling == "weak":
        pos, mass, eps = make_sources(base, seed=1776 + rank)
    else:
        pos, mass, eps = (np.ascontiguousarray(a[rank::world]) for a in make_sources(base, seed=1776))
    n_src = pos.shape[0]
    inter_rank = float(n_src) * n_tgt
    inter_total = float(args.n_src_total) * n_tgt

    # pinned host buffers (e2e path) and resident device inputs (value path)
    h_pos, h_mass, h_eps = (torch.from_numpy(a).pin_memory() for a in (pos, mass, eps))
    h_tgt = torch.from_numpy(np.ascontiguousarray(g.evolved_grid)).pin_memory()
    d_src = torch.empty((n_src, 4), dtype=torch.float32, device=dev)
    d_eps = torch.empty(n_src, dtype=torch.float32, device=dev)
    d_tgt = torch.empty((n_tgt, 4), dtype=torch.float32, device=dev)
    ctx.recentre_f64(h_pos.to(dev), h_mass.to(dev), CENTER, d_src)
    ctx.cast_f64_f32(h_eps.to(dev), d_eps)
    ctx.recentre_f64(h_tgt.to(dev), None, CENTER, d_tgt)
    acc = torch.empty((3, n_tgt), dtype=torch.float64, device=dev)
    ctx.set_source_shards(world)  # the FP64 near set is sized for the whole build, 1/world of it on every rank
    torch.cuda.synchronize()

    def step_resident(subtract=True):
        ctx.field_direct(d_src, d_eps, d_tgt, 0, G_KPC, acc)
        if world > 1:
            dist.all_reduce(acc)
        if subtract:
            ctx.frame_subtract(acc, g.origin_row)

    def step_e2e():
        if world == 1:
            return ctx.field_build_host(h_pos.numpy(), h_mass.numpy(), h_eps.numpy(), h_tgt.numpy(), CENTER, g.origin_row, 0,
                                        G_KPC)
        p64, m64, e64 = h_pos.to(dev, non_blocking=True), h_mass.to(dev, non_blocking=True), h_eps.to(dev, non_blocking=True)
        t64 = h_tgt.to(dev, non_blocking=True)
        ctx.recentre_f64(p64, m64, CENTER, d_src)
        ctx.cast_f64_f32(e64, d_eps)
        ctx.recentre_f64(t64, None, CENTER, d_tgt)
        ctx.field_direct(d_src, d_eps, d_tgt, 0, G_KPC, acc)
        dist.all_reduce(acc)
        ctx.frame_subtract(acc, g.origin_row)
        return acc.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # timing rule: inputs larger than L2, or L2 flushed between timed iterations.  One rank's source tiles are
    # n_src * 20 B: above the L2 at N = 1 on configs[1] (200 MB), below it once the snapshot is split over more ranks
    need_flush = not (n_src * 20 > 2 * 126e6 or (world == 1 and n_src * 20 > 126e6))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None

    def timed(fn, steps, kernel_ms=None):
        barrier()
        if flush_buf is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
                if kernel_ms is not None:
                    kernel_ms.append(ctx.last_direct_kernel_ms())
            e1.record()
            barrier()
            total = e0.elapsed_time(e1)
        else:  # per-step events, the flush between them untimed; the steps' device times are added up
            evs = []
            for _ in range(steps):
                flush_buf.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                evs.append((e0, e1))
                if kernel_ms is not None:
                    kernel_ms.append(ctx.last_direct_kernel_ms())
            barrier()
            total = sum(a.elapsed_time(b) for a, b in evs)
        ms = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- parity self-check (untimed): rows of the (all-reduced) raw field against the FP64 oracle over ALL shards ----
    parity = None
    step_resident(subtract=False)
    rows = np.unique(np.concatenate([np.linspace(0, n_tgt - 2, 16).astype(np.int64), [g.origin_row]]))
    got_rows = acc[:, torch.from_numpy(rows).to(dev)].cpu().numpy()
    if rank == 0 and not args.no_cpu_baseline:
        import oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from util import rel_err
        oracle.set_num_threads(host_threads())
        fp, fm, fe = make_sources(base, seed=1776) if (world > 1 and args.scaling == "strong") else (pos, mass, eps)
        if args.scaling == "weak" and world > 1:
            parts = [make_sources(base, seed=1776 + r) for r in range(world)]
            fp, fm, fe = (np.concatenate([q[k] for q in parts]) for k in range(3))
        s32 = oracle.recentre(fp, fm, CENTER)
        t32 = oracle.recentre(g.evolved_grid[rows], None, CENTER)
        ref = oracle.field_direct(s32, fe.astype(np.float32), t32, oracle.KERNEL_PLUMMER, G_KPC)
        o = int(np.nonzero(rows == g.origin_row)[0][0])
        keep = np.arange(len(rows)) != o
        parity = {"rows": int(len(rows)), "sources": int(s32.shape[0]), "tolerance": 1e-5,
                  "metric": "max_c |a - a_ref| / max(|a_ref,c|, 1e-3 ||a_ref||), FP64 oracle on the same FP32-rounded inputs",
                  "raw_field": rel_err(got_rows, ref),
                  "tidal_residual": rel_err((got_rows - got_rows[:, o:o + 1])[:, keep], (ref - ref[:, o:o + 1])[:, keep])}
        parity["ok"] = bool(parity["raw_field"] <= 1e-5 and parity["tidal_residual"] <= 1e-5)
        del s32, fp, fm, fe

    # ---- resident (device-timed) ----
    for _ in range(args.warmup):
        step_resident()
    ctx.set_kernel_timing(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    kernel_ms = []
    total_ms = timed(step_resident, args.steps, kernel_ms)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_kernel_timing(False)
    ms_per_step = total_ms / args.steps
    value = inter_total / (ms_per_step * 1e-3) / 1e9
    result_check = float(acc[:, g.origin_row].abs().max().item())

    # ---- end to end (host buffers) ----
    for _ in range(1):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps
    e2e_value = inter_total / (e2e_ms * 1e-3) / 1e9
    h2d = n_src * (24 + 8 + 8) + n_tgt * 24
    d2h = n_tgt * 24

    # ---- roofline of the dominant kernel (direct_sum_tp_kernel), FP32 pipe ----
    k_ms = float(np.mean(kernel_ms))
    achieved = FLOP_PER_INTERACTION * inter_rank / (k_ms * 1e-3) / 1e12
    nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    ffma = ctx.probe_throughput(0)
    ffma2 = ctx.probe_throughput(1)
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this same command
    # (dram__bytes_read.sum + dram__bytes_write.sum, profiles/); null when no capture exists for this configuration.
    # traffic_model is what the launch must move: source tiles + targets once, the FP64 field, partial slots written + read once.
    traffic = NCU_DRAM_BYTES_PER_LAUNCH.get((args.grid, n_src))
    roofline = {"bound": "fp32", "achieved": achieved, "peak": nominal, "unit": "TFLOP/s", "frac": achieved / nominal,
                "traffic": traffic, "traffic_model": ctx.last_direct_traffic_model(),
                "kernel": "direct_sum_tp_kernel (target-paired, mass-folded tiles, FP32 runs of 64 sources folded into FP64, folds staggered between the warps of a scheduler; K1 fast set)",
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                "bound_note": "FP32 FMA-pipe bound (north_star: no tensor cores; HBM traffic negligible)",
                "peak_kind": "nominal FP32: %d SM x 128 lanes x 2 flop x %.3f GHz (MEASURED_PEAKS.json has no FP32 entry)" % (
                    ctx.sm_count, ctx.sm_clock_khz / 1e6),
                "peak_measured_ffma": ffma, "peak_measured_ffma2": ffma2, "frac_of_measured": achieved / max(ffma, ffma2),
                "flop_per_interaction": FLOP_PER_INTERACTION,
                "interactions_per_launch": inter_rank}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    extras = world == 1 and not args.no_extras
    bridge = bridge_step_times(ctx) if extras else None
    roofline_k4 = k4_roofline(ctx, nominal) if extras else None
    roofline_k3 = None
    if extras:
        try:
            roofline_k3 = k3_roofline(ctx)
        except Exception as exc:  # noqa: BLE001  (an extra must never cost the headline line)
            roofline_k3 = {"error": str(exc)[:200]}
    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        oracle.set_num_threads(host_threads())
        sp, sm, se = sample_sources(args)
        r, desc, threads = cpu_sample_rate(g.evolved_grid, sp, sm, se, 15.0)
        cpu = {"value": r / 1e9, "unit": "G/s", "cores": threads, "kind": "port", "sample": desc}
        if extras:
            cpu["bridge_step"] = {"unit": "ms", "cores": threads, "kind": "port",
                                  "1024_stars": cpu_bridge_step_ms(1024, 8), "65536_stars": cpu_bridge_step_ms(65536, 2),
                                  "what": "oracle pipeline: trilinear + linear-in-time tidal kick (16^3 grid), FP64 OpenMP direct-sum "
                                          "self-gravity, KDK leapfrog (the CPU side of tests/test_gpu_fullsize.py::test_config0_full_bridge_run)"}
    line = {
        "metric": "pairwise_grav_interactions_per_sec", "value": value, "unit": "G/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "dtype_note": "f32 pair arithmetic, f64 per-target accumulation", "data": "synthetic",
        "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "G/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "api": "ocg_field_build_host (C ABI, host buffers)" if world == 1 else
                       "pinned host tensors -> ocg_recentre_f64/ocg_field_direct/all_reduce/ocg_frame_subtract -> host, every rank"},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_k4": roofline_k4,
        "roofline_k3": roofline_k3,
        "cpu_baseline": cpu,
        "parity_check": parity,
        "pct_fp32_peak": 100.0 * achieved / nominal,
        "bridge_step": bridge,
        "origin_row_abs_max": result_check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: parity self-check failed: %r" % (parity,))


if __name__ == "__main__":
    main()

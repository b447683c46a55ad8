"""N-GPU BRIDGE step: stars block-sharded over the ranks, positions NCCL-all-gathered at every self-gravity evaluation
(K4 target-sharded), tidal kick (K3) and leapfrog (K5) local.  Verifies the sharded trajectory against the single-GPU one
computed on every rank, then times the step.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bridge_multi.py [--stars 65536] [--steps 5]            (N = 1 works too)
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stars", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--graph", action="store_true", help="replay the sharded step (NCCL all-gathers included) as one CUDA graph")
    ap.add_argument("--integrator", default="leapfrog", choices=("leapfrog", "hermite"),
                    help="hermite: K6 (acc + jerk) target-sharded, positions AND velocities all-gathered per evaluation")
    ap.add_argument("--exchange", default="peer", choices=("peer", "nccl"),
                    help="peer: position gather fused into the tile pack over NVLink peer memory (ocg_self_gravity_sharded); nccl: all_gather + copy + pack")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oc_nbody_b200 import Context
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code, sharded_cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_plummer_cluster, make_snapshot
    from oc_nbody_b200.units import units
    ctx = Context(local)
    center = np.array([8.0, 0.0, 0.0])
    snaps = [make_snapshot(200000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / args.grid,
                with_potential=False)
    field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)   # replicated on every rank
    pos_pc, vel, mass = make_plummer_cluster(args.stars)
    pos = pos_pc * 1e-3 + center[:, None]
    dt = 0.1

    def run(code, steps, timed=False):
        field.evolve_grid(center)
        field.evolve_model(0.0 | units.Myr)
        system = Bridge(timestep=dt | units.Myr, use_threading=False, use_cuda_graph=args.graph)
        system.add_system(code, (field,))
        system.add_system(field)
        system.evolve_model(0.0 | units.Myr, timestep=dt | units.Myr)
        system.evolve_model(dt | units.Myr, timestep=dt | units.Myr)  # warm-up step (also validates)
        if args.graph:  # two more: the step that sizes the scratch buffers and the one that captures
            system.evolve_model(2 * dt | units.Myr, timestep=dt | units.Myr)
            system.evolve_model(3 * dt | units.Myr, timestep=dt | units.Myr)
        first = 4 if args.graph else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(first, steps + first):
            system.evolve_model(i * dt | units.Myr, timestep=dt | units.Myr)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        run.replays = system.graph_replays
        return float(ms.item())

    sh = sharded_cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx, integrator=args.integrator, exchange=args.exchange)
    # stage checks: the all-gather reproduces the full arrays; the first sharded force equals the unsharded one
    g_pos, g_vel = (t.cpu().numpy() for t in sh.gather_state())
    gather_ok = bool(np.array_equal(g_pos, pos) and np.array_equal(g_vel, vel))
    ref_code = cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx)
    a_full = ref_code.compute_self_gravity().cpu().numpy()
    a_loc = sh.compute_self_gravity().cpu().numpy()
    force_err = float(np.max(np.abs(a_loc - a_full[:, sh.a:sh.b])) / np.max(np.abs(a_full)))
    ms_sharded = run(sh, args.steps)
    x_sh, v_sh = (t.cpu().numpy() for t in sh.gather_state())
    single = cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx, integrator=args.integrator)
    ms_single = run(single, args.steps)
    x_1, v_1 = single.pos.cpu().numpy(), single.vel.cpu().numpy()
    dx = float(np.max(np.abs(x_sh - x_1)) / np.max(np.abs(x_1 - center[:, None])))
    dv = float(np.max(np.abs(v_sh - v_1)) / np.max(np.abs(v_1)))
    ok = dx < 1e-10 and dv < 1e-10 and gather_ok and force_err < 1e-12
    if getattr(sh, "_peer", False):
        ctx.comm_status()
    print("rank %d: gather_ok %s force_err %.3e dx %.3e dv %.3e" % (rank, gather_ok, force_err, dx, dv), file=sys.stderr)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "n_stars": args.stars, "integrator": args.integrator, "exchange": "peer" if getattr(sh, "_peer", False) else "nccl", "cuda_graph": bool(args.graph), "graph_replays_last_run": getattr(run, "replays", 0), "steps": args.steps, "ms_per_bridge_step_sharded": ms_sharded,
                          "ms_per_bridge_step_single_gpu": ms_single, "max_rel_dx": dx, "max_rel_dv": dv, "match": ok,
                          "allgather_exact_rank0": gather_ok, "first_force_rel_err_rank0": force_err}))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("sharded BRIDGE trajectory differs from the single-GPU one: dx %g dv %g" % (dx, dv))


if __name__ == "__main__":
    main()

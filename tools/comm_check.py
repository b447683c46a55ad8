"""N-rank check of the peer-memory exchange (include/ocg.h ocg_comm_*), run under torch.distributed.run:
  * ocg_comm_allreduce_f64 vs NCCL all-reduce; result bit-identical on every rank (rank-order sum)
  * source-sharded K1 field build (every rank: its share of the snapshot, full grid) + all-reduce vs the FP64 oracle over
    ALL shards (strict metric)
  * ocg_self_gravity_sharded (position gather fused into the pack, over NVLink) vs the single-GPU K4 on every rank
Prints one JSON line on rank 0; exit code 1 on a mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import oracle
    from bench import CENTER, G_KPC, make_sources, make_targets
    from oc_nbody_b200 import Context
    from oc_nbody_b200.distributed import connect_comm, shard_range
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from util import rel_err
    ctx = Context(local)
    connect_comm(ctx, None, window_bytes=8 << 20)
    out = {"n_gpus": world}
    dev = torch.device("cuda", local)
    # ---- all-reduce
    ok = True
    for n in (7, 100000, 700001):  # the last needs 2 window passes
        x = torch.from_numpy(np.random.default_rng(100 + rank).normal(size=n)).to(dev)
        mine, theirs = x.clone(), x.clone()
        ctx.comm_allreduce_f64(mine)
        if world > 1:
            dist.all_reduce(theirs)
        torch.cuda.synchronize()
        err = float((mine - theirs).abs().max() / theirs.abs().max())
        ok &= err < 1e-14
        if world > 1:  # bit-identical across ranks
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            ok &= all(torch.equal(gathered[0], t) for t in gathered)
        out["allreduce_rel_err_n%d" % n] = err
    # ---- source-sharded field build vs the oracle over all shards
    g = make_targets(12)
    pos, mass, eps = make_sources(300000, seed=1776)
    sl = slice(rank, None, world)
    s32 = torch.from_numpy(oracle.recentre(pos[sl], mass[sl], CENTER)).to(dev)
    t32 = oracle.recentre(g.evolved_grid, None, CENTER)
    acc = torch.empty((3, len(g)), dtype=torch.float64, device=dev)
    ctx.set_source_shards(world)
    ctx.field_direct(s32, torch.from_numpy(eps[sl].astype(np.float32)).to(dev), torch.from_numpy(t32).to(dev), 0, G_KPC, acc)
    ctx.set_source_shards(1)
    ctx.comm_allreduce_f64(acc)
    torch.cuda.synchronize()
    ref = oracle.field_direct(oracle.recentre(pos, mass, CENTER), eps.astype(np.float32), t32, 0, G_KPC)
    cond = oracle.field_direct_abs(oracle.recentre(pos, mass, CENTER), eps.astype(np.float32), t32, 0, G_KPC)
    out["field_build_sharded_err"] = rel_err(acc.cpu().numpy(), ref, abs_sum=cond)
    ok &= out["field_build_sharded_err"] <= 1e-5
    # ---- star-sharded K4
    n = 20001
    p, _, m = make_plummer_cluster(n, seed=3)
    x = p * 1e-3 + CENTER[:, None]
    a, b = shard_range(n, rank, world)
    eps2 = (0.01e-3) ** 2
    full = torch.empty((3, n), dtype=torch.float64, device=dev)
    fpot = torch.empty(n, dtype=torch.float64, device=dev)
    d_x, d_m = torch.from_numpy(np.ascontiguousarray(x)).to(dev), torch.from_numpy(m).to(dev)
    ctx.debug_set("direct_variant", 27)
    ctx.self_gravity(d_x, d_m, eps2, G_KPC, full, fpot)
    ctx.debug_set("direct_variant", -1)
    for it in range(3):
        loc = torch.full((3, b - a), float("nan"), dtype=torch.float64, device=dev)
        lpot = torch.full((b - a,), float("nan"), dtype=torch.float64, device=dev)
        ctx.self_gravity_sharded(d_x[:, a:b].contiguous(), d_m, eps2, G_KPC, loc, lpot)
        torch.cuda.synchronize()
        same = bool(torch.equal(loc, full[:, a:b]) and torch.equal(lpot, fpot[a:b]))
        dmax = float((loc - full[:, a:b]).abs().max() / full.abs().max())
        ok &= same or dmax < 1e-12
        out["k4_sharded_bit_identical_iter%d" % it] = same
        out["k4_sharded_rel_diff_iter%d" % it] = dmax
    ctx.comm_status()
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if not out["ok"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()

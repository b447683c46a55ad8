"""CPU, world_size 2, gloo: the multi-GPU partition/reduction logic with the oracle standing in for the GPU
compute step (the real N>1 path runs under `gpurun --gpus N`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 4.3986004135e-09


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oc_nbody_b200.distributed import allgather_particles, field_build_sharded, self_gravity_sharded, shard_range
    from util import grid_targets, random_sources
    rng = np.random.default_rng(99)  # same data on every rank
    src, soft = random_sources(rng, 3001, box=2.0)
    tgt = grid_targets(5)
    row = tgt.shape[0] - 1

    def partial(a, b):
        acc, pot = oracle.field_direct(src[a:b], soft[a:b], tgt, oracle.KERNEL_SPLINE, G, want_pot=True)
        return torch.from_numpy(np.concatenate([acc, pot[None]]))

    def fsub(field):
        field[:3] -= field[:3, row:row + 1].clone()

    field = field_build_sharded(partial, src.shape[0], fsub, row).numpy()

    # K4, target-sharded with an all-gather of unequal blocks
    n = 257
    pos = rng.normal(0, 1e-3, (3, n)) + np.array([[8.0], [0.0], [0.0]])
    mass = rng.uniform(0.5, 2.0, n)
    a, b = shard_range(n, rank, world)

    def acc_fn(pos_all, t0, t1):
        return torch.from_numpy(oracle.self_gravity(pos_all.numpy(), mass, 1e-10, G, t0=t0, t1=t1))

    acc_local = self_gravity_sharded(torch.from_numpy(np.ascontiguousarray(pos[:, a:b])),
                                     lambda p: allgather_particles(p, n), acc_fn, n).numpy()
    # equal blocks take the single-collective path (all_gather_into_tensor + one strided copy)
    n2 = 256
    vel = rng.normal(0, 1.0, (3, n2))
    a2, b2 = shard_range(n2, rank, world)
    gathered = allgather_particles(torch.from_numpy(np.ascontiguousarray(vel[:, a2:b2])), n2)
    assert gathered.shape == (3, n2) and np.array_equal(gathered.numpy(), vel)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), field=field, acc_local=acc_local, a=a, b=b)
    dist.barrier()
    dist.destroy_process_group()


def test_source_sharded_field_and_target_sharded_self_gravity(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from util import grid_targets, random_sources
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(99)
    src, soft = random_sources(rng, 3001, box=2.0)
    tgt = grid_targets(5)
    row = tgt.shape[0] - 1
    acc, pot = oracle.field_direct(src, soft, tgt, oracle.KERNEL_SPLINE, G, want_pot=True)
    want = np.concatenate([oracle.frame_subtract(acc, row), pot[None]])
    n = 257
    pos = rng.normal(0, 1e-3, (3, n)) + np.array([[8.0], [0.0], [0.0]])
    mass = rng.uniform(0.5, 2.0, n)
    want_acc = oracle.self_gravity(pos, mass, 1e-10, G)
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert np.array_equal(outs[0]["field"], outs[1]["field"])            # identical on every rank
    assert np.all(outs[0]["field"][:3, row] == 0.0)
    scale = np.abs(want).max(axis=1, keepdims=True)
    assert np.max(np.abs(outs[0]["field"] - want) / scale) < 1e-13       # sum of shard partials == whole (fp64)
    got = np.concatenate([o["acc_local"] for o in outs], axis=1)
    assert np.array_equal(got, want_acc)

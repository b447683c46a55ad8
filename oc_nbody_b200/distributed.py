"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the exchange.

Field build (K1) — source-sharded, BASELINE.json configs[4] / SURVEY §8(e):
    rank p holds sources [p*Ns/P, (p+1)*Ns/P), every rank holds all grid targets, each computes the full-grid
    partial field in FP64, then ONE all-reduce(sum, fp64, 3*(Ngrid+1)) and the frame subtraction (which must
    follow the reduction: gizmo_interface.py:569-571 subtracts the TOTAL field at the centre).
    The exchange is 6.3 MB (64^3) .. 50 MB (128^3) against seconds of compute per rank, so it is a plain NCCL
    call on the compute stream; there is nothing to fuse it with tile by tile.

Cluster self-gravity (K4) — target-sharded: every rank owns a contiguous block of stars, all-gathers positions
    each kick, computes accelerations for its block.

The functions take the per-rank compute step as a callable so the partition/reduction logic can be exercised
on CPU with the gloo backend (tests/test_cpu_distributed.py) without any GPU.
"""
import numpy as np


def shard_bounds(n, world):
    """Contiguous, balanced split of range(n): bounds[p] .. bounds[p+1] for rank p (sizes differ by <= 1)."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if p < extra else 0) for p in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_range(n, rank, world):
    b = shard_bounds(n, world)
    return int(b[rank]), int(b[rank + 1])


def _world(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allreduce_sum_(t, group=None):
    """In-place sum over ranks of a torch tensor (cuda -> NCCL, cpu -> gloo). No-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def field_build_sharded(partial_fn, n_src_total, frame_subtract_fn, center_row, group=None):
    """Source-sharded field build.

    partial_fn(src_begin, src_end) -> torch tensor [3 or 4, n_tgt] fp64: this rank's partial field (+ potential
        row) from its source shard — on the GPU path ``Context.field_direct`` on resident shard buffers.
    frame_subtract_fn(field) -> None: K1b on the reduced field (in place), or None to skip.
    Returns the reduced (and frame-subtracted) field, identical on every rank."""
    rank, world = _world(group)
    a, b = shard_range(n_src_total, rank, world)
    field = partial_fn(a, b)
    allreduce_sum_(field, group)
    if frame_subtract_fn is not None and center_row is not None and center_row >= 0:
        frame_subtract_fn(field)
    return field


def self_gravity_sharded(pos_local, gather_fn, acc_fn, n_total, group=None):
    """Target-sharded self-gravity for one kick.

    pos_local: this rank's block of positions, torch [3, n_local] fp64 (block = shard_range(n_total, rank, world)).
    gather_fn(pos_local) -> pos_all [3, n_total]: all-gather along the particle axis.
    acc_fn(pos_all, tgt_begin, tgt_end) -> acc [3, n_total] with [tgt_begin, tgt_end) filled (K4 with a target range).
    Returns acc_local [3, n_local]."""
    rank, world = _world(group)
    a, b = shard_range(n_total, rank, world)
    pos_all = gather_fn(pos_local)
    acc = acc_fn(pos_all, a, b)
    return acc[:, a:b]


def connect_comm(ctx, group=None, window_bytes=64 << 20):
    """Set up the peer-memory exchange of include/ocg.h (ocg_comm_*) for `ctx` over the ranks of a torch.distributed
    group: every rank allocates its window, the opaque handles are all-gathered with torch.distributed (any transport
    would do: they are 128 plain bytes), every rank maps its peers.  Returns (rank, world).  Idempotent per ctx."""
    import torch.distributed as dist
    rank, world = _world(group)
    if getattr(ctx, "comm_connected", False):
        return rank, world
    handle = ctx.comm_create(rank, world, window_bytes)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
    else:
        handles = [handle]
    ctx.comm_connect(handles)
    if world > 1:
        dist.barrier(group=group)  # every rank has mapped every window before the first exchange kernel runs
    return rank, world


def allgather_particles(pos_local, n_total, group=None):
    """All-gather [3, n_local] blocks (possibly of unequal size) into [3, n_total] on every rank.
    Equal blocks (n_total divisible by the world size): ONE NCCL all-gather into a [world, 3, n_local] buffer and one
    strided copy — two launches, capturable in a CUDA graph.  Unequal blocks are padded to the widest."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return pos_local
    if n_total % world == 0 and pos_local.is_contiguous():
        width = n_total // world
        buf = torch.empty((world * 3, width), dtype=pos_local.dtype, device=pos_local.device)  # rank blocks concatenated
        dist.all_gather_into_tensor(buf, pos_local, group=group)
        return buf.view(world, 3, width).permute(1, 0, 2).reshape(3, n_total)   # one copy kernel: [3][world][width] is [3][n_total]
    bounds = shard_bounds(n_total, world)
    width = int(np.max(np.diff(bounds)))
    send = torch.zeros((width, 3), dtype=pos_local.dtype, device=pos_local.device)
    send[: pos_local.shape[1]] = pos_local.t()
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out = torch.empty((3, n_total), dtype=pos_local.dtype, device=pos_local.device)
    for p in range(world):
        a, b = int(bounds[p]), int(bounds[p + 1])
        out[:, a:b] = recv[p][: b - a].t()
    return out

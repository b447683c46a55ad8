"""Peer-memory exchange (include/ocg.h ocg_comm_*, csrc/comm.cu) on the GPU.  One rank exercises every kernel path (a rank
signals and waits for itself); the 2-rank cases run tools/comm_check.py under torch.distributed.run and need 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from util import TOL, dev, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 4.3986004e-09


@pytest.fixture()
def comm_ctx():
    from oc_nbody_b200._lib import Context
    c = Context(0)
    h = c.comm_create(0, 1, 2 << 20)   # first half: force exchanges (96 B per star of a block), second half: all-reduce
    assert len(h) == Context.COMM_HANDLE_BYTES
    c.comm_connect([h])
    yield c
    c.close()


def test_allreduce_single_rank_is_identity_and_chunks(comm_ctx):
    import torch
    rng = np.random.default_rng(5)
    for n in (1, 1000, 65536, 200001):   # the last exceeds the 2 MiB window's single-pass capacity (65536 doubles): 4 passes
        x = rng.normal(size=n)
        d = dev(x)
        comm_ctx.comm_allreduce_f64(d)
        comm_ctx.comm_allreduce_f64(d)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), x)
    comm_ctx.comm_status()


def test_comm_errors(comm_ctx):
    import torch
    from oc_nbody_b200._lib import Context, OcgError
    plain = Context(0)
    try:
        with pytest.raises(OcgError, match="no communicator"):
            plain.comm_allreduce_f64(torch.zeros(4, dtype=torch.float64, device="cuda"))
        with pytest.raises(OcgError, match="rank"):
            plain.comm_create(3, 2, 1 << 20)
    finally:
        plain.close()
    with pytest.raises(OcgError, match="already"):
        comm_ctx.comm_create(0, 1, 1 << 20)
    with pytest.raises(OcgError, match="window"):   # 1 MiB window: 2 buffers x 3 x n doubles do not fit for 40 000 stars
        n = 40000
        comm_ctx.self_gravity_sharded(torch.zeros((3, n), dtype=torch.float64, device="cuda"),
                                      torch.ones(n, dtype=torch.float64, device="cuda"), 1e-10, G,
                                      torch.zeros((3, n), dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("n", [5000, 20000])
def test_self_gravity_sharded_single_rank_equals_k4(comm_ctx, n):
    """nranks = 1: the fused publish / gather / pack kernel + stream-K force kernel against the plain K4 call (same tiles,
    same kernel shape at these sizes: bit-identical) and the oracle; twice, so both window buffers are used."""
    import torch
    from oc_nbody_b200.synthetic import make_plummer_cluster
    p, _, mass = make_plummer_cluster(n, seed=n)
    pos = p * 1e-3 + np.array([[8.0], [0.0], [0.1]])
    eps2 = (0.01e-3) ** 2
    d_pos, d_m = dev(pos), dev(mass)
    ref_acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
    ref_pot = torch.empty(n, dtype=torch.float64, device="cuda")
    comm_ctx.debug_set("small_cluster_path", 0)
    comm_ctx.debug_set("direct_variant", 27)
    comm_ctx.self_gravity(d_pos, d_m, eps2, G, ref_acc, ref_pot)
    comm_ctx.debug_set("direct_variant", -1)
    for _ in range(2):
        acc = torch.full((3, n), np.nan, dtype=torch.float64, device="cuda")
        pot = torch.full((n,), np.nan, dtype=torch.float64, device="cuda")
        comm_ctx.self_gravity_sharded(d_pos, d_m, eps2, G, acc, pot)
        torch.cuda.synchronize()
        assert torch.equal(acc, ref_acc) and torch.equal(pot, ref_pot)
    comm_ctx.comm_status()
    o_acc, o_pot = oracle.self_gravity(pos, mass, eps2, G, want_pot=True)
    assert rel_err(acc.cpu().numpy(), o_acc, abs_sum=oracle.self_gravity_abs(pos, mass, eps2, G)) <= TOL
    assert np.max(np.abs(pot.cpu().numpy() - o_pot) / np.abs(o_pot)) <= TOL


def _torchrun(script_args, nproc=2, timeout=300):
    env = dict(os.environ)
    env.pop("OCG_TUNING_LIB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29577"] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_two_gpu_peer_exchange():
    """2 ranks: ocg_comm_allreduce_f64 against NCCL's all-reduce (and bit-identical on both ranks), the source-sharded
    field build against the FP64 oracle over BOTH shards, ocg_self_gravity_sharded against the single-GPU K4."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun([os.path.join(ROOT, "tools", "comm_check.py")])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"], line

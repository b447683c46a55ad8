import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, oracle
from util import *
from oc_nbody_b200 import Context
from test_gpu_direct import run_k1, G
ctx = Context(0)
for seed, n, box, cap in [(77, 4300, 2.0, 64), (5, 4300, 2.0, 64), (5, 2100, 2.0, 64), (77, 4300, 2.0, 512), (77, 4300, 5.0, 64), (5, 4300, 5.0, 64)]:
    rng = np.random.default_rng(seed)
    src, soft = random_sources(rng, n, box=box)
    tgt = grid_targets(41)
    ref, pref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    ctx.debug_set("near_cap", cap)
    out = []
    for pb in (0, 3 * 6 * 512 * 4, 1):
        ctx.debug_set("pass_bytes", pb)
        a, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER)
        out.append(rel_err(a, ref, abs_sum=cond))
    print(seed, n, box, cap, out)

"""GPU parity: K1 field build and K4 self-gravity through the C ABI vs the FP64 oracle."""
import numpy as np
import pytest

import oracle
from util import TOL, dev, grid_targets, random_sources, rel_err, rel_err_scalar

pytestmark = pytest.mark.gpu
G = 4.3986004135e-09
# The synthetic K1 cases below put the lattice targets INSIDE a Gaussian cloud of weakly softened sources (random_sources):
# single pair terms exceed the net field there, so they are gated on "strict bound OR backward-error bound of the sum"
# (util.rel_err with abs_sum = the oracle's sum of |pair terms|).  The galaxy-field cases (test_gpu_fullsize.py:
# configs[0], configs[1]) are gated on the strict bound alone.


def run_k1(ctx, src, soft, tgt, kernel, want_pot=False, accumulate_parts=1):
    import torch
    n_tgt = tgt.shape[0]
    acc = torch.full((3, n_tgt), 7.0, dtype=torch.float64, device="cuda")
    pot = torch.full((n_tgt,), 7.0, dtype=torch.float64, device="cuda") if want_pot else None
    d_tgt = dev(tgt)
    bounds = np.linspace(0, src.shape[0], accumulate_parts + 1).astype(int)
    for k in range(accumulate_parts):
        s, e = bounds[k], bounds[k + 1]
        ctx.field_direct(dev(src[s:e]), None if soft is None else dev(soft[s:e]), d_tgt, kernel, G, acc, pot,
                         accumulate=(k > 0))
    torch.cuda.synchronize()
    return acc.cpu().numpy(), (pot.cpu().numpy() if want_pot else None)


@pytest.mark.parametrize("kernel", [oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE])
@pytest.mark.parametrize("n_src,n_grid", [(20000, 12), (1537, 5), (513, 3)])
def test_field_direct_parity(ctx, kernel, n_src, n_grid):
    rng = np.random.default_rng(1776 + n_src)
    src, soft = random_sources(rng, n_src, box=2.0)
    tgt = grid_targets(n_grid)
    acc, pot = run_k1(ctx, src, soft, tgt, kernel, want_pot=True)
    ref, pref = oracle.field_direct(src, soft, tgt, kernel, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, kernel, G)
    assert rel_err(acc, ref, abs_sum=cond) <= TOL
    assert rel_err_scalar(pot, pref) <= TOL


@pytest.mark.parametrize("n_grid,n_src", [(27, 3000), (42, 1500)])
def test_field_direct_many_targets(ctx, n_grid, n_src):
    """Target counts that select the MID (>= 16k) and BIG (>= 64k) production kernels."""
    rng = np.random.default_rng(n_grid)
    src, soft = random_sources(rng, n_src, box=2.0)
    tgt = grid_targets(n_grid)
    for kernel in (oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE):
        acc, pot = run_k1(ctx, src, soft, tgt, kernel, want_pot=True)
        ref, pref = oracle.field_direct(src, soft, tgt, kernel, G, want_pot=True)
        cond = oracle.field_direct_abs(src, soft, tgt, kernel, G)
        assert rel_err(acc, ref, abs_sum=cond) <= TOL
        assert rel_err_scalar(pot, pref) <= TOL


def test_field_direct_sources_inside_support(ctx):
    """Spline sources sitting inside the grid box with supports that cover many targets (the near set)."""
    rng = np.random.default_rng(7)
    src, soft = random_sources(rng, 4000, box=0.4, soft_lo=0.05, soft_hi=0.8)
    tgt = grid_targets(9)
    for kernel in (oracle.KERNEL_SPLINE, oracle.KERNEL_PLUMMER):
        acc, pot = run_k1(ctx, src, soft, tgt, kernel, want_pot=True)
        ref, pref = oracle.field_direct(src, soft, tgt, kernel, G, want_pot=True)
        cond = oracle.field_direct_abs(src, soft, tgt, kernel, G)
        assert rel_err(acc, ref, abs_sum=cond) <= TOL
        assert rel_err_scalar(pot, pref) <= TOL


def test_field_direct_no_potential_and_accumulate(ctx):
    rng = np.random.default_rng(11)
    src, soft = random_sources(rng, 9000, box=3.0)
    tgt = grid_targets(7)
    ref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    acc1, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER)
    acc3, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, accumulate_parts=3)
    assert rel_err(acc1, ref, abs_sum=cond) <= TOL
    assert rel_err(acc3, ref, abs_sum=cond) <= TOL


def test_field_direct_deterministic(ctx):
    rng = np.random.default_rng(12)
    src, soft = random_sources(rng, 30000, box=3.0)
    tgt = grid_targets(10)
    a, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_SPLINE)
    b, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_SPLINE)
    assert np.array_equal(a, b)


def test_field_direct_edge_cases(ctx):
    import torch
    rng = np.random.default_rng(13)
    tgt = grid_targets(4)
    # no sources: overwrite with zeros
    acc = torch.full((3, tgt.shape[0]), 3.0, dtype=torch.float64, device="cuda")
    ctx.field_direct(torch.empty((0, 4), dtype=torch.float32, device="cuda"), None, dev(tgt), 0, G, acc)
    assert float(acc.abs().max()) == 0.0
    # one source, one target, no softening array (Newtonian two-body known answer)
    src = np.array([[0.3, -0.4, 1.2, 5.0e4]], np.float32)
    one = np.array([[0.0, 0.0, 0.0, 0.0]], np.float32)
    acc, pot = run_k1(ctx, src, None, one, oracle.KERNEL_PLUMMER, want_pot=True)
    r = np.sqrt(np.float64(src[0, 0]) ** 2 + np.float64(src[0, 1]) ** 2 + np.float64(src[0, 2]) ** 2)
    exact = G * np.float64(src[0, 3]) * src[0, :3].astype(np.float64) / r ** 3
    assert np.max(np.abs(acc[:, 0] - exact) / np.abs(exact)) <= TOL
    assert abs(pot[0] + G * np.float64(src[0, 3]) / r) / (G * src[0, 3] / r) <= TOL
    # coincident target and unsoftened source contributes nothing (r == 0 rule) in both kernels
    src = np.concatenate([random_sources(rng, 700, box=1.0)[0], np.array([[0.0, 0.0, 0.0, 1e9]], np.float32)])
    soft = np.concatenate([np.full(700, 0.05, np.float32), np.zeros(1, np.float32)])
    for kernel in (oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE):
        acc, pot = run_k1(ctx, src, soft, tgt, kernel, want_pot=True)
        ref, pref = oracle.field_direct(src, soft, tgt, kernel, G, want_pot=True)
        cond = oracle.field_direct_abs(src, soft, tgt, kernel, G)
        assert np.all(np.isfinite(acc)) and np.all(np.isfinite(pot))
        assert rel_err(acc, ref, abs_sum=cond) <= TOL
        assert rel_err_scalar(pot, pref) <= TOL


# Shape ids of the streaming kernel (direct_sum.cu table; stable across builds).  The shipped library carries only the
# production shapes; the sweep shapes and the timing experiments live in the separate OCG_TUNING build (tools/probe.py).
PRODUCTION_PLAIN = [1, 4, 26, 27, 31]  # SMALL, MID_GUARD, WIDE, MID, BIG (plain tiles: K4, and K1 with mass folding off)
PRODUCTION_MF_POT = [81, 80]           # BIG_MF (= BIG_MF_POT), MID_MF (mass-folded tiles, with or without potential)
PRODUCTION_MASS_FOLDED = PRODUCTION_MF_POT
TIMING_EXPERIMENTS = list(range(37, 40)) + list(range(48, 58))  # wrong results by construction


@pytest.mark.parametrize("variant", PRODUCTION_PLAIN + PRODUCTION_MASS_FOLDED)
def test_field_direct_variants(ctx, variant):
    """Every kernel shape of the shipped library, forced regardless of the target count."""
    rng = np.random.default_rng(21)
    src, soft = random_sources(rng, 11000, box=2.0)
    tgt = grid_targets(11)
    want_pot = variant in PRODUCTION_PLAIN + PRODUCTION_MF_POT
    ref, pref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    assert ctx.variant_built(variant)
    ctx.debug_set("direct_variant", variant)
    try:
        acc, pot = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=want_pot)
    finally:
        ctx.debug_set("direct_variant", -1)
    assert rel_err(acc, ref, abs_sum=cond) <= TOL
    if want_pot:
        assert rel_err_scalar(pot, pref) <= TOL


def test_shipped_library_has_no_timing_experiments(ctx):
    """The wrong-result timing kernels and the sweep shapes are not in the product library, and the knob refuses them."""
    from oc_nbody_b200._lib import OcgError
    n = ctx.variant_count()
    built = [v for v in range(n) if ctx.variant_built(v)]
    assert sorted(built) == sorted(PRODUCTION_PLAIN + PRODUCTION_MASS_FOLDED)
    for v in TIMING_EXPERIMENTS:
        with pytest.raises(OcgError):
            ctx.debug_set("direct_variant", v)


def test_knobs_are_per_context():
    """Tuning knobs are state of one ocg_ctx, not of the process (include/ocg.h: one ctx per GPU)."""
    from oc_nbody_b200._lib import Context
    rng = np.random.default_rng(5)
    src, soft = random_sources(rng, 2100, box=2.0)
    tgt = grid_targets(41)
    a, b = Context(0), Context(0)
    try:
        for c in (a, b):
            c.debug_set("precise_near", 0)  # 2100 sources would all fit the FP64 precision-radius set: keep them in the FP32 tiles
        a.debug_set("mass_fold", 0)
        ra, _ = run_k1(a, src, soft, tgt, oracle.KERNEL_PLUMMER)
        rb, _ = run_k1(b, src, soft, tgt, oracle.KERNEL_PLUMMER)
        assert not np.array_equal(ra, rb)      # b still takes the mass-folded kernel
        b.debug_set("mass_fold", 0)
        rb, _ = run_k1(b, src, soft, tgt, oracle.KERNEL_PLUMMER)
        assert np.array_equal(ra, rb)
    finally:
        a.close(), b.close()


@pytest.mark.parametrize("kernel", [oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE])
def test_field_direct_mass_folded_edge_masses(ctx, kernel):
    """Mass-folded tiles (w = (m/M0)^-1/2 folded into the coordinates): masses spanning 12 decades, zero and
    negative masses (FP64 near set), sources far enough that w^2 r^2 would overflow (also near set), sources
    inside the target box."""
    rng = np.random.default_rng(77)
    src, soft = random_sources(rng, 9000, box=1.5)
    src[:, 3] = np.exp(rng.uniform(np.log(1e-4), np.log(1e8), src.shape[0])).astype(np.float32)
    src[::97, 3] = 0.0
    src[5::211, 3] *= -1.0
    src[7::301, :3] *= 3.0e4          # tiny masses at huge distances: folded r^6 leaves the FP32 range
    src[7::301, 3] = 1e-3
    tgt = grid_targets(7)
    ref, pref = oracle.field_direct(src, soft, tgt, kernel, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, kernel, G)
    for variant in PRODUCTION_MASS_FOLDED:
        ctx.debug_set("direct_variant", variant)
        try:
            acc, pot = run_k1(ctx, src, soft, tgt, kernel, want_pot=variant in PRODUCTION_MF_POT)
        finally:
            ctx.debug_set("direct_variant", -1)
        assert np.all(np.isfinite(acc))
        assert rel_err(acc, ref, abs_sum=cond) <= TOL
        if pot is not None:
            assert np.all(np.isfinite(pot)) and rel_err_scalar(pot, pref) <= TOL


@pytest.mark.parametrize("variant", PRODUCTION_MASS_FOLDED)
@pytest.mark.parametrize("n_src", [511, 513, 1000, 1025, 1537])
def test_field_direct_swizzled_tiles_ragged_tail(ctx, variant, n_src):
    """The w array of a mass-folded tile is stored with adjacent sources swapped; odd and even source counts must
    leave the last real source's w intact when the tail of the tile is padded."""
    rng = np.random.default_rng(n_src)
    src, soft = random_sources(rng, n_src, box=2.0)
    src[:, :3] += np.float32(3.0)  # every source outside the precision radius: all of them in the FP32 tiles
    tgt = grid_targets(5)
    ref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    ctx.debug_set("direct_variant", variant)
    try:
        acc, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER)
    finally:
        ctx.debug_set("direct_variant", -1)
    assert rel_err(acc, ref, abs_sum=cond) <= TOL


def test_field_direct_production_big_is_mass_folded(ctx):
    """>= 64k targets takes the mass-folded kernels (12 targets/thread without potential, 8 with); with mass folding
    switched off the plain-tile kernel; all agree with the oracle to the parity tolerance."""
    rng = np.random.default_rng(5)
    src, soft = random_sources(rng, 2100, box=2.0)
    tgt = grid_targets(41)
    ref, pref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    ctx.debug_set("near_cap", 64)  # 2100 sources would all fit the FP64 precision-radius set: keep most in the FP32 tiles
    a_mf, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=False)
    a_mp, pot = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=True)
    assert rel_err(a_mf, ref, abs_sum=cond) <= TOL and rel_err(a_mp, ref, abs_sum=cond) <= TOL
    assert rel_err_scalar(pot, pref) <= TOL
    ctx.debug_set("mass_fold", 0)
    try:
        a_off, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=False)
        a_pl, pot_pl = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=True)
    finally:
        ctx.debug_set("mass_fold", 1)
        ctx.debug_set("near_cap", 0)
    assert np.array_equal(a_off, a_pl)     # mass folding off: the plain-tile kernel, potential or not
    assert not np.array_equal(a_mf, a_pl)  # two different kernels really ran
    assert rel_err(a_pl, ref, abs_sum=cond) <= TOL and rel_err_scalar(pot_pl, pref) <= TOL


@pytest.mark.parametrize("n_grid,want_pot", [(41, False), (41, True), (12, False), (4, True)])
def test_field_direct_stream_k_passes(ctx, n_grid, want_pot):
    """The source tiles are streamed in L2-sized passes (streamk.cuh, PASSES): with the pass shrunk to 3 tiles a 9-tile
    problem takes 3-4 passes (the last one shorter), every row's slots span the passes, and the result is the single-pass
    one to FP64 rounding, deterministic, and the oracle's to the parity tolerance."""
    rng = np.random.default_rng(77)
    src, soft = random_sources(rng, 4300, box=2.0)
    tgt = grid_targets(n_grid)
    ref, pref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=True)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    ctx.debug_set("near_cap", 512)  # keep most of the 4300 sources in the FP32 tiles (9 tiles)
    try:
        ctx.debug_set("pass_bytes", 0)
        a1, p1 = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=want_pot)
        ctx.debug_set("pass_bytes", 3 * 6 * 512 * 4)
        a3, p3 = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=want_pot)
        b3, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=want_pot)
        ctx.debug_set("pass_bytes", 1)  # one tile per pass
        a9, p9 = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER, want_pot=want_pot)
    finally:
        ctx.debug_set("pass_bytes", 32 << 20)
        ctx.debug_set("near_cap", 0)
    assert np.array_equal(a3, b3)
    scale = np.abs(a1).max()
    assert np.max(np.abs(a3 - a1)) <= 1e-13 * scale and np.max(np.abs(a9 - a1)) <= 1e-13 * scale
    assert rel_err(a3, ref, abs_sum=cond) <= TOL and rel_err(a9, ref, abs_sum=cond) <= TOL
    if want_pot:
        assert rel_err_scalar(p3, pref) <= TOL and rel_err_scalar(p9, pref) <= TOL


def test_field_direct_source_shards_share_the_near_set_limit(ctx):
    """ocg_set_source_shards(P): each of P strided source shards takes 1/P of the whole build's FP64 near-set limit; the
    partial fields add up to the unsharded result within the parity tolerance of the oracle, and the shards' near sets
    together are no larger than the unsharded one (the FP64 pass must not grow with the number of ranks)."""
    rng = np.random.default_rng(91)
    src, soft = random_sources(rng, 60000, box=3.0)
    tgt = grid_targets(9)
    ref = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=False)
    cond = oracle.field_direct_abs(src, soft, tgt, oracle.KERNEL_PLUMMER, G)
    ctx.debug_set("near_cap", 0)
    full, _ = run_k1(ctx, src, soft, tgt, oracle.KERNEL_PLUMMER)
    P = 4
    total = np.zeros_like(full)
    try:
        ctx.set_source_shards(P)
        for r in range(P):
            part, _ = run_k1(ctx, np.ascontiguousarray(src[r::P]), np.ascontiguousarray(soft[r::P]), tgt, oracle.KERNEL_PLUMMER)
            total += part
    finally:
        ctx.set_source_shards(1)
    assert rel_err(full, ref, abs_sum=cond) <= TOL and rel_err(total, ref, abs_sum=cond) <= TOL
    with pytest.raises(Exception):
        ctx.set_source_shards(0)


def test_frame_subtract_and_host_form(ctx):
    import torch
    rng = np.random.default_rng(31)
    center = np.array([8.0, 0.1, -0.2])
    n_src = 6000
    pos = rng.normal(0, 4.0, (n_src, 3)) + center * 0.5
    mass = np.exp(rng.uniform(np.log(1e3), np.log(1e5), n_src))
    soft = np.exp(rng.uniform(np.log(0.004), np.log(0.3), n_src))
    tgt4 = grid_targets(6)
    tgt = tgt4[:, :3].astype(np.float64) + center
    row = tgt.shape[0] - 1
    acc = ctx.field_build_host(pos, mass, soft, tgt, center, row, oracle.KERNEL_SPLINE, G)
    assert np.all(acc[:, row] == 0.0)
    s32 = oracle.recentre(pos, mass, center)
    t32 = oracle.recentre(tgt, None, center)
    raw = oracle.field_direct(s32, soft.astype(np.float32), t32, oracle.KERNEL_SPLINE, G)
    cond = oracle.field_direct_abs(s32, soft.astype(np.float32), t32, oracle.KERNEL_SPLINE, G)
    ref = oracle.frame_subtract(raw, row)
    # gate on the raw field: add the frame term back
    assert rel_err(acc + raw[:, row:row + 1], raw, abs_sum=cond) <= TOL
    # report-only residual metric
    print("tidal residual rel err:", rel_err(acc, ref))
    # device K1b alone
    d = dev(raw.copy())
    ctx.frame_subtract(d, row)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), ref)
    # the host form streams the sources through HBM in chunks (2^26 by default): force 4 ragged chunks
    ctx.debug_set("host_chunk", 1777)
    try:
        acc_c, pot_c = ctx.field_build_host(pos, mass, soft, tgt, center, row, oracle.KERNEL_SPLINE, G, want_pot=True)
    finally:
        ctx.debug_set("host_chunk", 0)
    assert np.all(acc_c[:, row] == 0.0)
    assert rel_err(acc_c + raw[:, row:row + 1], raw, abs_sum=cond) <= TOL
    _, pref = oracle.field_direct(s32, soft.astype(np.float32), t32, oracle.KERNEL_SPLINE, G, want_pot=True)
    assert rel_err_scalar(pot_c, pref) <= TOL


@pytest.mark.parametrize("n,eps_pc", [(1024, 0.01), (777, 0.0), (4100, 0.05), (20000, 0.01), (17000, 0.0), (66000, 0.01)])
def test_self_gravity_parity(ctx, n, eps_pc):
    import torch
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, _, mass = make_plummer_cluster(n, seed=n)
    pos = pos_pc * 1e-3 + np.array([[8.0], [0.01], [-0.02]])  # kpc, far from the origin
    eps2 = (eps_pc * 1e-3) ** 2
    acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
    pot = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.self_gravity(dev(pos), dev(mass), eps2, G, acc, pot)
    torch.cuda.synchronize()
    ref, pref = oracle.self_gravity(pos, mass, eps2, G, want_pot=True)
    # cluster self-gravity: close pairs make single terms exceed the net field => strict bound OR the backward-error
    # bound of the sum (util.rel_err, abs_sum)
    assert rel_err(acc.cpu().numpy(), ref, abs_sum=oracle.self_gravity_abs(pos, mass, eps2, G)) <= TOL
    assert rel_err_scalar(pot.cpu().numpy(), pref) <= TOL


def test_self_gravity_segments_and_shards(ctx):
    """Ragged batch of clusters; target range sharded as two ranks would do it."""
    import torch
    from oc_nbody_b200.synthetic import make_plummer_cluster
    sizes = [300, 1, 513, 1024, 0, 77]
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(seg[-1])
    pos = np.empty((3, n))
    mass = np.empty(n)
    for k, sz in enumerate(sizes):
        if sz == 0:
            continue
        p, _, m = make_plummer_cluster(max(sz, 2), seed=100 + k)
        ang = 2 * np.pi * k / len(sizes)
        pos[:, seg[k]:seg[k + 1]] = p[:, :sz] * 1e-3 + 8.0 * np.array([[np.cos(ang)], [np.sin(ang)], [0.0]])
        mass[seg[k]:seg[k + 1]] = m[:sz]
    eps2 = (0.01e-3) ** 2
    ref, pref = oracle.self_gravity(pos, mass, eps2, G, seg_offsets=seg, want_pot=True)
    acc = torch.zeros((3, n), dtype=torch.float64, device="cuda")
    pot = torch.zeros(n, dtype=torch.float64, device="cuda")
    half = n // 2
    for a, b in ((0, half), (half, n)):
        ctx.self_gravity(dev(pos), dev(mass), eps2, G, acc, pot, seg_offsets=seg, tgt_begin=a, tgt_end=b)
    torch.cuda.synchronize()
    assert rel_err(acc.cpu().numpy(), ref, abs_sum=oracle.self_gravity_abs(pos, mass, eps2, G, seg_offsets=seg)) <= TOL
    assert rel_err_scalar(pot.cpu().numpy()[pref != 0], pref[pref != 0]) <= TOL
    # lone particle: exactly zero
    assert np.all(acc.cpu().numpy()[:, seg[1]] == 0.0)


def test_errors_are_reported_not_fatal(ctx):
    import torch
    from oc_nbody_b200 import OcgError
    acc = torch.empty((3, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(OcgError, match="unknown softening kernel"):
        ctx.field_direct(torch.zeros((2, 4), dtype=torch.float32, device="cuda"), None,
                         torch.zeros((4, 4), dtype=torch.float32, device="cuda"), 9, G, acc)
    with pytest.raises(OcgError, match="centre row"):
        ctx.frame_subtract(acc, 17)
    with pytest.raises(OcgError, match="eps2"):
        ctx.self_gravity(torch.zeros((3, 4), dtype=torch.float64, device="cuda"),
                         torch.ones(4, dtype=torch.float64, device="cuda"), -1.0, G, acc)


@pytest.mark.parametrize("n,world", [(65536, 8), (16384, 8), (20000, 3)])
def test_self_gravity_large_cluster_target_shards(ctx, n, world):
    """One large cluster with its targets sharded as `world` ranks would (the star-sharded BRIDGE step): every shard
    takes the few-targets kernel variant against ALL source tiles; the union must equal the unsharded call."""
    import torch
    from oc_nbody_b200.distributed import shard_bounds
    from oc_nbody_b200.synthetic import make_plummer_cluster
    p, _, mass = make_plummer_cluster(n, seed=4)
    pos = p * 1e-3 + np.array([[8.0], [0.0], [0.0]])
    eps2 = (0.01e-3) ** 2
    d_pos, d_m = dev(pos), dev(mass)
    full = torch.empty((3, n), dtype=torch.float64, device="cuda")
    ctx.self_gravity(d_pos, d_m, eps2, G, full)
    acc = torch.full((3, n), np.nan, dtype=torch.float64, device="cuda")
    b = shard_bounds(n, world)
    for r in range(world):
        ctx.self_gravity(d_pos, d_m, eps2, G, acc, tgt_begin=int(b[r]), tgt_end=int(b[r + 1]))
    torch.cuda.synchronize()
    ref, cond = oracle.self_gravity(pos, mass, eps2, G), oracle.self_gravity_abs(pos, mass, eps2, G)
    assert rel_err(full.cpu().numpy(), ref, abs_sum=cond) <= TOL
    assert rel_err(acc.cpu().numpy(), ref, abs_sum=cond) <= TOL
    if n == 65536:
        # enough source tiles per shard for the target-paired kernels: the sharded result is bit-identical to the
        # unsharded one (same per-target accumulation order), so an N-GPU BRIDGE run reproduces the 1-GPU trajectory
        assert np.array_equal(acc.cpu().numpy(), full.cpu().numpy())


@pytest.mark.parametrize("n,eps_pc", [(1024, 0.01), (4096, 0.01), (333, 0.0), (2, 0.01), (1, 0.01)])
def test_self_gravity_small_cluster_path(ctx, n, eps_pc):
    """The fused one-launch kernel for a single small cluster (N <= 4096; the reference runs N = 1024, test_options:57)
    against the oracle and against the streaming kernel it stands in for, full range and a target shard."""
    import torch
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, _, mass = make_plummer_cluster(max(n, 2), seed=n)
    pos = (pos_pc * 1e-3 + np.array([[8.0], [0.01], [-0.02]]))[:, :n]
    mass = mass[:n]
    eps2 = (eps_pc * 1e-3) ** 2
    d_pos, d_m = dev(np.ascontiguousarray(pos)), dev(np.ascontiguousarray(mass))
    res = {}
    for small in (1, 0):
        ctx.debug_set("small_cluster_path", small)
        try:
            acc = torch.zeros((3, n), dtype=torch.float64, device="cuda")
            pot = torch.zeros(n, dtype=torch.float64, device="cuda")
            ctx.self_gravity(d_pos, d_m, eps2, G, acc, pot)
            sh = torch.full((3, n), np.nan, dtype=torch.float64, device="cuda")
            ctx.self_gravity(d_pos, d_m, eps2, G, sh, None, tgt_begin=n // 3, tgt_end=n - n // 4)
        finally:
            ctx.debug_set("small_cluster_path", 1)
        torch.cuda.synchronize()
        res[small] = (acc.cpu().numpy(), pot.cpu().numpy(), sh.cpu().numpy())
    ref, pref = oracle.self_gravity(pos, mass, eps2, G, want_pot=True)
    cond = oracle.self_gravity_abs(pos, mass, eps2, G)
    for small in (1, 0):
        a, p, sh = res[small]
        if n > 1:
            assert rel_err(a, ref, abs_sum=cond) <= TOL
            assert rel_err_scalar(p, pref) <= TOL
        else:
            assert np.all(a == 0.0) and np.all(p == 0.0)
        lo, hi = n // 3, n - n // 4
        assert np.array_equal(sh[:, lo:hi], a[:, lo:hi]) and np.all(np.isnan(sh[:, :lo])) and np.all(np.isnan(sh[:, hi:]))

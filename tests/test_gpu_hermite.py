"""K6 on the GPU: the Hermite force loop (acceleration + jerk) and the 4th-order Hermite predictor / corrector —
ph4's own arithmetic (oc_code.py:218-229) — against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import oracle
from util import EPS_PAIR, TOL, dev, rel_err, rel_err_norm, rel_err_scalar

pytestmark = pytest.mark.gpu

G = 4.3986004e-09      # kpc^2 km/s /Myr /Msun (gizmo_interface.py:70)
VTL = 1.022712165045695e-3  # kpc per Myr at 1 km/s
EPS2 = (0.01e-3) ** 2  # (0.01 pc)^2 in kpc^2 (test_options:63, oc_code.py:225)


def cluster(n, seed=1777, center=(8.0, 0.0, 0.0), vsys=(0.0, 220.0, 0.0), kroupa=False):
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, vel, mass = make_plummer_cluster(n, seed=seed)
    if kroupa:
        mass = np.exp(np.random.default_rng(seed).uniform(np.log(0.1), np.log(20.0), n))
    return pos_pc * 1e-3 + np.asarray(center)[:, None], vel + np.asarray(vsys)[:, None], mass


def gpu_force(ctx, pos, vel, mass, eps2=EPS2, seg=None, t0=0, t1=None, want_pot=True):
    import torch
    n = pos.shape[1]
    acc = torch.full((3, n), np.nan, dtype=torch.float64, device="cuda")
    jerk = torch.full((3, n), np.nan, dtype=torch.float64, device="cuda")
    pot = torch.full((n,), np.nan, dtype=torch.float64, device="cuda") if want_pot else None
    ctx.self_gravity_hermite(dev(pos), dev(vel), dev(mass), eps2, G, VTL, acc, jerk, pot, seg_offsets=seg, tgt_begin=t0,
                             tgt_end=t1)
    torch.cuda.synchronize()
    return acc.cpu().numpy(), jerk.cpu().numpy(), (pot.cpu().numpy() if want_pot else None)


# the jerk term chains ~twice as many FP32 operations as the acceleration term (d.w, alpha, alpha*d + w)
EPS_PAIR_JERK = 2 * EPS_PAIR


def hermite_ref(pos, vel, mass, eps2=EPS2, seg=None, t0=0, t1=None):
    """FP64 oracle (acc, jerk, pot) plus the condition numbers of the two sums (sum of |pair terms|): cluster
    self-gravity is where close pairs make single terms exceed the net field, see util.rel_err."""
    a, j, p = oracle.self_gravity_hermite(pos, vel, mass, eps2, G, VTL, seg_offsets=seg, t0=t0, t1=t1, want_pot=True)
    sa, sj = oracle.self_gravity_abs(pos, mass, eps2, G, vel=vel, vel_to_len=VTL, seg_offsets=seg, t0=t0, t1=t1)
    return a, j, p, sa, sj


def check(got, ref, sl=slice(None)):
    a, j, p = got
    ra, rj, rp, sa, sj = ref
    assert rel_err(a[:, sl], ra[:, sl], abs_sum=sa[:, sl]) <= TOL
    assert rel_err(j[:, sl], rj[:, sl], abs_sum=sj[:, sl], eps_pair=EPS_PAIR_JERK) <= TOL
    if p is not None:
        assert rel_err_scalar(p[sl], rp[sl]) <= TOL


@pytest.mark.parametrize("n", [2, 33, 1024, 4096])
def test_small_cluster_path_matches_oracle(ctx, n):
    pos, vel, mass = cluster(n, kroupa=(n == 1024))
    ref = hermite_ref(pos, vel, mass)
    check(gpu_force(ctx, pos, vel, mass), ref)


@pytest.mark.parametrize("n", [700, 5000, 12289])
def test_streaming_path_matches_oracle_and_small_path(ctx, n):
    pos, vel, mass = cluster(n, seed=3)
    ref = hermite_ref(pos, vel, mass)
    ctx.debug_set("hermite_small_path", 0)
    try:
        got = gpu_force(ctx, pos, vel, mass)
        again = gpu_force(ctx, pos, vel, mass)
    finally:
        ctx.debug_set("hermite_small_path", 1)
    check(got, ref)
    for x, y in zip(got, again):
        assert np.array_equal(x, y)  # run-to-run deterministic
    if n <= 4096:
        small = gpu_force(ctx, pos, vel, mass)
        assert rel_err(got[0], small[0], abs_sum=ref[3]) <= TOL and rel_err(got[1], small[1], abs_sum=ref[4]) <= TOL


def test_every_kernel_variant_matches_oracle(ctx):
    pos, vel, mass = cluster(5000, seed=8)
    ref = hermite_ref(pos, vel, mass)
    nv = ctx.variant_count(family=1)
    assert nv >= 3
    try:
        for v in range(nv):  # a shape this build does not carry falls back to the production one
            ctx.debug_set("hermite_variant", v)
            check(gpu_force(ctx, pos, vel, mass), ref)
            check(gpu_force(ctx, pos, vel, mass, want_pot=False), ref)
    finally:
        ctx.debug_set("hermite_variant", -1)


def test_batch_of_ragged_clusters_and_target_shards(ctx):
    lens = [4096, 0, 1, 777, 2048, 5000, 513]
    seg = np.concatenate([[0], np.cumsum(lens)])
    parts = [cluster(l, seed=20 + i, center=(8.0 * np.cos(i), 8.0 * np.sin(i), 0.1 * i)) for i, l in enumerate(lens) if l]
    pos, vel, mass = (np.concatenate([p[k] for p in parts], axis=-1) for k in range(3))
    ref = hermite_ref(pos, vel, mass, seg=seg)
    got = gpu_force(ctx, pos, vel, mass, seg=seg)
    check(got, ref)
    # the lone star feels nothing
    i1 = seg[2]
    assert np.all(got[0][:, i1] == 0) and np.all(got[1][:, i1] == 0) and got[2][i1] == 0
    # a rank's target shard: written for [t0, t1) only, same values
    n = pos.shape[1]
    t0, t1 = 3000, 9001
    sh = gpu_force(ctx, pos, vel, mass, seg=seg, t0=t0, t1=t1)
    for x, y in zip(sh, got):
        assert np.array_equal(x[..., t0:t1], y[..., t0:t1])
        assert np.all(np.isnan(x[..., :t0])) and np.all(np.isnan(x[..., t1:]))
    assert n > t1


def test_unsoftened_form_skips_coincident_pairs(ctx):
    pos, vel, mass = cluster(3000, seed=4)
    pos[:, 17] = pos[:, 5]       # a coincident pair: contributes nothing when eps2 == 0
    pos[:, 2999] = pos[:, 0]     # ... including one with the recentring origin (where tile padding sits)
    ref = hermite_ref(pos, vel, mass, eps2=0.0)
    for small in (1, 0):
        ctx.debug_set("hermite_small_path", small)
        try:
            got = gpu_force(ctx, pos, vel, mass, eps2=0.0)
        finally:
            ctx.debug_set("hermite_small_path", 1)
        assert all(np.all(np.isfinite(x)) for x in got)
        check(got, ref)


def test_argument_errors(ctx):
    import torch
    from oc_nbody_b200._lib import OcgError
    pos, vel, mass = cluster(64)
    a = torch.empty((3, 64), dtype=torch.float64, device="cuda")
    with pytest.raises(OcgError):
        ctx.self_gravity_hermite(dev(pos), dev(vel), dev(mass), -1.0, G, VTL, a, a.clone())
    with pytest.raises(OcgError):
        ctx.self_gravity_hermite(dev(pos), dev(vel), dev(mass), EPS2, G, VTL, a, a.clone(), tgt_begin=10, tgt_end=65)
    with pytest.raises(OcgError):
        ctx.self_gravity_hermite(dev(pos), dev(vel), dev(mass), EPS2, G, VTL, a, a.clone(), seg_offsets=[0, 40, 30, 64])
    with pytest.raises(OcgError):
        ctx.hermite_correct(dev(pos), dev(vel), a, a, dev(pos), dev(vel), a, a, 0.0, VTL, 0.14)


def test_predictor_and_corrector_are_bit_exact(ctx):
    import torch
    rng = np.random.default_rng(12)
    n = 3001
    pos, vel, mass = cluster(n, seed=6)
    a0, j0 = oracle.self_gravity_hermite(pos, vel, mass, EPS2, G, VTL)
    dt = 0.0123
    xp_ref, vp_ref = oracle.hermite_predict(pos, vel, a0, j0, dt, VTL)
    xp, vp = torch.empty((3, n), dtype=torch.float64, device="cuda"), torch.empty((3, n), dtype=torch.float64, device="cuda")
    ctx.hermite_predict(dev(pos), dev(vel), dev(a0), dev(j0), dt, VTL, xp, vp)
    assert np.array_equal(xp.cpu().numpy(), xp_ref) and np.array_equal(vp.cpu().numpy(), vp_ref)
    a1, j1 = oracle.self_gravity_hermite(xp_ref, vp_ref, mass, EPS2, G, VTL)
    x_ref, v_ref, dtm_ref = oracle.hermite_correct(xp_ref, vp_ref, a0, j0, a1, j1, dt, VTL, 0.14)
    x, v, da0, dj0 = dev(pos), dev(vel), dev(a0), dev(j0)
    dtm = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.hermite_correct(x, v, da0, dj0, xp, vp, dev(a1), dev(j1), dt, VTL, 0.14, dtm)
    assert np.array_equal(x.cpu().numpy(), x_ref) and np.array_equal(v.cpu().numpy(), v_ref)
    assert np.array_equal(da0.cpu().numpy(), a1) and np.array_equal(dj0.cpu().numpy(), j1)
    assert abs(dtm.item() - dtm_ref) <= 1e-12 * dtm_ref
    del rng


@pytest.mark.parametrize("n,substeps", [(1024, 4), (6000, 2)])
def test_cluster_code_hermite_evolve_matches_oracle(ctx, n, substeps):
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.units import units
    pos, vel, mass = cluster(n, seed=31)
    cl = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=substeps, ctx=ctx, integrator="hermite")
    span = 0.05
    cl.evolve_model(span | units.Myr)
    x_ref, v_ref, dtm_ref = oracle.hermite_evolve(pos, vel, mass, EPS2, G, span, substeps, VTL, 0.14)
    x, v = cl.pos.cpu().numpy(), cl.vel.cpu().numpy()
    # displacements over the span agree to the force tolerance (norm metric: a trajectory is not a sum of pair terms,
    # and a velocity change component can pass through zero)
    assert rel_err_norm(x - pos, x_ref - pos) <= TOL
    assert rel_err_norm(v - vel, v_ref - vel) <= TOL
    assert abs(cl.dt_min.item() - dtm_ref) <= 1e-3 * dtm_ref
    assert cl.suggested_substeps(span) >= 1
    # Hermite and leapfrog integrate the same dynamics
    lf = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=64 * substeps, ctx=ctx)
    lf.evolve_model(span | units.Myr)
    scale = np.abs(x_ref - pos).max()
    assert np.max(np.abs(lf.pos.cpu().numpy() - x)) <= 2e-2 * scale


def test_bridge_with_hermite_cluster_eager_equals_cuda_graph(ctx):
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    from oc_nbody_b200.units import units
    center = np.array([8.0, 0.0, 0.0])
    snaps = [make_snapshot(20000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / 8)
    field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    pos, vel, mass = cluster(1024, seed=2)
    out = []
    for graph in (False, True):
        field.evolve_grid(center)
        field.evolve_model(0.0 | units.Myr)
        cl = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=2, ctx=ctx, integrator="hermite")
        system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=graph)
        system.add_system(cl, (field,))
        system.add_system(field)
        system.evolve_model(0.5 | units.Myr, timestep=0.1 | units.Myr)
        out.append((cl.pos.cpu().numpy(), cl.vel.cpu().numpy()))
        if graph:
            assert system.graph_replays >= 3
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-13 * 8.0
    assert np.max(np.abs(out[0][1] - out[1][1])) <= 1e-12 * 220.0


def test_block_time_steps_equal_level_limit_is_the_shared_step_scheme(ctx):
    """ocg_hermite_block_evolve with every star forced onto the smallest step (eta -> 0) takes exactly 2^k block steps of n
    stars and reproduces cluster_code's shared-step Hermite run with substeps = 2^k (same force kernel, same predictor /
    corrector arithmetic; the force over 'active targets x all sources' walks the tiles in another order, so equal to FP64
    rounding of the sums)."""
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.units import units
    n, k, span = 5000, 3, 0.02
    pos, vel, mass = cluster(n, seed=12)
    shared = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=2 ** k, ctx=ctx, integrator="hermite")
    shared.evolve_model(span | units.Myr)
    block = cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx, integrator="hermite", block_steps=True, eta=1e-12, max_level=k)
    block.evolve_model(span | units.Myr)
    assert block.block_step_count == 2 ** k and block.star_step_count == n * 2 ** k
    xs, vs, xb, vb = (t.cpu().numpy() for t in (shared.pos, shared.vel, block.pos, block.vel))
    assert np.max(np.abs(xb - xs)) <= 1e-15 * 8.0 * 4 and np.max(np.abs(vb - vs)) <= 1e-13 * np.max(np.abs(vs))


def test_block_time_steps_match_the_oracle_and_save_work(ctx):
    """A 400-star cluster with a hard binary in it: the GPU block-step run against the oracle's (FP64 forces; same schedule
    arithmetic).  The binary sits on steps hundreds of times shorter than the bulk, so the star-steps taken are a small
    fraction of what one shared step of the smallest size would cost; energy is conserved to the scheme's accuracy."""
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR, cluster_code
    from oc_nbody_b200.units import units
    n, span, eta, max_level = 400, 0.05, 0.02, 14
    pos, vel, mass = cluster(n, seed=21)
    # hard binary: stars 0 and 1 at 40 AU-ish (2e-4 pc) on a circular orbit about their barycentre
    sep = 2e-7  # kpc
    pos[:, 1] = pos[:, 0] + np.array([sep, 0.0, 0.0])
    vc = np.sqrt(G * (mass[0] + mass[1]) / sep / KMS_TO_KPC_PER_MYR)  # km/s
    vel[:, 1] = vel[:, 0] + np.array([0.0, vc, 0.0])
    eps2 = (1e-5 * 1e-3) ** 2   # 1e-5 pc softening: the binary is resolved
    cl = cluster_code(mass, pos, vel, softening_pc=1e-5, ctx=ctx, integrator="hermite", block_steps=True, eta=eta, max_level=max_level)

    def energy(x, v):
        _, _, pot = oracle.self_gravity_hermite(x, v, mass, eps2, G, VTL, want_pot=True)
        vv = v - (v * mass).sum(axis=1, keepdims=True) / mass.sum()
        return float((0.5 * mass * (vv * vv).sum(axis=0)).sum() + 0.5 * (mass * pot).sum() / KMS_TO_KPC_PER_MYR)
    e0 = energy(pos, vel)
    cl.evolve_model(span | units.Myr)
    x, v = cl.pos.cpu().numpy(), cl.vel.cpu().numpy()
    xo, vo, _, _, steps_o, star_o = oracle.hermite_block_evolve(pos, vel, mass, eps2, G, span, VTL, eta, max_level)
    # the schedules agree to a few per cent (FP32 pair arithmetic can move a borderline step choice), far below a shared step's cost
    assert abs(cl.block_step_count - steps_o) <= 0.05 * steps_o and abs(cl.star_step_count - star_o) <= 0.05 * star_o
    smallest = span / 2 ** max_level
    assert cl.block_step_count > 50 and cl.star_step_count < 0.05 * n * (span / smallest)
    # single stars follow the oracle closely; the binary's phase is the sensitive quantity (hundreds of orbits): compare its
    # barycentre and separation instead
    # (1e-4 of the largest displacement: FP32 pair arithmetic against the oracle's FP64 sums, through ~100 block steps whose
    # borderline step choices may differ)
    rest = np.arange(2, n)
    assert np.max(np.abs(x[:, rest] - xo[:, rest])) <= 1e-4 * np.max(np.abs(xo[:, rest] - pos[:, rest]))
    bc = lambda q: (q[:, 0] * mass[0] + q[:, 1] * mass[1]) / (mass[0] + mass[1])  # noqa: E731
    assert np.max(np.abs(bc(x) - bc(xo))) <= 1e-4 * np.max(np.abs(bc(xo) - bc(pos)))
    assert abs(np.linalg.norm(x[:, 1] - x[:, 0]) - sep) <= 2e-2 * sep
    # energy: the binary (eta = 0.02, hundreds of orbits) carries the error budget; the oracle's own run lands at the same level
    de, de_o = abs(energy(x, v) - e0) / abs(e0), abs(energy(xo, vo) - e0) / abs(e0)
    print("block steps: energy error GPU %.2e, oracle %.2e" % (de, de_o))
    assert de <= 1e-3 and de_o <= 1e-3   # FP32 pair arithmetic adds ~3e-4 through the hard binary's ~1e4 force evaluations
    with pytest.raises(ValueError):
        cluster_code(mass, pos, vel, ctx=ctx, block_steps=True)


def test_fullsize_65536_third_law_and_rows(ctx):
    """configs[2] size: N = 65 536.  Newton's third law for the acceleration AND the jerk (sum m a = sum m j = 0 up to
    rounding), and 150 rows against the oracle over all sources."""
    n = 65536
    pos, vel, mass = cluster(n, seed=1777, kroupa=True)
    a, j, p = gpu_force(ctx, pos, vel, mass)
    for f in (a, j):
        net = (mass * f).sum(axis=1)
        assert np.max(np.abs(net)) <= 1e-6 * (mass * np.sqrt((f * f).sum(axis=0))).sum()
    rows = np.unique(np.random.default_rng(1).integers(0, n, 150))
    ref = [hermite_ref(pos, vel, mass, t0=int(t), t1=int(t) + 1) for t in rows]
    ea = max(rel_err(a[:, t:t + 1], r[0][:, t:t + 1], abs_sum=r[3][:, t:t + 1]) for t, r in zip(rows, ref))
    ej = max(rel_err(j[:, t:t + 1], r[1][:, t:t + 1], abs_sum=r[4][:, t:t + 1], eps_pair=EPS_PAIR_JERK) for t, r in zip(rows, ref))
    ep = max(rel_err_scalar(p[t:t + 1], r[2][t:t + 1]) for t, r in zip(rows, ref))
    assert ea <= TOL and ej <= TOL and ep <= TOL


@pytest.mark.parametrize("exchange,graph", [("peer", False), ("peer", True), ("nccl", False)])
def test_two_gpu_sharded_hermite_matches_single_gpu(exchange, graph):
    """K6 target-sharded over 2 GPUs: positions and velocities gathered per force evaluation — over peer memory inside the
    tile pack (ocg_self_gravity_hermite_sharded, also as a captured CUDA graph) or NCCL-all-gathered — predictor /
    corrector local, Aarseth step all-reduced (tools/bridge_multi.py --integrator hermite): bit-identical to one GPU."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(root, "tools", "bridge_multi.py"), "--stars", "8192", "--steps", "2",
           "--integrator", "hermite", "--exchange", exchange] + (["--graph"] if graph else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["match"] and line["n_gpus"] == 2 and line["integrator"] == "hermite" and line["exchange"] == exchange
    assert line["max_rel_dx"] == 0.0 and line["max_rel_dv"] <= 1e-15


def test_driver_loop_with_hermite_graph_steps_and_k4_bookkeeping_between_them(ctx):
    """The reference's driver loop (oc_nbody.py:49-66) above the small-cluster size: every step replays the captured
    Hermite step (K6, streaming kernel with its own work plan and scratch) and then runs K4 with the potential for the
    bound-subset centre of mass.  The two plans must not disturb each other: graph and eager runs agree."""
    from oc_nbody_b200.driver import evolve_cluster_in_galaxy
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    center = np.array([8.0, 0.0, 0.0])
    snaps = [make_snapshot(20000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / 8)
    pos, vel, mass = cluster(5000, seed=12)
    out = []
    for graph in (False, True):
        field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
        field.evolve_grid(center)
        cl, rec = evolve_cluster_in_galaxy(field, mass, pos, vel, timestep=0.1, tend=0.65, softening_pc=0.01, substeps=1,
                                           use_cuda_graph=graph, ctx=ctx, integrator="hermite")
        out.append((cl.pos.cpu().numpy(), cl.vel.cpu().numpy(), len(rec.frames)))
    assert out[0][2] == out[1][2] == 7
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-12 * 8.0
    assert np.max(np.abs(out[0][1] - out[1][1])) <= 1e-11 * 220.0


def test_hermite_auto_substeps_follow_the_aarseth_criterion(ctx):
    """substeps="auto": the count is a power of two that brings the shared step under the Aarseth minimum measured on the
    device; the result equals the fixed-count run with the same count and conserves energy far better than one step."""
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR, cluster_code
    from oc_nbody_b200.units import units
    import torch
    pos, vel, mass = cluster(1024, seed=44)
    span = 0.05

    def energy(cl):
        pot = torch.empty(cl.n, dtype=torch.float64, device="cuda")
        ctx.self_gravity_hermite(cl.pos, cl.vel, cl.mass, EPS2, G, VTL, cl.acc1, cl.jerk1, pot)
        v = cl.vel - (cl.vel * cl.mass).sum(dim=1, keepdim=True) / cl.mass.sum()
        return float((0.5 * cl.mass * (v * v).sum(dim=0)).sum() + 0.5 * (cl.mass * pot).sum() / KMS_TO_KPC_PER_MYR)
    auto = cluster_code(mass, pos, vel, softening_pc=0.01, substeps="auto", ctx=ctx, integrator="hermite")
    e0 = energy(auto)
    auto.evolve_model(span | units.Myr)
    n1 = None
    first = auto.substeps  # already the suggestion for the NEXT call
    assert first >= 1 and first & (first - 1) == 0
    # the count used for the first call came from eta * min |a|/|j|
    a, j = oracle.self_gravity_hermite(pos, vel, mass, EPS2, G, VTL)
    dt0 = 0.14 * np.sqrt(((a * a).sum(axis=0) / (j * j).sum(axis=0)).min())
    n1 = int(2 ** max(0, int(np.ceil(np.log2(span / dt0)))))
    fixed = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=n1, ctx=ctx, integrator="hermite")
    fixed.evolve_model(span | units.Myr)
    assert np.array_equal(auto.pos.cpu().numpy(), fixed.pos.cpu().numpy())
    assert span / first <= auto.dt_min.item() * (1 + 1e-12) or first == auto.max_substeps
    one = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=1, ctx=ctx, integrator="hermite")
    one.evolve_model(span | units.Myr)
    de_auto, de_one = abs(energy(auto) - e0), abs(energy(one) - e0)
    assert de_auto <= 1e-5 * abs(e0) and (n1 == 1 or de_auto < de_one)
    with pytest.raises(ValueError):
        cluster_code(mass, pos, vel, substeps="auto", ctx=ctx)

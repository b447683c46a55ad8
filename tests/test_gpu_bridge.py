"""GPU integration: the plugin classes (field code, cluster code, Bridge) against the oracle pipeline.
Config 1 of BASELINE.json in miniature: 1k-star Plummer cluster BRIDGE-kicked by a 16^3 grid field."""
import os

import numpy as np
import pytest

import oracle
from util import TOL, rel_err, rel_err_norm

pytestmark = pytest.mark.gpu


def oracle_field(field, snap, center, want_pot=True):
    """The oracle's version of _populate_grid_acceleration_ for one snapshot."""
    r, m, soft = field._source_arrays_(snap)
    g = field.grid
    tgt = g.init_grid + center
    s32 = oracle.recentre(r, m, center)
    t32 = oracle.recentre(tgt, None, center)
    kern = oracle.KERNEL_SPLINE if field.softening_kernel == "spline" else oracle.KERNEL_PLUMMER
    raw, pot = oracle.field_direct(s32, soft.astype(np.float32), t32, kern, field.G, want_pot=True)
    # miniature snapshots (1e4..1e5 particles): single pair terms rival the net field at some grid points, so the raw field
    # is gated on "strict bound OR backward-error bound of the sum" (util.rel_err, abs_sum); the full-size configs are not
    oracle_field.cond = oracle.field_direct_abs(s32, soft.astype(np.float32), t32, kern, field.G)
    return raw, oracle.frame_subtract(raw, g.origin_row), pot


@pytest.fixture(scope="module")
def small_world(ctx):
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    snap0 = make_snapshot(40000, seed=1776)
    snap1 = advance_snapshot(snap0, 23.0)
    opts = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 8,
                softening_kernel="spline")
    centers = np.array([[8.0, 0.0, 0.0], [7.99, 0.19, 0.0]])
    field = gizmo_field(opts, [snap0, snap1], chosen_positions=centers, chosen_id=5, ctx=ctx)
    return field, (snap0, snap1), centers


def test_field_code_build_matches_oracle(small_world):
    field, snaps, centers = small_world
    g = field.grid
    assert g.snapshot_acceleration_x.shape == (2, 8 ** 3 + 1)
    for i in range(2):
        raw, sub, pot = oracle_field(field, snaps[i], centers[i])
        got = np.stack([g.snapshot_acceleration_x[i], g.snapshot_acceleration_y[i], g.snapshot_acceleration_z[i]])
        assert np.all(got[:, g.origin_row] == 0.0)
        assert rel_err(got + raw[:, g.origin_row:g.origin_row + 1], raw, abs_sum=oracle_field.cond) <= TOL
        assert np.max(np.abs(g.snapshot_potential[i] - pot) / np.abs(pot)) <= TOL


def test_get_gravity_at_point_forms(small_world):
    from oc_nbody_b200.units import units
    field, _, _ = small_world
    rng = np.random.default_rng(3)
    g = field.grid
    field.evolve_grid(np.array([8.0, 0.0, 0.0]))
    field.evolve_model(9.2 | units.Myr)
    a, b, w = field._bracket
    assert (a, b) == (0, 1) and abs(w - 0.4) < 1e-12
    p = rng.uniform(-0.05, 0.05, (257, 3)) + np.array([8.0, 0.0, 0.0])
    ax, ay, az = field.get_gravity_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    got = np.stack([c.value_in(units.kms / units.Myr) for c in (ax, ay, az)])
    # oracle on the arrays the field code itself exposes (fp64 stacks -> fp32 records)
    recs = [oracle.pack_planes(np.stack([g.snapshot_acceleration_x[i], g.snapshot_acceleration_y[i],
                                         g.snapshot_acceleration_z[i]]), g.snapshot_potential[i]) for i in range(2)]
    ref, refpot = oracle.grid_interp(g.nodes, np.array([[8.0, 0.0, 0.0]]), recs[0], recs[1], w, p[:, 0], p[:, 1], p[:, 2],
                                     want_pot=True)
    assert np.array_equal(got, ref)
    # scalar form, parsec input
    sx, sy, sz = field.get_gravity_at_point(0 | units.kpc, p[0, 0] * 1e3 | units.parsec, p[0, 1] * 1e3 | units.parsec,
                                            p[0, 2] * 1e3 | units.parsec)
    assert abs(sx.value_in(units.kms / units.Myr) - ref[0, 0]) <= 1e-9 * abs(ref[0, 0])
    phi = field.get_potential_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    assert np.allclose(phi.value_in(units.kms ** 2), refpot / 1.022712165045695e-3, rtol=1e-14)
    # materialised blend equals the oracle's
    assert np.array_equal(field.evolved_acceleration, oracle.time_blend(recs[0], recs[1], w))
    # bare floats (already kpc) are accepted too
    fx, _, _ = field.get_gravity_at_point(0.0, p[:, 0], p[:, 1], p[:, 2])
    assert np.array_equal(fx.value_in(units.kms / units.Myr), ref[0])


def test_bridge_steps_match_oracle(small_world, ctx):
    """K(dt/2) D(dt) K(dt/2) for a few steps: device-resident path and generic-partner path vs the CPU BRIDGE."""
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR, cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import units
    field, _, _ = small_world
    g = field.grid
    pos_pc, vel, mass = make_plummer_cluster(1024)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    field.evolve_grid(center)
    field.evolve_model(0.0 | units.Myr)
    dt, nstep = 0.1, 3
    eps2 = (0.01e-3) ** 2
    recs = [oracle.pack_planes(np.stack([g.snapshot_acceleration_x[i], g.snapshot_acceleration_y[i],
                                         g.snapshot_acceleration_z[i]]), g.snapshot_potential[i]) for i in range(2)]

    # CPU BRIDGE with the oracle pieces; field time advances inside the drift like amuse's bridge does
    x, v, t = pos.copy(), vel.copy(), 0.0
    for _ in range(nstep):
        def tidal(xx, tt):
            _, _, w = oracle.time_bracket(field.time_in_Myr, tt)
            return oracle.grid_interp(g.nodes, center[None], recs[0], recs[1], w, xx[0], xx[1], xx[2])
        v = oracle.kick(v, tidal(x, t), 0.5 * dt)
        a = oracle.self_gravity(x, mass, eps2, field.G)
        v = oracle.kick(v, a, 0.5 * dt)
        x = oracle.drift(x, v, dt, KMS_TO_KPC_PER_MYR)
        a = oracle.self_gravity(x, mass, eps2, field.G)
        v = oracle.kick(v, a, 0.5 * dt)
        t += dt
        v = oracle.kick(v, tidal(x, t), 0.5 * dt)

    for generic in (False, True):
        cl = cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx)
        field.evolve_model(0.0 | units.Myr)
        partner = field
        if generic:
            class Wrapped(object):  # hides kick_device: forces the public get_gravity_at_point path
                def get_gravity_at_point(self, *a):
                    return field.get_gravity_at_point(*a)

                def evolve_model(self, *a, **k):
                    return field.evolve_model(*a, **k)
            partner = Wrapped()
        system = Bridge(timestep=dt | units.Myr, use_threading=False)
        system.add_system(cl, (partner,))
        system.add_system(partner)
        for i in range(nstep + 1):
            system.evolve_model(i * dt | units.Myr, timestep=dt | units.Myr)
        p = system.particles
        gx = p.position.value_in(units.kpc).T - center[:, None]
        gv = p.velocity.value_in(units.kms).T
        # offsets from the cluster centre (pc scale) and velocities both within 1e-5
        # trajectories: norm metric (a position / velocity component is not a sum of pair terms and may pass through 0)
        assert rel_err_norm(gx, x - center[:, None]) <= TOL
        assert rel_err_norm(gv, v) <= TOL
    com = cl.bound_center_of_mass()
    assert np.linalg.norm(com - center) < 2e-3


# ------------------------------------------------------------------ the reference's default: nested fine grid ----
@pytest.mark.parametrize("time_interpolation", ["linear", "cubic"])
def test_field_code_with_fine_grid(ctx, time_interpolation):
    """options with grid_fine_* (test_options:93-100 in miniature): field build on the reference's point list
    (kept coarse | fine | origin), two-level kick, potential and tidal tensor, against the oracle pipeline."""
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    from oc_nbody_b200.units import units
    from oc_nbody_b200 import time_spline
    nsnap = 2 if time_interpolation == "linear" else 4
    snaps = [make_snapshot(15000, seed=1776)]
    for _ in range(nsnap - 1):
        snaps.append(advance_snapshot(snaps[-1], 23.0))
    center = np.array([8.0, 0.0, 0.0])
    opts = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 7,
                fine_grid=True, grid_fine_x_size_in_kpc=0.02, grid_fine_y_size_in_kpc=0.02, grid_fine_z_size_in_kpc=0.02,
                grid_fine_resolution=0.02 / 6, softening_kernel="spline", time_interpolation=time_interpolation)
    field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (nsnap, 1)), ctx=ctx)
    g = field.grid
    assert g.has_fine_grid and g.snapshot_acceleration_x.shape == (nsnap, len(g))
    assert len(g) == g.n_lattice - len(g.coarse_hole_index) + int(np.prod(g.fine_shape)) + 1 and len(g.coarse_hole_index) > 0
    # field build on the nested point list vs the oracle
    raw, sub, pot = oracle_field(field, snaps[0], center)
    got = np.stack([g.snapshot_acceleration_x[0], g.snapshot_acceleration_y[0], g.snapshot_acceleration_z[0]])
    assert np.all(got[:, g.origin_row] == 0.0)
    assert rel_err(got + raw[:, g.origin_row:g.origin_row + 1], raw, abs_sum=oracle_field.cond) <= TOL
    # the kick: oracle layout (hole filled from the fine lattice) + two-level interpolation, bit for bit
    planes = field._planes_()
    t = 9.2 if time_interpolation == "linear" else 30.0
    field.evolve_grid(center)
    field.evolve_model(t | units.Myr)
    if time_interpolation == "cubic":
        knots, coef = time_spline.fit(field.time_in_Myr, planes)
        first, w = time_spline.basis(knots, t)
        use, weights = coef[first:first + 4], list(w)
    else:
        wb = np.float32(0.4)
        use, weights = planes, [float(np.float32(np.float32(1.0) - wb)), float(wb)]
    coarse, fine = oracle.layout_nested(use, g.n_lattice, g.coarse_keep_index, g.coarse_hole_index, g.coarse_hole_points(),
                                        g.fine_nodes, g.fine_row0)
    rng = np.random.default_rng(8)
    p = rng.uniform(-0.06, 0.06, (600, 3))
    p[:300] = rng.uniform(-0.025, 0.025, (300, 3))
    x = p + center
    ref = oracle.grid_interp_nested(g.nodes, g.fine_nodes, center, list(coarse), list(fine), weights, x[:, 0], x[:, 1], x[:, 2],
                                    want_pot=True, want_tensor=True, want_level=True)
    assert 100 < ref["level"].sum() < 500
    ax, ay, az = field.get_gravity_at_point(0 | units.kpc, x[:, 0] | units.kpc, x[:, 1] | units.kpc, x[:, 2] | units.kpc)
    got = np.stack([c.value_in(units.kms / units.Myr) for c in (ax, ay, az)])
    assert np.array_equal(got, ref["acc"])
    phi = field.get_potential_at_point(0 | units.kpc, x[:, 0] | units.kpc, x[:, 1] | units.kpc, x[:, 2] | units.kpc)
    assert np.allclose(phi.value_in(units.kms ** 2), ref["pot"] / 1.022712165045695e-3, rtol=1e-14)
    T = field.get_tidal_tensor_at_point(0 | units.kpc, x[:, 0] | units.kpc, x[:, 1] | units.kpc, x[:, 2] | units.kpc)
    Tv = T.value_in(units.kms / units.Myr / units.kpc)
    assert Tv.shape == (600, 3, 3)
    assert np.array_equal(Tv, ref["tensor"].T.reshape(-1, 3, 3))
    T0 = field.get_tidal_tensor_at_point(0 | units.kpc, x[0, 0] | units.kpc, x[0, 1] | units.kpc, x[0, 2] | units.kpc)
    assert np.array_equal(T0.value_in(units.kms / units.Myr / units.kpc), Tv[0])
    # physics: the tidal tensor of a smooth field is nearly symmetric and nearly trace-free away from sources
    assert field.evolved_acceleration.shape == (3, len(g))


# ------------------------------------------------------------------ per-step cluster bookkeeping on the device ----
@pytest.mark.parametrize("n", [1000, 1025, 4096])
def test_bound_com_eject_and_compaction_match_oracle(ctx, n):
    """ocg_bound_com / ocg_eject_mask / ocg_compact_rows vs the numpy restatement of oc_nbody.py:60-61 and
    oc_code.py:231-246: masks and medians bit for bit, centre of mass to FP64 summation-order accuracy."""
    import torch
    from util import dev
    rng = np.random.default_rng(n)
    pos = rng.normal(0, 1e-3, (3, n)) + np.array([[8.0], [0.1], [-0.2]])
    vel = rng.normal(0, 0.5, (3, n)) + np.array([[10.0], [200.0], [5.0]])
    vel[:, ::7] += rng.normal(0, 30.0, (3, len(range(0, n, 7))))      # unbound escapers
    pos[:, 3::50] += rng.normal(0, 0.2, (3, len(range(3, n, 50))))     # far-flung stars for the ejection cut
    mass = rng.uniform(0.3, 3.0, n)
    pot_v2 = -rng.uniform(0.5, 20.0, n)
    d_pos, d_vel, d_mass, d_pot = dev(pos), dev(vel), dev(mass), dev(pot_v2)
    out = torch.empty((1, 8), dtype=torch.float64, device="cuda")
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.bound_com(d_pos, d_vel, d_mass, d_pot, 1.0, out, None, mask)
    torch.cuda.synchronize()
    com, bound, vcom = oracle.bound_com(pos, vel, mass, pot_v2)
    res = out.cpu().numpy()[0]
    assert 0 < bound.sum() < n
    assert np.array_equal(mask.cpu().numpy().astype(bool), bound)
    assert res[4] == bound.sum() and np.isclose(res[3], mass[bound].sum(), rtol=1e-13)
    assert np.allclose(res[:3], com, rtol=1e-13, atol=0) and np.allclose(res[5:], vcom, rtol=1e-13, atol=0)
    # nothing bound -> every star counts
    ctx.bound_com(d_pos, d_vel, d_mass, dev(np.full(n, 1e9)), 1.0, out, None, mask)
    torch.cuda.synchronize()
    assert int(out[0, 4].item()) == n and bool(mask.all())
    assert np.allclose(out.cpu().numpy()[0, :3], (pos * mass).sum(axis=1) / mass.sum(), rtol=1e-13)
    # a batch of clusters (segments), one result row each
    seg = np.array([0, n // 3, n // 3 + 1, n], np.int64)
    out3 = torch.empty((3, 8), dtype=torch.float64, device="cuda")
    ctx.bound_com(d_pos, d_vel, d_mass, d_pot, 1.0, out3, dev(seg), None)
    torch.cuda.synchronize()
    for k in range(3):
        a, b = seg[k], seg[k + 1]
        c, bd, _ = oracle.bound_com(pos[:, a:b], vel[:, a:b], mass[a:b], pot_v2[a:b])
        assert np.allclose(out3.cpu().numpy()[k, :3], c, rtol=1e-13) and out3[k, 4].item() == bd.sum()

    # ejection cut: exact medians (odd and even n), keep mask, stable compaction
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    med = torch.empty(3, dtype=torch.float64, device="cuda")
    ctx.eject_mask(d_pos, 1000.0, 20.0, keep, med)
    torch.cuda.synchronize()
    want_keep, want_med = oracle.eject_keep(pos, 1000.0, 20.0)
    assert np.array_equal(med.cpu().numpy(), want_med)
    assert np.array_equal(keep.cpu().numpy().astype(bool), want_keep) and 0 < want_keep.sum() < n
    rows = np.concatenate([mass[None], pos, vel])
    nk = torch.zeros(1, dtype=torch.int64, device="cuda")
    outr = torch.full((7, int(want_keep.sum())), np.nan, dtype=torch.float64, device="cuda")
    ctx.compact_rows(dev(rows), keep, outr, nk)
    torch.cuda.synchronize()
    assert int(nk.item()) == want_keep.sum()
    assert np.array_equal(outr.cpu().numpy(), rows[:, want_keep])


def test_cluster_code_clean_ejections_and_bound_com(ctx):
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR, cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, vel, mass = make_plummer_cluster(2000)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    pos[:, 5] += 0.5          # 500 pc away: ejected
    vel[:, 11] += 300.0       # unbound but still inside the cut
    cl = cluster_code(mass, pos, vel, softening_pc=0.01, eject_cut=100.0, ctx=ctx)
    removed = cl.clean_ejections()
    want_keep, _ = oracle.eject_keep(pos, 1000.0, 100.0)
    assert list(removed) == list(np.where(~want_keep)[0]) and 5 in removed and cl.n == want_keep.sum()
    p = cl.particles
    assert np.array_equal(p.position.value_in(p.position.unit).T, pos[:, want_keep]) and np.array_equal(p.key, np.where(want_keep)[0])
    com, mask = cl.bound_center_of_mass(return_mask=True)
    x, v, m = pos[:, want_keep], vel[:, want_keep], mass[want_keep]
    _, phi = oracle.self_gravity(x, m, (0.01e-3) ** 2, cl.G, want_pot=True)
    want_com, want_mask, _ = oracle.bound_com(x, v, m, phi / KMS_TO_KPC_PER_MYR)
    assert np.array_equal(mask, want_mask) and not want_mask[list(np.where(want_keep)[0]).index(11)]
    assert np.allclose(com - center, want_com - center, rtol=1e-6, atol=1e-12)


def test_cluster_code_iterated_bound_subset(ctx):
    """bound_center_of_mass(iterations=8): with a one-sided tail of escapers the frame is re-taken from the bound stars until
    the set is stable; mask and centre of mass equal the oracle's iteration (that the iteration changes the answer on such a
    configuration is shown on the CPU, tests/test_cpu_host.py::test_iterated_bound_subset_logic_matches_the_oracle)."""
    from oc_nbody_b200.cluster import KMS_TO_KPC_PER_MYR, cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, vel, mass = make_plummer_cluster(3000, seed=3)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    tail = np.arange(0, 3000, 4)
    vel[0, tail] += 6.0
    pos[0, tail] += 0.004
    cl = cluster_code(mass, pos, vel, softening_pc=0.01, ctx=ctx)
    com1, mask1 = cl.bound_center_of_mass(return_mask=True)
    com8, mask8 = cl.bound_center_of_mass(return_mask=True, iterations=8)
    _, phi = oracle.self_gravity(pos, mass, (0.01e-3) ** 2, cl.G, want_pot=True)
    w1, wm1, _ = oracle.bound_com(pos, vel, mass, phi / KMS_TO_KPC_PER_MYR)
    w8, wm8, _ = oracle.bound_com(pos, vel, mass, phi / KMS_TO_KPC_PER_MYR, iterations=8)
    print("bound stars, one pass / iterated:", int(wm1.sum()), int(wm8.sum()))
    assert np.array_equal(mask1, wm1)
    assert np.array_equal(mask8, wm8)
    assert np.allclose(com1 - center, w1 - center, rtol=1e-6, atol=1e-12)
    assert np.allclose(com8 - center, w8 - center, rtol=1e-6, atol=1e-12)
    assert cl.n_bound == int(wm8.sum()) and np.isclose(cl.bound_mass, mass[wm8].sum())


def test_field_code_snapshot_caches_round_trip(ctx, tmp_path):
    """Second construction with the same options is served from the reference-format caches, bit for bit."""
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    snaps = [make_snapshot(5000, seed=3)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 5,
                cache_directory=str(tmp_path / "cache"))
    a = gizmo_field(opts, snaps, ctx=ctx)
    # 2 snapshots x (x, y, z, pot) + the whole-grid pickle (gizmo_interface.py:510), + one provenance sidecar for each of the 3
    names = sorted(p.name for p in (tmp_path / "cache").iterdir())
    assert a.cache_hits == 0 and not a.grid_cache_hit and len(names) == 12
    assert sum(n.endswith(".provenance.json") for n in names) == 3
    os.remove(a._grid_cache_name_()[1])   # without the whole-grid file the per-snapshot caches serve the build
    b = gizmo_field(opts, snaps, ctx=ctx)
    assert b.cache_hits == 2 and not b.grid_cache_hit
    for k in ("snapshot_acceleration_x", "snapshot_acceleration_y", "snapshot_acceleration_z", "snapshot_potential"):
        assert np.array_equal(getattr(a.grid, k), getattr(b.grid, k))
    c = gizmo_field(dict(opts, with_potential=False), snaps, ctx=ctx)   # b rewrote the whole-grid file: served from it
    assert c.grid_cache_hit and c.cache_hits == 0 and c.grid.snapshot_potential is None
    assert np.array_equal(c.grid.snapshot_acceleration_x, a.grid.snapshot_acceleration_x)
    assert a.cache_unverified == [] and b.cache_unverified == [] and c.cache_unverified == []
    # other softening settings under the same (reference-format) file names: the sidecars say so and the caches are rebuilt
    with pytest.warns(UserWarning, match="other settings"):
        d = gizmo_field(dict(opts, softening_kernel="plummer"), snaps, ctx=ctx)
    assert not d.grid_cache_hit and d.cache_hits == 0
    assert not np.array_equal(d.grid.snapshot_acceleration_x, a.grid.snapshot_acceleration_x)
    # a cache without sidecars (what the reference writes) is accepted as the reference accepts it, flagged and warned about
    for p in (tmp_path / "cache").iterdir():
        if p.name.endswith(".provenance.json"):
            p.unlink()
    with pytest.warns(UserWarning, match="no provenance sidecar"):
        e = gizmo_field(dict(opts, softening_kernel="plummer"), snaps, ctx=ctx)
    assert e.grid_cache_hit and len(e.cache_unverified) == 1
    assert np.array_equal(e.grid.snapshot_acceleration_x, d.grid.snapshot_acceleration_x)


@pytest.mark.parametrize("graph", [False, True])
def test_two_gpu_sharded_bridge_matches_single_gpu(graph):
    """K4 target-sharded over 2 GPUs with an NCCL all-gather of the positions per evaluation (tools/bridge_multi.py);
    graph=True replays the whole sharded step, collectives included, as one CUDA graph per rank."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29518" if graph else "29517", os.path.join(root, "tools", "bridge_multi.py"), "--stars", "8192",
           "--steps", "2"] + (["--graph"] if graph else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["match"] and line["n_gpus"] == 2
    assert not graph or line["graph_replays_last_run"] >= 2


def test_bridge_cuda_graph_step_is_bit_identical(small_world, ctx):
    """use_cuda_graph=True: eager first steps, then captured, replayed afterwards — the eager bridge's trajectory to
    FP64 rounding, including after the grid is re-centred (evolve_grid only rewrites a device buffer the graph reads)."""
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import units
    field, _, _ = small_world
    pos_pc, vel, mass = make_plummer_cluster(1500)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    dt, nstep = 0.1, 7
    out = []
    for use_graph in (False, True):
        field.evolve_grid(center)
        field.evolve_model(0.0 | units.Myr)
        cl = cluster_code(mass, pos, vel, softening_pc=0.01, substeps=2, ctx=ctx)
        system = Bridge(timestep=dt | units.Myr, use_threading=False, use_cuda_graph=use_graph)
        system.add_system(cl, (field,))
        system.add_system(field)
        for i in range(nstep + 1):
            system.evolve_model(i * dt | units.Myr, timestep=dt | units.Myr)
            if i == 4:
                field.evolve_grid(center + np.array([1e-4, -2e-4, 5e-5]))
        out.append((cl.pos.cpu().numpy(), cl.vel.cpu().numpy(), cl.model_time, field.time, system.graph_replays,
                    system.graph_captures))
    (x0, v0, t0, ft0, r0, c0), (x1, v1, t1, ft1, r1, c1) = out
    # step 1 is eager (no force yet), step 2 eager with the slot kernels (sizes the scratch), step 3 captures, 3.. replay
    assert r0 == 0 and c0 == 0 and c1 == 1 and r1 == nstep - 2, (r0, c0, r1, c1)
    # the eager drift integrates over span = (t + dt) - model_time, which differs from dt in the last bit at some
    # steps; the graph bakes dt itself into the captured kernels.  Everything else is the same arithmetic.
    assert np.max(np.abs(x0 - x1)) <= 1e-14 * 8.0 and np.max(np.abs(v0 - v1)) <= 1e-13 * np.max(np.abs(v0))
    assert abs(t0 - t1) < 1e-12 and ft0 == ft1 and abs(t1 - nstep * dt) < 1e-12


def _run_bridge_(field, cl, use_graph, nstep, dt, between=None):
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.units import units
    system = Bridge(timestep=dt | units.Myr, use_threading=False, use_cuda_graph=use_graph)
    system.add_system(cl, (field,))
    system.add_system(field)
    for i in range(1, nstep + 1):
        system.evolve_model(i * dt | units.Myr, timestep=dt | units.Myr)
        if between is not None:
            between(i, cl)
    return system


def test_bridge_cuda_graph_survives_scratch_growth_and_particle_removal(small_world, ctx):
    """More than 4096 stars: the streaming K4 path, whose tiles / targets / partials / work plan live in ctx scratch that a
    captured graph freezes.  Between replays the driver loop's own calls run on the same ctx — bound_center_of_mass()
    (K4 WITH the potential: 4 partial components instead of 3) and clean_ejections() (the star count changes) — and a
    larger foreign call grows the scratch buffers.  The graph must notice (ocg_capture_epoch) and re-capture; the
    trajectory must equal the eager bridge's doing the same calls."""
    import torch
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import units
    field, _, _ = small_world
    pos_pc, vel, mass = make_plummer_cluster(6000, seed=9)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    pos[:, 17] += np.array([0.03, 0.0, 0.0])   # a star 30 pc out: clean_ejections removes it (and the halo beyond 10 pc)
    dt, nstep = 0.05, 9
    big = make_plummer_cluster(30000, seed=10)
    big_pos, big_m = torch.from_numpy(big[0] * 1e-3 + center[:, None]).cuda(), torch.from_numpy(big[2]).cuda()
    big_acc = torch.empty((3, 30000), dtype=torch.float64, device="cuda")
    coms = {False: [], True: []}
    out = {}
    for use_graph in (False, True):
        field.evolve_grid(center)
        field.evolve_model(0.0 | units.Myr)
        cl = cluster_code(mass, pos, vel, softening_pc=0.01, eject_cut=10.0, ctx=ctx)

        def between(i, c, use_graph=use_graph):
            if i in (4, 7):
                coms[use_graph].append(c.bound_center_of_mass())
            if i == 5:
                c.clean_ejections()
            if i == 6:   # a larger self-gravity call of someone else on the SAME ctx: scratch buffers are reallocated
                c.ctx.self_gravity(big_pos, big_m, 1e-10, 4.3986004e-09, big_acc)
        sysm = _run_bridge_(field, cl, use_graph, nstep, dt, between)
        out[use_graph] = (cl.pos.cpu().numpy(), cl.vel.cpu().numpy(), cl.n, sysm.graph_captures, sysm.graph_replays)
    (x0, v0, n0, c0, r0), (x1, v1, n1, c1, r1) = out[False], out[True]
    assert n0 == n1 < 6000 and c0 == 0 and c1 >= 2 and r1 >= 3, (n0, n1, c1, r1)
    assert np.max(np.abs(x0 - x1)) <= 1e-14 * 8.0 and np.max(np.abs(v0 - v1)) <= 1e-13 * np.max(np.abs(v0))
    for a, b in zip(coms[False], coms[True]):
        assert np.max(np.abs(a - b)) <= 1e-14 * 8.0


def test_two_clusters_with_graphs_on_one_ctx(small_world, ctx):
    """Two cluster codes stepped alternately, both captured as graphs, both on ONE shared ctx: each one's K4 uploads its own
    work plan into the same device buffer.  A replay with the other cluster's plan resident would give wrong forces;
    the capture epoch makes the bridge fall back to an eager step instead.  Reference: each cluster alone, eager, on a
    ctx of its own."""
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import units
    field, _, _ = small_world
    center = np.array([8.0, 0.0, 0.0])
    dt, nstep = 0.05, 6
    sets = []
    for n, seed in ((5000, 3), (7001, 4)):
        p, v, m = make_plummer_cluster(n, seed=seed)
        sets.append((m, p * 1e-3 + center[:, None], v))
    ref = []
    for m, x, v in sets:
        field.evolve_grid(center)
        field.evolve_model(0.0 | units.Myr)
        cl = cluster_code(m, x, v, softening_pc=0.01)          # private ctx
        assert cl.ctx is not ctx
        _run_bridge_(field, cl, False, nstep, dt)
        ref.append((cl.pos.cpu().numpy(), cl.vel.cpu().numpy()))
    from oc_nbody_b200.bridge import Bridge
    field.evolve_grid(center)
    cls = [cluster_code(m, x, v, softening_pc=0.01, ctx=ctx) for m, x, v in sets]
    systems = []
    for cl in cls:
        s = Bridge(timestep=dt | units.Myr, use_threading=False, use_cuda_graph=True)
        s.add_system(cl, (field,))
        s.add_system(field)
        systems.append(s)
    for i in range(1, nstep + 1):
        for s in systems:
            field.evolve_model((i - 1) * dt | units.Myr)   # both bridges share the field code: rewind it for the second
            s.evolve_model(i * dt | units.Myr, timestep=dt | units.Myr)
    for cl, (x, v) in zip(cls, ref):
        assert np.max(np.abs(cl.pos.cpu().numpy() - x)) <= 1e-14 * 8.0
        assert np.max(np.abs(cl.vel.cpu().numpy() - v)) <= 1e-13 * np.max(np.abs(v))


@pytest.mark.parametrize("opts", [dict(), dict(star_char_mass=7100.0, dark_char_mass=35000.0, softening_kernel="plummer"),
                                  dict(clean_Rmax=True, Rmax=9.0)])
def test_device_source_assembly_matches_the_host_rules(ctx, opts):
    """ocg_assemble_sources (tracked star dropped, Rmax cut, per-species softening, star|dark|gas concatenation, FP64
    recentring + FP32 rounding) against the host restatement of gizmo_interface.py:297-304,515-558 — records equal bit for
    bit (the cube-root rule to 1 FP32 ulp: device pow vs numpy power) — and the field built from them equal to the
    host-assembled build."""
    from oc_nbody_b200.gizmo_field import clean_Rmag, gizmo_field
    from oc_nbody_b200.synthetic import make_snapshot
    snap = make_snapshot(30000, seed=5)
    center = np.array([8.0, 0.0, 0.0])
    base = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 6)
    base.update(opts)
    chosen = int(snap["star"]["id"][123])
    f_dev = gizmo_field(dict(base, source_assembly="device"), [snap], chosen_positions=center[None], chosen_id=chosen, ctx=ctx)
    host_snap = make_snapshot(30000, seed=5)
    if opts.get("clean_Rmax"):
        n_before = sum(len(host_snap[k]["mass"]) for k in ("star", "dark", "gas"))
        clean_Rmag(host_snap, 9.0)   # what the reference does at ingestion (gizmo_interface.py:254-255,297-304)
        assert sum(len(host_snap[k]["mass"]) for k in ("star", "dark", "gas")) < n_before
    f_host = gizmo_field(dict(base, source_assembly="host", clean_Rmax=False), [host_snap], chosen_positions=center[None],
                         chosen_id=chosen, ctx=ctx)
    r, m, soft = f_host._source_arrays_(host_snap)
    want = oracle.recentre(r, m, center)
    xyzm, dsoft = f_dev._assemble_device_(snap, center)
    got, gsoft = xyzm.cpu().numpy(), dsoft.cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    if "star_char_mass" in opts:
        assert np.max(np.abs(gsoft - soft.astype(np.float32)) / soft) <= 1.2e-7
    else:
        assert np.array_equal(gsoft, soft.astype(np.float32))
    gd, gh = f_dev.grid, f_host.grid
    for a, b in ((gd.snapshot_acceleration_x, gh.snapshot_acceleration_x), (gd.snapshot_acceleration_z, gh.snapshot_acceleration_z),
                 (gd.snapshot_potential, gh.snapshot_potential)):
        assert np.max(np.abs(a - b)) <= 1e-6 * np.max(np.abs(b)) if "star_char_mass" in opts else np.array_equal(a, b)


def test_whole_grid_cache_and_interface_dump_round_trip(ctx, tmp_path):
    """gizmo_interface.py:395-398,510 and oceanic_io.py:72-145: a second field code on the same cache directory loads the
    whole-grid pickle (no GPU build), and dump_interface / load_interface reproduce the kick bit for bit."""
    from oc_nbody_b200 import cache_compat
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    from oc_nbody_b200.units import units
    snaps = [make_snapshot(20000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    center = np.array([8.0, 0.0, 0.0])
    opts = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 6,
                cache_directory=str(tmp_path / "cache"))
    a = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    assert not a.grid_cache_hit and os.path.exists(a._grid_cache_name_()[1])
    launches = ctx.launch_count()
    b = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    assert b.grid_cache_hit and np.array_equal(b.grid.snapshot_acceleration_y, a.grid.snapshot_acceleration_y)
    assert np.array_equal(b.grid.snapshot_potential, a.grid.snapshot_potential)
    # only the plane upload ran on the GPU: no classify / direct-sum launches
    assert ctx.launch_count() - launches < 10
    a.output_directory = str(tmp_path)
    a.evolve_grid(center)
    out = cache_compat.dump_interface(a, "interface")
    assert sorted(os.listdir(out)) == sorted(["evolved_grid", "init_grid", "interface", "grid_accx_interpolators",
                                              "grid_accy_interpolators", "grid_accz_interpolators", "snapshot_acceleration_x.npz",
                                              "snapshot_acceleration_y.npz", "snapshot_acceleration_z.npz", "snapshot_potential.npz"])
    c = cache_compat.load_interface(out, ctx=ctx)
    x = center[:, None] + np.random.default_rng(2).normal(0, 0.01, (3, 200))
    for t in (0.0, 7.5):
        a.evolve_model(t | units.Myr), c.evolve_model(t | units.Myr)
        ga = a.get_gravity_at_point(0 | units.kpc, x[0] | units.kpc, x[1] | units.kpc, x[2] | units.kpc)
        gc = c.get_gravity_at_point(0 | units.kpc, x[0] | units.kpc, x[1] | units.kpc, x[2] | units.kpc)
        for u, v in zip(ga, gc):
            assert np.array_equal(u.value_in(units.kms / units.Myr), v.value_in(units.kms / units.Myr))


def test_pykdgrav_compat_call_sites(ctx):
    """The reference's own three calls (gizmo_interface.py:561,564,566) through the drop-in module:
    tree = ConstructKDTree(r, m, soft); GetAccelParallel(grid, tree, G, theta); GetAccelParallel([centre], ...)."""
    from oc_nbody_b200.grid_cartesian import grid
    from oc_nbody_b200.pykdgrav_compat import ConstructKDTree, GetAccelParallel, GetPotentialParallel
    from oc_nbody_b200.synthetic import make_snapshot
    snap = make_snapshot(30000, seed=11)
    r = np.concatenate([snap[s]["position"] for s in ("star", "dark", "gas")])
    m = np.concatenate([snap[s]["mass"] for s in ("star", "dark", "gas")])
    soft = np.concatenate([np.full(len(snap["star"]["mass"]), 11.2e-3), np.full(len(snap["dark"]["mass"]), 112e-3),
                           2.8e-3 * snap["gas"]["smooth.length"]])
    g = grid(0.05, 0.05, 0.05, 0.05 / 6)
    center = np.array([8.0, 0.0, 0.0])
    g.gen_evolved_grid(center)
    G_ref, theta = 4.398600413517813e-09, 0.5
    tree = ConstructKDTree(np.float64(r), np.float64(m), np.float64(soft), ctx=ctx)
    accel = GetAccelParallel(g.evolved_grid, tree, G_ref, theta)
    accel_center = GetAccelParallel(np.array([g.ss_evolved_position]), tree, G_ref, theta)
    assert accel.shape == (len(g), 3) and accel_center.shape == (1, 3)
    s32 = oracle.recentre(r, m, center)
    ref, pref = oracle.field_direct(s32, soft.astype(np.float32), oracle.recentre(g.evolved_grid, None, center),
                                    oracle.KERNEL_SPLINE, G_ref, want_pot=True)
    cond = oracle.field_direct_abs(s32, soft.astype(np.float32), oracle.recentre(g.evolved_grid, None, center),
                                   oracle.KERNEL_SPLINE, G_ref)  # 30 000-particle miniature: see oracle_field
    assert rel_err(accel.T, ref, abs_sum=cond) <= TOL
    assert rel_err(accel_center.T, ref[:, -1:], abs_sum=cond[:, -1:]) <= TOL
    # the frame subtraction as the reference writes it (gizmo_interface.py:569-571) leaves the tidal field
    tidal = accel - accel_center[0]
    want = oracle.frame_subtract(ref, g.origin_row)
    scale = np.sqrt((ref * ref).sum(axis=0)).max()
    assert np.max(np.abs(tidal.T - want)) <= 2e-6 * scale
    pot = GetPotentialParallel(g.evolved_grid, tree, G_ref, theta)
    assert np.max(np.abs(pot - pref) / np.abs(pref)) <= TOL


def test_driver_loop_follows_the_cluster(small_world, ctx):
    """oc_nbody.py:15-70 restated over the GPU codes (driver.evolve_cluster_in_galaxy): the grid origin follows the
    bound centre of mass, ejected stars leave the system, frames carry the reference's keys, energy stays put."""
    from oc_nbody_b200.driver import evolve_cluster_in_galaxy
    from oc_nbody_b200.synthetic import make_plummer_cluster
    field, _, _ = small_world
    pos_pc, vel, mass = make_plummer_cluster(800)
    center = np.array([8.0, 0.0, 0.0])
    pos = pos_pc * 1e-3 + center[:, None]
    vel = vel + np.array([[0.0], [20.0], [0.0]])       # the cluster moves through the grid: 20 km/s ~ 0.02 kpc/Myr
    pos[:, 17] += 0.3                                   # one star far outside the ejection cut
    field.evolve_grid(center)
    field.evolve_model(0.0)
    cl, rec = evolve_cluster_in_galaxy(field, mass, pos, vel, timestep=0.05, tend=0.5, softening_pc=0.01, eject_cut=100.0,
                                       ctx=ctx)
    assert len(rec.frames) == 10 and cl.n == 799 and 17 not in cl.key
    f = rec.frames[-1]
    assert set(f) == {"time", "position", "velocity", "mass", "com", "chosen_position", "chosen_velocity"}
    assert f["position"].shape == (799, 3) and abs(f["time"] - 0.45) < 1e-12
    # the grid origin is the last bound centre of mass, which moved with the cluster: 0.45 Myr x 20 km/s ~ 9 pc in y
    assert np.allclose(field._origin, f["com"]) and 0.007 < f["com"][1] - rec.frames[0]["com"][1] < 0.011
    assert abs(f["com"][0] - 8.0) < 2e-3

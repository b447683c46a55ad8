"""GPU parity: K2/K3/K5 are FP64 with separately rounded ops -> bit-exact against the oracle."""
import numpy as np
import pytest

import oracle
from util import dev

pytestmark = pytest.mark.gpu


def make_planes(rng, nodes, n_cluster, nsnap=2):
    n_node = len(nodes[0]) * len(nodes[1]) * len(nodes[2]) + 1
    acc = rng.normal(0, 1e-2, (nsnap, n_cluster, 3, n_node))
    pot = rng.normal(-50, 1.0, (nsnap, n_cluster, n_node))
    rec = np.empty((nsnap, n_cluster, n_node, 4), np.float32)
    for s in range(nsnap):
        for c in range(n_cluster):
            rec[s, c] = oracle.pack_planes(acc[s, c], pot[s, c])
    return acc, pot, rec


def star_cloud(rng, nodes, origin, n):
    """random interior points + exact nodes + faces + outside the lattice, per cluster"""
    ncl = origin.shape[0]
    scl = rng.integers(0, ncl, n).astype(np.int32)
    half = np.array([a[-1] for a in nodes])
    p = rng.uniform(-1.15, 1.15, (n, 3)) * half + origin[scl]
    k = n // 8
    for d in range(3):  # exactly on evolved nodes (node + origin, the value the kernels compare against)
        idx = rng.integers(0, len(nodes[d]), k)
        p[:k, d] = nodes[d][idx] + origin[scl[:k], d]
    p[k:2 * k, 0] = nodes[0][0] + origin[scl[k:2 * k], 0]      # low face
    p[2 * k:3 * k, 2] = nodes[2][-1] + origin[scl[2 * k:3 * k], 2]  # high face
    p[3 * k:3 * k + 3] = origin[scl[3 * k:3 * k + 3]] + 100.0 * half   # far outside
    return p, scl


@pytest.mark.parametrize("shape,ncl", [((16, 16, 16), 1), ((32, 32, 32), 5), ((7, 12, 3), 3), ((2, 2, 2), 2)])
def test_grid_interp_bit_exact(ctx, shape, ncl):
    import torch
    rng = np.random.default_rng(sum(shape) + ncl)
    nodes = [np.linspace(-L, L, n) for L, n in zip((0.6, 0.45, 0.3), shape)]
    origin = rng.normal(0, 3.0, (ncl, 3))
    _, _, rec = make_planes(rng, nodes, ncl)
    n = 20000
    p, scl = star_cloud(rng, nodes, origin, n)
    for wb in (0.0, 0.37, 1.0):
        ref_acc, ref_pot, ref_cell = oracle.grid_interp(nodes, origin, rec[0], rec[1], wb, p[:, 0], p[:, 1], p[:, 2], scl,
                                                        want_pot=True, want_cell=True)
        acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
        pot = torch.empty(n, dtype=torch.float64, device="cuda")
        cell = torch.empty((3, n), dtype=torch.int32, device="cuda")
        ctx.grid_interp(shape, [dev(a) for a in nodes], dev(origin), dev(rec[0]), dev(rec[1]), wb,
                        dev(p[:, 0].copy()), dev(p[:, 1].copy()), dev(p[:, 2].copy()), dev(scl), acc, pot, cell)
        torch.cuda.synchronize()
        assert np.array_equal(cell.cpu().numpy(), ref_cell)
        assert np.array_equal(acc.cpu().numpy(), ref_acc)
        assert np.array_equal(pot.cpu().numpy(), ref_pot)
        # independent check of the cell rule: numpy searchsorted on the evolved node arrays
        for d in range(3):
            ev = nodes[d][None, :] + origin[:, d:d + 1]
            want = np.array([np.searchsorted(ev[c], x, side="right") - 1 for c, x in zip(scl[:2000], p[:2000, d])])
            want = np.clip(want, 0, shape[d] - 2)
            assert np.array_equal(ref_cell[d, :2000], want)


def test_grid_interp_affine_field_exact(ctx):
    """Trilinear interpolation reproduces an affine field (to FP64 rounding), single snapshot, no cluster ids."""
    import torch
    rng = np.random.default_rng(5)
    nodes = [np.linspace(-0.6, 0.6, 16)] * 3
    X, Y, Z = np.meshgrid(*nodes, indexing="ij")
    A = rng.normal(0, 1, (4, 3))
    b = rng.normal(0, 1, 4)
    vals = [(A[q, 0] * X + A[q, 1] * Y + A[q, 2] * Z + b[q]).reshape(-1) for q in range(4)]
    rec = np.zeros((1, 16 ** 3 + 1, 4), np.float64)
    for q in range(4):
        rec[0, :-1, q] = vals[q]
    rec32 = rec.astype(np.float32)
    n = 5000
    p = rng.uniform(-0.6, 0.6, (n, 3))
    acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
    pot = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.grid_interp((16, 16, 16), [dev(a) for a in nodes], dev(np.zeros((1, 3))), dev(rec32[0]), None, 0.0,
                    dev(p[:, 0].copy()), dev(p[:, 1].copy()), dev(p[:, 2].copy()), None, acc, pot)
    torch.cuda.synchronize()
    got = np.concatenate([acc.cpu().numpy(), pot.cpu().numpy()[None]])
    want = A @ p.T + b[:, None]
    assert np.max(np.abs(got - want)) < 5e-7  # FP32 storage of the node values bounds the error


def test_pack_blend_kick_drift_bit_exact(ctx):
    import torch
    rng = np.random.default_rng(6)
    n = 4097
    acc = rng.normal(0, 1e-2, (3, n))
    pot = rng.normal(-10, 1, n)
    rec = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    ctx.pack_planes(dev(acc), dev(pot), rec)
    ref_rec = oracle.pack_planes(acc, pot)
    assert np.array_equal(rec.cpu().numpy(), ref_rec)
    rec_b = oracle.pack_planes(acc * 1.1 + 1e-4, pot * 0.9)
    out = torch.empty((3, n), dtype=torch.float64, device="cuda")
    outp = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.grid_time_blend(rec, dev(rec_b), 0.3, out, outp)
    ra, rp = oracle.time_blend(ref_rec, rec_b, 0.3, want_pot=True)
    assert np.array_equal(out.cpu().numpy(), ra) and np.array_equal(outp.cpu().numpy(), rp)
    vel = rng.normal(0, 1, (3, n))
    pos = rng.normal(8, 1e-3, (3, n))
    dv, dp = dev(vel), dev(pos)
    ctx.kick(dv, dev(acc), 0.05)
    ctx.drift(dp, dv, 0.1, 1.022712165045695e-3)
    torch.cuda.synchronize()
    rv = oracle.kick(vel, acc, 0.05)
    rp = oracle.drift(pos, rv, 0.1, 1.022712165045695e-3)
    assert np.array_equal(dv.cpu().numpy(), rv) and np.array_equal(dp.cpu().numpy(), rp)


def test_grid_interp_multi_bit_exact_and_cubic_field_code(ctx):
    """1..4 record planes (cubic B-spline in time = the reference's splrep/splev): kernel vs oracle bit-exact, and the
    field code in time_interpolation='cubic' mode against scipy's splev per grid point."""
    import torch
    from scipy import interpolate
    from oc_nbody_b200 import time_spline
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.units import units
    rng = np.random.default_rng(17)
    nodes = [np.linspace(-0.06, 0.06, 8)] * 3
    origin = np.array([[8.0, 0.0, 0.0]])
    n_node = 8 ** 3 + 1
    recs = rng.normal(0, 1e-2, (4, n_node, 4)).astype(np.float32)
    w = np.array([0.1, 0.55, 0.3, 0.05])
    n = 5000
    p = rng.uniform(-0.07, 0.07, (n, 3)) + origin
    for nr in (1, 2, 3, 4):
        ref_acc, ref_pot = oracle.grid_interp_multi(nodes, origin, list(recs[:nr]), w[:nr], p[:, 0], p[:, 1], p[:, 2],
                                                    want_pot=True)
        acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
        pot = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.grid_interp_multi((8, 8, 8), [dev(a) for a in nodes], dev(origin), [dev(r) for r in recs[:nr]], w[:nr],
                              dev(p[:, 0].copy()), dev(p[:, 1].copy()), dev(p[:, 2].copy()), None, acc, pot)
        torch.cuda.synchronize()
        assert np.array_equal(acc.cpu().numpy(), ref_acc) and np.array_equal(pot.cpu().numpy(), ref_pot)

    # field code, cubic mode, 6 snapshots of a smooth synthetic field installed through set_snapshot_fields
    class Snap(object):
        snapshot = {"index": 0, "time": 0.0}
    times = np.array([0.0, 22.0, 45.5, 68.0, 91.2, 113.9])
    stacks = rng.normal(0, 1e-3, (4, 1, n_node)) + np.sin(times / 35.0)[None, :, None] * rng.normal(0, 1e-2, (4, 1, n_node))
    f = gizmo_field(dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 8,
                         time_interpolation="cubic"), [Snap() for _ in times], time_in_Myr=times, build=False, ctx=ctx)
    f.set_snapshot_fields(stacks[0], stacks[1], stacks[2], pot=stacks[3])
    f.evolve_grid(origin[0])
    for t in (0.0, 30.3, 68.0, 100.0):
        f.evolve_model(t | units.Myr)
        ax, ay, az = f.get_gravity_at_point(0.0, p[:200, 0], p[:200, 1], p[:200, 2])
        got = np.stack([c.value_in(units.kms / units.Myr) for c in (ax, ay, az)])
        # the reference's way: splev of a per-node splrep (gizmo_interface.py:591-597,607-620), then trilinear
        ev = np.empty((4, n_node))
        for q in range(4):
            for i in range(n_node):
                ev[q, i] = interpolate.splev(t, interpolate.splrep(times, stacks[q][:, i]))
        rec = oracle.pack_planes(ev[:3], ev[3])
        want = oracle.grid_interp(nodes, origin, rec, None, 0.0, p[:200, 0], p[:200, 1], p[:200, 2])
        assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want))   # FP32 coefficient planes
        assert np.allclose(f.evolved_acceleration, ev[:3], rtol=1e-10, atol=1e-14)


# ------------------------------------------------------------------ nested fine grid + tidal tensor ----
def _nested_case(rng, n_planes):
    from oc_nbody_b200.grid_cartesian import grid
    g = grid(0.6, 0.45, 0.3, 0.05)
    g.add_fine_grid(0.2, 0.1, 0.12, 0.013)
    planes = rng.normal(0, 1e-2, (n_planes, 4, len(g)))
    coarse, fine = oracle.layout_nested(planes, g.n_lattice, g.coarse_keep_index, g.coarse_hole_index, g.coarse_hole_points(),
                                        g.fine_nodes, g.fine_row0)
    return g, planes, coarse, fine


def _nested_stars(rng, g, origin, n):
    half_c = np.array([0.6, 0.45, 0.3])
    half_f = np.array([0.2, 0.1, 0.12])
    p = rng.uniform(-1.1, 1.1, (n, 3)) * half_c
    p[: n // 2] = rng.uniform(-1.3, 1.3, (n // 2, 3)) * half_f
    x = p + origin
    k = n // 10
    for d in range(3):  # exactly on the faces of the fine box and on fine / coarse nodes (evolved values)
        x[:k, d] = g.fine_nodes[d][rng.integers(0, len(g.fine_nodes[d]), k)] + origin[d]
        x[k:2 * k, d] = g.nodes[d][rng.integers(0, len(g.nodes[d]), k)] + origin[d]
    x[2 * k:2 * k + 50, 0] = g.fine_nodes[0][-1] + origin[0]
    x[2 * k + 50:2 * k + 100, 2] = np.nextafter(g.fine_nodes[2][0] + origin[2], -np.inf)
    return x


@pytest.mark.parametrize("n_planes,weights", [(1, [1.0]), (2, [0.625, 0.375]), (4, [0.1, 0.55, 0.4, -0.05])])
def test_grid_interp_nested_bit_exact(ctx, n_planes, weights):
    """Two-level K3 (grid_cartesian.py:34-53,71-91) + tensor + level + cell vs the oracle, bit for bit."""
    import torch
    rng = np.random.default_rng(40 + n_planes)
    g, _, coarse, fine = _nested_case(rng, n_planes)
    origin = np.array([8.0, -0.25, 0.125])
    n = 30000
    x = _nested_stars(rng, g, origin, n)
    ref = oracle.grid_interp_nested(g.nodes, g.fine_nodes, origin, list(coarse), list(fine), weights, x[:, 0], x[:, 1], x[:, 2],
                                    want_pot=True, want_tensor=True, want_level=True, want_cell=True)
    assert 0.2 * n < ref["level"].sum() < 0.8 * n
    acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
    pot = torch.empty(n, dtype=torch.float64, device="cuda")
    ten = torch.empty((9, n), dtype=torch.float64, device="cuda")
    lev = torch.empty(n, dtype=torch.int32, device="cuda")
    cell = torch.empty((3, n), dtype=torch.int32, device="cuda")
    d_c, d_f = dev(coarse), dev(fine)
    ctx.grid_interp_nested(g.shape, [dev(a) for a in g.nodes], dev(origin[None]), list(d_c), weights, dev(x[:, 0].copy()),
                           dev(x[:, 1].copy()), dev(x[:, 2].copy()), None, acc, pot, fine_n=g.fine_shape,
                           fine_nodes=[dev(a) for a in g.fine_nodes], recs_fine=list(d_f), tensor_out=ten, level_out=lev,
                           cell_out=cell)
    torch.cuda.synchronize()
    assert np.array_equal(lev.cpu().numpy(), ref["level"])
    assert np.array_equal(cell.cpu().numpy(), ref["cell"])
    assert np.array_equal(acc.cpu().numpy(), ref["acc"])
    assert np.array_equal(pot.cpu().numpy(), ref["pot"])
    assert np.array_equal(ten.cpu().numpy(), ref["tensor"])
    # without the optional outputs the same accelerations come back (different kernel instantiation)
    acc2 = torch.empty_like(acc)
    ctx.grid_interp_nested(g.shape, [dev(a) for a in g.nodes], dev(origin[None]), list(d_c), weights, dev(x[:, 0].copy()),
                           dev(x[:, 1].copy()), dev(x[:, 2].copy()), None, acc2, None, fine_n=g.fine_shape,
                           fine_nodes=[dev(a) for a in g.fine_nodes], recs_fine=list(d_f))
    torch.cuda.synchronize()
    assert np.array_equal(acc2.cpu().numpy(), ref["acc"])


def test_grid_interp_tensor_single_level_batched(ctx):
    """Tensor output on a batch of single-level grids (no fine lattice) vs the oracle."""
    import torch
    rng = np.random.default_rng(77)
    shape, ncl = (9, 7, 5), 4
    nodes = [np.linspace(-L, L, n) for L, n in zip((0.6, 0.45, 0.3), shape)]
    origin = rng.normal(0, 3.0, (ncl, 3))
    _, _, rec = make_planes(rng, nodes, ncl)
    n = 6000
    p, scl = star_cloud(rng, nodes, origin, n)
    w = [0.25, 0.75]
    ref = oracle.grid_interp_nested(nodes, None, origin, [rec[0], rec[1]], None, w, p[:, 0], p[:, 1], p[:, 2], scl,
                                    want_tensor=True)
    acc = torch.empty((3, n), dtype=torch.float64, device="cuda")
    ten = torch.empty((9, n), dtype=torch.float64, device="cuda")
    ctx.grid_interp_nested(shape, [dev(a) for a in nodes], dev(origin), [dev(rec[0]), dev(rec[1])], w, dev(p[:, 0].copy()),
                           dev(p[:, 1].copy()), dev(p[:, 2].copy()), dev(scl), acc, tensor_out=ten)
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy(), ref["acc"])
    assert np.array_equal(ten.cpu().numpy(), ref["tensor"])

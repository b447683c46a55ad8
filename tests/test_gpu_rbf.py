"""K7 on the GPU: get_gravity_at_point / get_tidal_tensor_at_point with the reference's own spatial interpolation
(kNN(150) + RBF-PHS(phs3, order 5), gizmo_interface.py:651-756) against the CPU oracle (itself pinned against
scipy's RBFInterpolator and cKDTree in tests/test_cpu_rbf.py), through the C ABI."""
import numpy as np
import pytest

import oracle
from util import dev

pytestmark = pytest.mark.gpu


def lattice(n=16, half=0.05, origin=(8.0, 0.1, -0.2)):
    ax = np.linspace(-half, half, n)
    o = np.asarray(origin, np.float64)
    g = np.stack(np.meshgrid(ax + o[0], ax + o[1], ax + o[2], indexing="ij"), -1).reshape(-1, 3)
    return (ax, ax, ax), o, np.concatenate([g, o[None]])


def fields_of(pts, o):
    x, y, z = (pts - o).T * 25.0
    return np.stack([np.sin(x) + y * z, np.cos(y) * x - 0.2 * z, x * x - z + 0.3 * y ** 3, np.exp(0.5 * x) - y])


def run(ctx, nodes, origin, f, p, star_cluster=None, want_tensor=True, **kw):
    import torch
    n = p.shape[0]
    nclose = kw.get("nclose", 150)
    out = torch.full((f.shape[0], n), np.nan, dtype=torch.float64, device="cuda")
    tensor = torch.full((3, f.shape[0], n), np.nan, dtype=torch.float64, device="cuda") if want_tensor else None
    status = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    nb = torch.full((nclose, n), -1, dtype=torch.int64, device="cuda")
    shape = tuple(len(a) for a in nodes)
    ctx.grid_interp_rbf(shape, [dev(a) for a in nodes], dev(np.atleast_2d(origin)), dev(f), dev(p[:, 0]), dev(p[:, 1]), dev(p[:, 2]),
                        None if star_cluster is None else dev(star_cluster), out, tensor_out=tensor, status_out=status,
                        neighbors_out=nb, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), (tensor.cpu().numpy() if want_tensor else None), status.cpu().numpy() & 0xff, nb.cpu().numpy()


def close(got, ref, tol, overall=False):
    """max |got - ref| relative to the largest value of each component (or of the whole array: the tidal tensor has
    components that vanish identically)."""
    scale = np.abs(ref).max() if overall else np.abs(ref).max(axis=-1, keepdims=True)
    return np.max(np.abs(got - ref) / scale) <= tol


def test_rbf_matches_oracle_values_neighbours_and_tensor(ctx):
    nodes, o, pts = lattice()
    f = fields_of(pts, o)
    rng = np.random.default_rng(7)
    p = o + rng.uniform(-0.012, 0.012, (40, 3))
    p[0] = o                      # exactly on the origin row: shells of equidistant nodes, ties broken by index
    p[1] = pts[16 * 16 * 7 + 16 * 8 + 8]  # exactly on a lattice node
    p[2] = o + np.array([0.0, 0.0033, 0.0])  # on a symmetry plane of the lattice
    out, tensor, status, nb = run(ctx, nodes, o, f, p)
    ref = oracle.rbf_interp(nodes, o, f, p[:, 0], p[:, 1], p[:, 2], want_tensor=True, want_neighbors=True)
    assert np.all(status == 0)
    assert np.array_equal(nb, ref["neighbors"])
    assert close(out, ref["out"], 1e-9)
    assert close(tensor, ref["tensor"], 1e-7, overall=True)
    # interpolation property: exact at data points
    assert abs(out[0, 0] - f[0, -1]) <= 1e-9 * np.abs(f[0]).max()
    assert abs(out[2, 1] - f[2, 16 * 16 * 7 + 16 * 8 + 8]) <= 1e-9 * np.abs(f[2]).max()
    # values only (one right-hand side) gives the same numbers
    out1, _, status1, _ = run(ctx, nodes, o, f, p, want_tensor=False)
    assert np.all(status1 == 0) and np.array_equal(out1, out)


def test_rbf_reproduces_degree_five_polynomials(ctx):
    nodes, o, pts = lattice()
    x, y, z = (pts - o).T * 10.0
    f = np.stack([1.0 + x - 2 * y + 0.5 * z, x ** 5 - 3 * x * y ** 3 * z + z ** 4])
    rng = np.random.default_rng(9)
    p = o + rng.uniform(-0.01, 0.01, (64, 3))
    out, tensor, status, _ = run(ctx, nodes, o, f, p)
    X, Y, Z = (p - o).T * 10.0
    assert np.all(status == 0)
    assert np.allclose(out, np.stack([1.0 + X - 2 * Y + 0.5 * Z, X ** 5 - 3 * X * Y ** 3 * Z + Z ** 4]), rtol=0, atol=1e-9)
    assert np.allclose(tensor[0], 10.0 * np.stack([np.ones_like(X), 5 * X ** 4 - 3 * Y ** 3 * Z]), rtol=0, atol=1e-6)
    assert np.allclose(tensor[1], 10.0 * np.stack([-2 * np.ones_like(X), -9 * X * Y ** 2 * Z]), rtol=0, atol=1e-6)


@pytest.mark.parametrize("nclose,order,phs", [(80, 3, 5), (150, 5, 1), (120, 4, 3), (60, 2, 3), (149, 5, 7),
                                              (100, 3, 2), (80, 2, 4), (120, 4, 6), (150, 5, 8), (150, 5, 4)])
def test_rbf_other_options(ctx, nclose, order, phs):
    nodes, o, pts = lattice(n=12, half=0.06)
    f = fields_of(pts, o)[:3]
    p = o + np.random.default_rng(nclose).uniform(-0.01, 0.01, (9, 3))
    out, _, status, nb = run(ctx, nodes, o, f, p, want_tensor=False, nclose=nclose, order=order, phs=phs)
    ref = oracle.rbf_interp(nodes, o, f, p[:, 0], p[:, 1], p[:, 2], nclose=nclose, order=order, phs=phs, want_neighbors=True)
    assert np.all(status == 0)
    assert np.array_equal(nb, ref["neighbors"])
    assert close(out, ref["out"], 1e-8)


def test_rbf_batched_grids_odd_lattice_and_edges(ctx):
    # two clusters with their own origins and fields; odd node counts: the origin row duplicates the central node and is
    # left out of the neighbour search (include_origin = 0)
    ax = np.linspace(-0.05, 0.05, 11)
    origins = np.array([[8.0, 0.0, 0.0], [-3.0, 7.0, 0.5]])
    fs, ps, refs = [], [], []
    rng = np.random.default_rng(2)
    for o in origins:
        g = np.stack(np.meshgrid(ax + o[0], ax + o[1], ax + o[2], indexing="ij"), -1).reshape(-1, 3)
        pts = np.concatenate([g, o[None]])
        f = fields_of(pts, o)
        p = o + rng.uniform(-0.008, 0.008, (10, 3))
        fs.append(f), ps.append(p)
        refs.append(oracle.rbf_interp((ax, ax, ax), o, f, p[:, 0], p[:, 1], p[:, 2], include_origin=False)["out"])
    f = np.ascontiguousarray(np.stack(fs, axis=1).reshape(4, -1))  # [n_comp][n_cluster][n_node]
    p = np.concatenate(ps)
    scl = np.repeat(np.arange(2, dtype=np.int32), 10)
    import torch
    out = torch.empty((4, 20), dtype=torch.float64, device="cuda")
    status = torch.empty(20, dtype=torch.int32, device="cuda")
    ctx.grid_interp_rbf((11, 11, 11), [dev(ax)] * 3, dev(origins), dev(f), dev(p[:, 0]), dev(p[:, 1]), dev(p[:, 2]), dev(scl), out,
                        include_origin=False, status_out=status)
    assert np.all(status.cpu().numpy() & 0xff == 0)
    assert close(out.cpu().numpy(), np.concatenate(refs, axis=1), 1e-9)
    # stars at the grid edge and far outside: the stencil is one-sided and the system ill-conditioned or singular, as it is
    # for the reference; the call must come back, flag what it can detect, and still pick the true nearest neighbours
    o = origins[0]
    g = np.stack(np.meshgrid(ax + o[0], ax + o[1], ax + o[2], indexing="ij"), -1).reshape(-1, 3)
    pe = o + np.array([[0.0499, 0.001, 0.002], [0.2, 0.0, 0.0], [0.049, 0.049, 0.049]])
    nb = torch.empty((150, 3), dtype=torch.int64, device="cuda")
    oute = torch.empty((4, 3), dtype=torch.float64, device="cuda")
    ste = torch.empty(3, dtype=torch.int32, device="cuda")
    ctx.grid_interp_rbf((11, 11, 11), [dev(ax)] * 3, dev(o[None]), dev(fs[0]), dev(pe[:, 0]), dev(pe[:, 1]), dev(pe[:, 2]), None, oute,
                        include_origin=False, status_out=ste, neighbors_out=nb)
    torch.cuda.synchronize()
    from scipy.spatial import cKDTree
    _, ids = cKDTree(g).query(pe, 150)
    assert np.array_equal(np.sort(nb.cpu().numpy().T, axis=1), np.sort(ids, axis=1))


def test_rbf_argument_errors(ctx):
    import torch
    from oc_nbody_b200._lib import OcgError
    nodes, o, pts = lattice(n=8)
    f = fields_of(pts, o)
    p = o + np.zeros((2, 3))
    out = torch.empty((4, 2), dtype=torch.float64, device="cuda")
    args = ((8, 8, 8), [dev(a) for a in nodes], dev(o[None]), dev(f), dev(p[:, 0]), dev(p[:, 1]), dev(p[:, 2]), None, out)
    for kw in (dict(nclose=151), dict(order=6), dict(phs=9), dict(phs=0), dict(nclose=30, order=5), dict(phs=7, order=2),
               dict(phs=8, order=3)):
        with pytest.raises(OcgError):
            ctx.grid_interp_rbf(*args, **kw)
    small = lattice(n=4)
    with pytest.raises(OcgError):  # 65 points < nclose
        ctx.grid_interp_rbf((4, 4, 4), [dev(a) for a in small[0]], dev(o[None]), dev(fields_of(small[2], o)), dev(p[:, 0]), dev(p[:, 1]),
                            dev(p[:, 2]), None, out)


def test_field_code_rbf_mode_matches_oracle_and_drives_the_bridge(ctx):
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_plummer_cluster, make_snapshot
    from oc_nbody_b200.units import units
    center = np.array([8.0, 0.0, 0.0])
    snaps = [make_snapshot(20000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / 12,
                space_interpolation="rbf", nclose=150, basis="phs3", order=5)
    field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    field.evolve_grid(center)
    field.evolve_model(7.5 | units.Myr)
    g = field.grid
    rng = np.random.default_rng(5)
    p = center + rng.normal(0.0, 0.002, (33, 3))
    ax, ay, az = field.get_gravity_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    got = np.stack([c.value_in(units.kms / units.Myr) for c in (ax, ay, az)])
    assert np.all(field.rbf_status.cpu().numpy() & 0xff == 0)
    fld = np.concatenate([field.evolved_acceleration, field.evolved_potential[None]])
    ref = oracle.rbf_interp(g.nodes, center, fld, p[:, 0], p[:, 1], p[:, 2], want_tensor=True)
    assert close(got, ref["out"][:3], 1e-9)
    phi = field.get_potential_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    assert np.allclose(phi.value_in(units.kms ** 2) * 1.022712165045695e-3, ref["out"][3], rtol=1e-9)
    T = field.get_tidal_tensor_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    T = T.value_in(units.kms / units.Myr / units.kpc)  # [n, 3, 3], T[s][i][j] = d a_j / d x_i
    assert close(np.transpose(T, (1, 2, 0)), ref["tensor"][:, :3, :], 1e-6, overall=True)
    # the smooth field is interpolated consistently by both schemes
    tri = gizmo_field(dict(opts, space_interpolation="trilinear"), snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    tri.set_snapshot_fields(g.snapshot_acceleration_x, g.snapshot_acceleration_y, g.snapshot_acceleration_z, pot=g.snapshot_potential)
    tri.evolve_grid(center)
    tri.evolve_model(7.5 | units.Myr)
    tx, ty, tz = tri.get_gravity_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    tgot = np.stack([c.value_in(units.kms / units.Myr) for c in (tx, ty, tz)])
    assert np.max(np.abs(tgot - got)) <= 0.05 * np.abs(got).max()
    # and a BRIDGE step runs with the RBF field code as the kicker
    pos_pc, vel, mass = make_plummer_cluster(256)
    cl = cluster_code(mass, pos_pc * 1e-3 + center[:, None], vel, softening_pc=0.01, ctx=ctx)
    field.evolve_model(0.0 | units.Myr)
    system = Bridge(timestep=0.1 | units.Myr, use_threading=False, use_cuda_graph=True)
    system.add_system(cl, (field,))
    system.add_system(field)
    system.evolve_model(0.2 | units.Myr, timestep=0.1 | units.Myr)
    assert system.graph_replays == 0 and np.all(np.isfinite(cl.pos.cpu().numpy()))
    assert np.all(field.rbf_status.cpu().numpy() & 0xff == 0)


def test_field_code_rbf_on_the_nested_grid(ctx):
    """The reference's default grid (coarse lattice with a fine lattice nested around the cluster, test_options:93-100):
    the RBF kick searches both levels of the reference's point list (kept coarse + fine + origin) — also for a star beyond
    the fine-box surface, whose stencil mixes the levels."""
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    from oc_nbody_b200.units import units
    center = np.array([8.0, 0.0, 0.0])
    snaps = [make_snapshot(8000, seed=1776)]
    snaps.append(advance_snapshot(snaps[0], 23.0))
    opts = dict(grid_x_size_in_kpc=0.06, grid_y_size_in_kpc=0.06, grid_z_size_in_kpc=0.06, grid_resolution=0.06 / 8, fine_grid=True,
                grid_fine_x_size_in_kpc=0.012, grid_fine_y_size_in_kpc=0.012, grid_fine_z_size_in_kpc=0.012,
                grid_fine_resolution=0.012 / 14, space_interpolation="rbf")
    opts.update(grid_resolution=0.06 / 24, grid_fine_x_size_in_kpc=0.0135, grid_fine_y_size_in_kpc=0.0135, grid_fine_z_size_in_kpc=0.0135,
                grid_fine_resolution=0.0135 / 12)   # coarse spacing ~2x the fine one: well-conditioned mixed-level stencils
    field = gizmo_field(opts, snaps, chosen_positions=np.tile(center, (2, 1)), ctx=ctx)
    field.evolve_grid(center)
    field.evolve_model(11.0 | units.Myr)
    g = field.grid
    assert g.has_fine_grid
    rng = np.random.default_rng(8)
    p = center + rng.uniform(-0.002, 0.002, (12, 3))
    p[-1] = center + np.array([0.0145, 0.0003, -0.0002])  # just beyond the surface of the fine box: mixed-level stencil
    ax, ay, az = field.get_gravity_at_point(0 | units.kpc, p[:, 0] | units.kpc, p[:, 1] | units.kpc, p[:, 2] | units.kpc)
    got = np.stack([c.value_in(units.kms / units.Myr) for c in (ax, ay, az)])
    st = field.rbf_status.cpu().numpy() & 0xff
    assert np.all(st == 0) and field.rbf_bad_count == 0
    fld = np.concatenate([field.evolved_acceleration, field.evolved_potential[None]])
    h = g.fine_nodes[0][1] - g.fine_nodes[0][0]
    ref = oracle.rbf_interp_points(g.evolved_grid, fld, p[:, 0], p[:, 1], p[:, 2], h, want_neighbors=True)
    assert np.all(ref["neighbors"][:, :-1] >= g.fine_row0)  # inside the fine box: every neighbour a fine point or the origin row
    assert np.any(ref["neighbors"][:, -1] < g.fine_row0)    # ... the last star's stencil reaches the coarse level
    assert close(got, ref["out"][:3], 1e-9)


def test_rbf_matches_the_golden_fixture(ctx):
    """K7 against tests/golden/rbf_reference.npz — cKDTree neighbours and scipy RBFInterpolator values on the point list of
    the REAL grid class: a single lattice, the nested grid with every stencil inside the fine level, and the nested grid
    with MIXED-LEVEL stencils (stars near the fine-box surface, up to 87 of the 150 neighbours on the coarse level)."""
    import torch
    from test_cpu_rbf import _golden_cases
    mixed_seen = 0
    for g, origin, fields, stars, nbr, vals in _golden_cases():
        n = stars.shape[0]
        out = torch.empty((3, n), dtype=torch.float64, device="cuda")
        st = torch.empty(n, dtype=torch.int32, device="cuda")
        nb = torch.empty((150, n), dtype=torch.int64, device="cuda")
        sx, sy, sz = (dev(stars[:, k].copy()) for k in range(3))
        if g.has_fine_grid:
            row = np.full(int(np.prod(g.coarse_shape)), -1, np.int32)
            row[g.coarse_keep_index] = np.arange(len(g.coarse_keep_index), dtype=np.int32)
            ctx.grid_interp_rbf_nested(g.coarse_shape, [dev(a) for a in g.nodes], g.fine_shape, [dev(a) for a in g.fine_nodes],
                                       dev(origin[None]), dev(row), g.fine_row0, dev(fields), sx, sy, sz, None, out, status_out=st,
                                       neighbors_out=nb)
            mixed_seen += int((nbr < g.fine_row0).any(axis=1).sum())
        else:
            ctx.grid_interp_rbf(g.shape, [dev(a) for a in g.nodes], dev(origin[None]), dev(fields), sx, sy, sz, None, out,
                                status_out=st, neighbors_out=nb)
        assert np.all(st.cpu().numpy() & 0xff == 0)
        assert np.array_equal(np.sort(nb.cpu().numpy().T, axis=1), nbr)   # rows of the reference's point list
        assert close(out.cpu().numpy(), vals, 1e-9)
    assert mixed_seen >= 8

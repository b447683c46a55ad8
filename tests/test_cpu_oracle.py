"""CPU: the oracle against the reference's golden vectors and analytic known answers (no GPU)."""
import os

import numpy as np
import pytest
from scipy import integrate

import oracle
from util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
G = 4.3986004135e-09


# ------------------------------------------------------------------ grid layout: pinned by the reference ----
def _grid_cases():
    z = np.load(os.path.join(GOLD, "grid_reference.npz"))
    return z, int(z["n_cases"])


@pytest.mark.parametrize("k", range(5))
def test_grid_layout_matches_reference_golden(k):
    """init_grid / node arrays bit-equal to the REAL grid_cartesian.py (fixture made by tests/golden/make_golden.py)."""
    from oc_nbody_b200.grid_cartesian import grid
    z, n = _grid_cases()
    assert k < n
    lx, ly, lz, res = z["case%d_args" % k]
    xg, yg, zg, init = oracle.grid_layout(lx, ly, lz, res)
    g = grid(lx, ly, lz, res)
    g.gen_evolved_grid(np.array([8.0, -0.25, 0.125]))
    for mine in ((xg, yg, zg, init), (g.x_grid, g.y_grid, g.z_grid, g.init_grid)):
        assert np.array_equal(mine[0], z["case%d_x_grid" % k])
        assert np.array_equal(mine[1], z["case%d_y_grid" % k])
        assert np.array_equal(mine[2], z["case%d_z_grid" % k])
        assert np.array_equal(mine[3], z["case%d_init_grid" % k])
    assert (g.x_n, g.y_n, g.z_n) == tuple(z["case%d_n" % k])
    assert np.array_equal(g.evolved_grid, z["case%d_evolved_grid" % k])
    # layout contract: C order, x outer / z inner, origin appended (grid_cartesian.py:59-69)
    i, j, kk = g.x_n - 1, g.y_n // 2, 1
    assert np.array_equal(g.init_grid[(i * g.y_n + j) * g.z_n + kk], [g.x_grid[i], g.y_grid[j], g.z_grid[kk]])
    assert np.all(g.init_grid[g.origin_row] == 0.0) and g.origin_row == g.x_n * g.y_n * g.z_n
    # quirk Q1: spacing is 2L/(n-1), not `resolution`
    assert np.isclose(g.x_grid[1] - g.x_grid[0], 2 * lx / (g.x_n - 1))


def test_grid_rejects_degenerate():
    from oc_nbody_b200.grid_cartesian import grid
    with pytest.raises(ValueError):
        grid(0.6, 0.6, 0.6, 0.5)  # int(L/res) = 1 node: nothing to interpolate between
    with pytest.raises(ValueError):
        grid(0.6, 0.6, 0.6, 0.1).add_fine_grid(0.1, 0.1, 0.1, 0.09)  # one fine node
    with pytest.raises(ValueError):
        grid(0.6, 0.6, 0.6, 0.1).add_fine_grid(0.7, 0.1, 0.1, 0.01)  # fine box sticks out of the coarse one


# ------------------------------------------------------------------ nested fine grid: pinned by the reference ----
@pytest.mark.parametrize("k", range(4))
def test_nested_grid_layout_matches_reference_golden(k):
    """grid.add_fine_grid: point list bit-equal to the REAL grid_cartesian.py:34-53,71-91 (kept coarse | fine | origin),
    and the lattice bookkeeping the CUDA kernel relies on is consistent with it."""
    from oc_nbody_b200.grid_cartesian import grid
    z = np.load(os.path.join(GOLD, "grid_nested_reference.npz"))
    assert k < int(z["n_cases"])
    g = grid(*z["case%d_coarse_args" % k])
    g.add_fine_grid(*z["case%d_fine_args" % k])
    g.gen_evolved_grid(np.array([8.0, -0.25, 0.125]))
    assert np.array_equal(g.init_grid, z["case%d_init_grid" % k])
    assert np.array_equal(g.evolved_grid, z["case%d_evolved_grid" % k])
    assert (g.x_n, g.y_n, g.z_n) == tuple(z["case%d_n" % k]) == g.fine_shape  # the reference's overwrite quirk
    for mine, key in zip(g.fine_nodes, ("x", "y", "z")):
        assert np.array_equal(mine, z["case%d_%s_fine_grid" % (k, key)])
    assert np.all(g.init_grid[g.origin_row] == 0.0) and g.origin_row == len(g) - 1
    # bookkeeping: kept rows are the full coarse lattice minus the hole, in lattice order; fine rows follow
    full = g._lattice_points_(g.x_grid, g.y_grid, g.z_grid)
    assert np.array_equal(full[g.coarse_keep_index], g.init_grid[:g.fine_row0])
    assert len(g.coarse_keep_index) + len(g.coarse_hole_index) == g.n_lattice
    fine = g._lattice_points_(*g.fine_nodes)
    assert np.array_equal(fine, g.init_grid[g.fine_row0:-1])
    hole = g.coarse_hole_points()
    half = np.array([g.fine_x_size_in_kpc, g.fine_y_size_in_kpc, g.fine_z_size_in_kpc])
    assert np.all(np.abs(hole) < half) and not np.any(np.all(np.abs(g.init_grid[:g.fine_row0]) < half, axis=1))


def _nested_affine_setup(rng):
    from oc_nbody_b200.grid_cartesian import grid
    g = grid(0.6, 0.45, 0.3, 0.05)
    g.add_fine_grid(0.2, 0.1, 0.12, 0.013)
    A = rng.normal(0, 1, (4, 3))
    b = rng.normal(0, 1, 4)
    planes = (A @ g.init_grid.T + b[:, None])[None]  # [1, 4, Npoints]: an affine field sampled on the point list
    coarse, fine = oracle.layout_nested(planes, g.n_lattice, g.coarse_keep_index, g.coarse_hole_index, g.coarse_hole_points(),
                                        g.fine_nodes, g.fine_row0)
    return g, A, b, coarse, fine


def test_nested_interp_affine_field_level_rule_and_tensor():
    """Two-level trilinear interpolation reproduces an affine field on both levels and across the hole (dropped
    coarse points are filled from the fine lattice); the level rule is the closed fine box; the tensor of an affine
    field is its constant gradient, T[i][j] = d a_j / d x_i (gizmo_interface.py:719-756)."""
    rng = np.random.default_rng(11)
    g, A, b, coarse, fine = _nested_affine_setup(rng)
    origin = np.array([8.0, -0.25, 0.125])
    n = 4000
    half_c = np.array([0.6, 0.45, 0.3])
    half_f = np.array([0.2, 0.1, 0.12])
    p = rng.uniform(-1, 1, (n, 3)) * half_c
    p[: n // 2] = rng.uniform(-1.3, 1.3, (n // 2, 3)) * half_f  # around / inside the fine box
    p[0] = half_f            # corner of the fine box: inside (closed)
    p[1] = -half_f
    p[2] = half_f * [1.0, 1.0, np.nextafter(1.0, 2.0)]  # just outside along z
    x = p + origin
    out = oracle.grid_interp_nested(g.nodes, g.fine_nodes, origin, [coarse[0]], [fine[0]], [1.0], x[:, 0], x[:, 1], x[:, 2],
                                    want_pot=True, want_tensor=True, want_level=True)
    # the level rule is applied to evolved coordinates: fine iff node_f[0]+o <= x <= node_f[-1]+o on every axis
    lo = np.array([a[0] for a in g.fine_nodes]) + origin
    hi = np.array([a[-1] for a in g.fine_nodes]) + origin
    want_level = np.all((x >= lo) & (x <= hi), axis=1).astype(np.int32)
    assert np.array_equal(out["level"], want_level)
    assert want_level[0] == 1 and want_level[1] == 1 and want_level[2] == 0 and 100 < want_level.sum() < n - 100
    want = A @ p.T + b[:, None]
    scale = np.abs(want).max()
    assert np.max(np.abs(out["acc"] - want[:3])) <= 3e-6 * scale     # FP32 records
    assert np.max(np.abs(out["pot"] - want[3])) <= 3e-6 * scale
    T = out["tensor"].reshape(3, 3, n)                              # T[i][j]
    cell_f = 2 * half_f.min() / 20
    assert np.max(np.abs(T - A[:3].T[:, :, None])) <= 2e-6 * scale / cell_f * 10


def test_nested_interp_single_level_equals_plain_and_tensor_matches_finite_differences():
    rng = np.random.default_rng(12)
    nodes = [np.linspace(-0.6, 0.6, 9), np.linspace(-0.45, 0.45, 7), np.linspace(-0.3, 0.3, 5)]
    n_node = 9 * 7 * 5 + 1
    recs = [oracle.pack_planes(rng.normal(0, 1, (3, n_node)), rng.normal(0, 1, n_node)) for _ in range(2)]
    origin = np.array([1.0, 2.0, 3.0])
    n = 500
    x = rng.uniform(-0.95, 0.95, (n, 3)) * [0.6, 0.45, 0.3] + origin
    ref = oracle.grid_interp(nodes, origin, recs[0], recs[1], 0.3, x[:, 0], x[:, 1], x[:, 2])
    wb = np.float32(0.3)
    w = [float(np.float32(np.float32(1.0) - wb)), float(wb)]
    out = oracle.grid_interp_nested(nodes, None, origin, recs, None, w, x[:, 0], x[:, 1], x[:, 2], want_tensor=True,
                                    want_cell=True)
    assert np.array_equal(out["acc"], ref)
    # central differences inside the cell (the interpolant is multilinear there)
    h = 1e-6
    T = out["tensor"].reshape(3, 3, n)
    for i in range(3):
        dx = np.zeros(3)
        dx[i] = h
        ap = oracle.grid_interp_nested(nodes, None, origin, recs, None, w, *(x + dx).T, want_cell=True)
        am = oracle.grid_interp_nested(nodes, None, origin, recs, None, w, *(x - dx).T, want_cell=True)
        same = np.all(ap["cell"] == am["cell"], axis=0)
        fd = (ap["acc"] - am["acc"]) / (2 * h)
        assert same.sum() > n * 0.9
        assert np.max(np.abs(fd[:, same] - T[i][:, same])) < 1e-6 * np.abs(T).max()


# ------------------------------------------------------------------ softening kernels ----
def test_plummer_two_body_known_answer():
    src = np.array([[3.0, 4.0, 12.0, 2.0e5]], np.float32)
    soft = np.array([0.5], np.float32)
    tgt = np.zeros((1, 4), np.float32)
    acc, pot = oracle.field_direct(src, soft, tgt, oracle.KERNEL_PLUMMER, G, want_pot=True)
    r2 = 169.0 + 0.25
    assert np.allclose(acc[:, 0], G * 2.0e5 * np.array([3.0, 4.0, 12.0]) / r2 ** 1.5, rtol=1e-14)
    assert np.isclose(pot[0], -G * 2.0e5 / np.sqrt(r2), rtol=1e-14)


def _spline_density(q):
    """Cubic spline (Monaghan & Lattanzio 1985) with compact support radius h = 1, normalised to unit mass."""
    return (8.0 / np.pi) * np.where(q <= 0.5, 1 - 6 * q ** 2 + 6 * q ** 3, 2 * (1 - np.minimum(q, 1.0)) ** 3)


@pytest.mark.parametrize("q", [0.05, 0.3, 0.5, 0.7, 0.95])
def test_spline_kernel_matches_quadrature_of_the_density(q):
    """pykdgrav's ForceKernel/PotentialKernel are restated from memory [3P]: check them against an independent
    quadrature of the cubic-spline mass distribution they claim to describe."""
    h = 0.37
    r = q * h
    menc = integrate.quad(lambda x: 4 * np.pi * x * x * _spline_density(x), 0, q, epsabs=1e-14)[0]
    outer = integrate.quad(lambda x: 4 * np.pi * x * _spline_density(x), q, 1.0, epsabs=1e-14)[0]
    assert np.isclose(oracle.spline_force(r, h), menc / r ** 3, rtol=1e-10)
    assert np.isclose(oracle.spline_pot(r, h), -(menc / q + outer) / h, rtol=1e-10)


def test_spline_kernel_continuity_and_newtonian_limit():
    h = 0.2
    for q in (0.5, 1.0):
        lo, hi = q * h * (1 - 1e-13), q * h * (1 + 1e-13)
        assert np.isclose(oracle.spline_force(lo, h), oracle.spline_force(hi, h), rtol=1e-9)
        assert np.isclose(oracle.spline_pot(lo, h), oracle.spline_pot(hi, h), rtol=1e-9)
    for r in (h, 1.5 * h, 40 * h):
        assert np.isclose(oracle.spline_force(r, h), r ** -3, rtol=1e-13)
        assert np.isclose(oracle.spline_pot(r, h), -1 / r, rtol=1e-13)
    assert np.isclose(oracle.spline_pot(0.0, h), -2.8 / h)  # the Plummer-equivalent relation h = 2.8 eps


def test_coincident_pairs_contribute_nothing():
    src = np.array([[0, 0, 0, 1e9], [1, 0, 0, 1.0]], np.float32)
    tgt = np.zeros((1, 4), np.float32)
    for kern in (oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE):
        acc, pot = oracle.field_direct(src, np.zeros(2, np.float32), tgt, kern, 1.0, want_pot=True)
        assert np.allclose(acc[:, 0], [1.0, 0, 0]) and np.isclose(pot[0], -1.0)
    # Plummer with soft > 0: r = 0 gives zero force but a finite potential -m/eps
    acc, pot = oracle.field_direct(src[:1], np.array([2.0], np.float32), tgt, oracle.KERNEL_PLUMMER, 1.0, want_pot=True)
    assert np.all(acc == 0) and np.isclose(pot[0], -1e9 / 2.0)


def test_direct_sum_invariances():
    rng = np.random.default_rng(2)
    n = 500
    src = np.concatenate([rng.normal(0, 2, (n, 3)), rng.uniform(1e3, 1e5, (n, 1))], 1).astype(np.float32)
    soft = rng.uniform(0.01, 0.4, n).astype(np.float32)
    tgt = np.concatenate([rng.uniform(-0.5, 0.5, (40, 3)), np.zeros((40, 1))], 1).astype(np.float32)
    for kern in (oracle.KERNEL_PLUMMER, oracle.KERNEL_SPLINE):
        a = oracle.field_direct(src, soft, tgt, kern, G)
        perm = rng.permutation(n)
        assert rel_err(oracle.field_direct(src[perm], soft[perm], tgt, kern, G), a) < 1e-12  # permutation
        s2 = src.copy()
        s2[:, 3] *= 2
        assert rel_err(oracle.field_direct(s2, soft, tgt, kern, G), 2 * a) < 1e-13            # linear in mass
        half = oracle.field_direct(src[: n // 2], soft[: n // 2], tgt, kern, G) + \
            oracle.field_direct(src[n // 2:], soft[n // 2:], tgt, kern, G)
        assert rel_err(half, a) < 1e-12                                                        # sources additive


def test_recentre_and_frame_subtraction():
    rng = np.random.default_rng(3)
    c = np.array([8.0, 0.1, -0.2])
    pos = rng.normal(0, 1e-3, (100, 3)) + c
    r = oracle.recentre(pos, np.ones(100), c)
    assert r.dtype == np.float32 and np.array_equal(r[:, :3], (pos - c).astype(np.float32))
    # the point of H3: the recentred FP32 offsets keep ~7 digits of the OFFSET, not of the 8 kpc coordinate
    assert np.max(np.abs(r[:, :3] - (pos - c))) < 3e-10  # vs ~5e-7 for (float32)pos
    acc = rng.normal(0, 1, (3, 17))
    sub = oracle.frame_subtract(acc, 16)
    assert np.all(sub[:, 16] == 0.0) and np.array_equal(sub[:, :16], acc[:, :16] - acc[:, 16:17])


# ------------------------------------------------------------------ interpolation ----
def test_cell_selection_is_searchsorted_on_evolved_nodes():
    rng = np.random.default_rng(4)
    nodes = [np.linspace(-0.6, 0.6, 16), np.linspace(-0.3, 0.3, 7), np.linspace(-0.45, 0.45, 11)]
    origin = np.array([[8.0, -1.0, 0.3]])
    rec = np.zeros((16 * 7 * 11 + 1, 4), np.float32)
    n = 4000
    p = rng.uniform(-1.2, 1.2, (n, 3)) * np.array([0.6, 0.3, 0.45]) + origin
    for d in range(3):  # exact node hits
        p[:300, d] = nodes[d][rng.integers(0, len(nodes[d]), 300)] + origin[0, d]
    _, cell = oracle.grid_interp(nodes, origin, rec, None, 0.0, p[:, 0], p[:, 1], p[:, 2], want_cell=True)
    for d in range(3):
        want = np.clip(np.searchsorted(nodes[d] + origin[0, d], p[:, d], side="right") - 1, 0, len(nodes[d]) - 2)
        assert np.array_equal(cell[d], want)


def test_trilinear_reproduces_affine_fields_and_time_lerp():
    rng = np.random.default_rng(5)
    nodes = [np.linspace(-0.6, 0.6, 9)] * 3
    X, Y, Z = np.meshgrid(*nodes, indexing="ij")
    A, b = rng.normal(0, 1, (4, 3)), rng.normal(0, 1, 4)
    f = np.stack([(A[q, 0] * X + A[q, 1] * Y + A[q, 2] * Z + b[q]).reshape(-1) for q in range(4)], 1)
    rec_a = np.concatenate([f, np.zeros((1, 4))]).astype(np.float32)
    rec_b = (3.0 * rec_a).astype(np.float32)
    p = rng.uniform(-0.6, 0.6, (1000, 3))
    want = A @ p.T + b[:, None]
    acc, pot = oracle.grid_interp(nodes, np.zeros((1, 3)), rec_a, None, 0.0, p[:, 0], p[:, 1], p[:, 2], want_pot=True)
    assert np.max(np.abs(np.concatenate([acc, pot[None]]) - want)) < 5e-7   # FP32 node storage
    acc2 = oracle.grid_interp(nodes, np.zeros((1, 3)), rec_a, rec_b, 0.25, p[:, 0], p[:, 1], p[:, 2])
    assert np.allclose(acc2, 1.5 * acc, rtol=1e-6, atol=1e-6)
    # linear extrapolation outside the lattice (documented behaviour)
    q = np.array([[0.9, -0.8, 0.7]])
    out = oracle.grid_interp(nodes, np.zeros((1, 3)), rec_a, None, 0.0, q[:, 0], q[:, 1], q[:, 2])
    assert np.allclose(out[:, 0], (A @ q.T + b[:, None])[:3, 0], atol=2e-6)


def test_time_bracket():
    t = [0.0, 23.0, 47.0]
    assert oracle.time_bracket(t, 0.0) == (0, 1, 0.0)
    assert oracle.time_bracket(t, 23.0) == (1, 2, 0.0)
    i, j, w = oracle.time_bracket(t, 35.0)
    assert (i, j) == (1, 2) and np.isclose(w, 0.5)
    assert oracle.time_bracket(t, 100.0) == (1, 2, 1.0) and oracle.time_bracket(t, -5.0) == (0, 1, 0.0)
    assert oracle.time_bracket([5.0], 7.0) == (0, 0, 0.0)


def test_reference_time_spline_golden_is_reproduced_by_scipy():
    """The reference's own time interpolation (splrep/splev per grid point, gizmo_interface.py:587-620) collapses to
    one vectorised make_interp_spline over all grid points; pinned by the fixture generated with the
    reference's exact calls."""
    from scipy.interpolate import make_interp_spline
    z = np.load(os.path.join(GOLD, "time_spline_reference.npz"))
    spl = make_interp_spline(z["times"], z["series"], k=3)
    assert np.allclose(spl(z["t_eval"]), z["values"], rtol=1e-12, atol=1e-14)
    # the product's host module: one fit, then four basis weights per time (what the GPU blend consumes)
    from oc_nbody_b200 import time_spline
    knots, coef = time_spline.fit(z["times"], z["series"])
    for k, x in enumerate(z["t_eval"]):
        first, w = time_spline.basis(knots, x)
        assert np.isclose(w.sum(), 1.0, rtol=1e-14)
        assert np.allclose((w[:, None] * coef[first:first + 4]).sum(axis=0), z["values"][k], rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        time_spline.fit(z["times"][:3], z["series"][:3])
    # and the north_star's linear-in-time substitute differs from it at the percent level (documented gap)
    lin = np.stack([np.interp(z["t_eval"], z["times"], z["series"][:, i]) for i in range(z["series"].shape[1])], 1)
    assert 1e-4 < np.max(np.abs(lin - z["values"])) < 0.2


# ------------------------------------------------------------------ self gravity + BRIDGE ----
def test_self_gravity_momentum_and_segments():
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos, vel, m = make_plummer_cluster(300)
    m = m * np.random.default_rng(6).uniform(0.5, 2.0, 300)
    pos_kpc = pos * 1e-3 + np.array([[8.0], [0.0], [0.0]])
    acc, pot = oracle.self_gravity(pos_kpc, m, (0.01e-3) ** 2, G, want_pot=True)
    f = (acc * m).sum(axis=1)
    assert np.max(np.abs(f)) < 1e-6 * np.abs(acc * m).sum()      # Newton's third law (FP32-rounded offsets)
    assert np.all(pot < 0)
    seg = np.array([0, 100, 300])
    a2 = oracle.self_gravity(pos_kpc, m, (0.01e-3) ** 2, G, seg_offsets=seg)
    a_first = oracle.self_gravity(pos_kpc[:, :100], m[:100], (0.01e-3) ** 2, G)
    assert np.array_equal(a2[:, :100], a_first)
    a_sh = oracle.self_gravity(pos_kpc, m, (0.01e-3) ** 2, G, t0=50, t1=120)
    assert np.array_equal(a_sh[:, 50:120], acc[:, 50:120]) and np.all(a_sh[:, :50] == 0)


def test_bridge_step_conserves_energy_and_is_time_reversible():
    from oc_nbody_b200.synthetic import make_plummer_cluster
    from oc_nbody_b200.units import G_PC_KMS2, KMS_TO_PC_PER_MYR
    pos, vel, m = make_plummer_cluster(128)
    eps2 = 0.05 ** 2

    def energy(x, v):
        _, phi = oracle.self_gravity(x, m, eps2, G_PC_KMS2, want_pot=True)
        return 0.5 * (m * (v * v).sum(0)).sum() + 0.5 * (m * phi).sum()

    def no_tide(x):
        return np.zeros_like(x)
    e0 = energy(pos, vel)
    x, v = pos, vel
    for _ in range(20):
        x, v = oracle.bridge_step(x, v, m, eps2, G_PC_KMS2, 0.005, no_tide, KMS_TO_PC_PER_MYR)
    assert abs(energy(x, v) - e0) < 2e-4 * abs(e0)
    for _ in range(20):
        x, v = oracle.bridge_step(x, -v, m, eps2, G_PC_KMS2, 0.005, no_tide, KMS_TO_PC_PER_MYR)
        v = -v
    assert np.max(np.abs(x - pos)) < 1e-7

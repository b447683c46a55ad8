"""Pins the oracle's restatement of the reference's own spatial interpolation (gizmo_interface.py:651-717:
cKDTree.query(nclose) + rbf.interpolate.RBFInterpolant(basis=phs3, order=5)).  The `rbf` package (T. Hines) is not in
this image; scipy.interpolate.RBFInterpolator — the same author's port of it into scipy — with kernel='cubic' (r^3)
and degree=5 is the same interpolant, and cKDTree is the reference's own neighbour search."""
import numpy as np
import pytest

import oracle


def lattice(n=12, half=0.06, origin=(8.0, 0.1, -0.2)):
    ax = np.linspace(-half, half, n)
    o = np.asarray(origin)
    g = np.stack(np.meshgrid(ax + o[0], ax + o[1], ax + o[2], indexing="ij"), -1).reshape(-1, 3)
    return (ax, ax, ax), o, np.concatenate([g, o[None]])


def smooth_fields(pts):
    x, y, z = (pts - pts[-1]).T * 20.0
    return np.stack([np.sin(x) + y * z, np.cos(y) * x, x * x - z + 0.3 * y ** 3, np.exp(0.5 * x) - y])


def test_rbf_oracle_matches_scipy_rbfinterpolator_and_ckdtree():
    from scipy.interpolate import RBFInterpolator
    from scipy.spatial import cKDTree
    nodes, o, pts = lattice()
    f = smooth_fields(pts)
    rng = np.random.default_rng(4)
    p = o + rng.uniform(-0.012, 0.012, (12, 3))
    res = oracle.rbf_interp(nodes, o, f, p[:, 0], p[:, 1], p[:, 2], want_neighbors=True)
    # the reference's neighbour search (gizmo_interface.py:654,664)
    _, ids = cKDTree(pts).query(p, 150)
    assert np.array_equal(np.sort(ids, axis=1), np.sort(res["neighbors"].T, axis=1))
    for c in range(4):
        ref = RBFInterpolator(pts, f[c], neighbors=150, kernel="cubic", degree=5)(p)
        assert np.allclose(res["out"][c], ref, rtol=1e-9, atol=1e-11)
    # other odd polyharmonic splines of options.py:178-246 and a lower order
    res5 = oracle.rbf_interp(nodes, o, f[:1], p[:3, 0], p[:3, 1], p[:3, 2], nclose=80, order=3, phs=5)
    ref5 = RBFInterpolator(pts, f[0], neighbors=80, kernel="quintic", degree=3)(p[:3])
    # scipy's quintic kernel is -r^5: the interpolant is the same, only the sign of the weights differs
    assert np.allclose(res5["out"][0], ref5, rtol=1e-8, atol=1e-10)


def test_rbf_even_polyharmonic_splines():
    """phs2/4/6/8 of options.py:178-202 (r^k log r).  phs2 is scipy's thin_plate_spline kernel; the others have no scipy
    counterpart and are pinned through the property that fixes them: with order >= k/2 the interpolant does not depend on
    the length unit (the oracle works in coordinates scaled by the spacing h; scaling by 1 or by 7 h must give the same
    values), it reproduces polynomials up to its order, and its gradient is the finite-difference one."""
    from scipy.interpolate import RBFInterpolator
    nodes, o, pts = lattice()
    f = smooth_fields(pts)
    rng = np.random.default_rng(14)
    p = o + rng.uniform(-0.012, 0.012, (4, 3))
    res2 = oracle.rbf_interp(nodes, o, f, p[:, 0], p[:, 1], p[:, 2], nclose=100, order=3, phs=2, want_tensor=True)
    for c in range(4):
        ref = RBFInterpolator(pts, f[c], neighbors=100, kernel="thin_plate_spline", degree=3)(p)
        assert np.allclose(res2["out"][c], ref, rtol=1e-9, atol=1e-11)
    h = nodes[0][1] - nodes[0][0]
    for phs, order, nclose in ((2, 1, 60), (4, 2, 80), (6, 4, 120), (8, 5, 150)):
        a = oracle.rbf_interp_points(pts, f, p[:, 0], p[:, 1], p[:, 2], h, nclose, order, phs, want_tensor=True)
        b = oracle.rbf_interp_points(pts, f, p[:, 0], p[:, 1], p[:, 2], 7.0 * h, nclose, order, phs, want_tensor=True)
        scale = np.abs(f).max()
        assert np.max(np.abs(a["out"] - b["out"])) <= 1e-7 * scale, phs
        assert np.max(np.abs(a["tensor"] - b["tensor"])) <= 1e-5 * np.abs(a["tensor"]).max(), phs
        # gradient against central differences of the interpolant itself (same stencil: the step is tiny)
        eps = 1e-7
        for ax in range(3):
            dp = np.zeros(3)
            dp[ax] = eps
            hi = oracle.rbf_interp_points(pts, f, p[:1, 0] + dp[0], p[:1, 1] + dp[1], p[:1, 2] + dp[2], h, nclose, order, phs)["out"]
            lo = oracle.rbf_interp_points(pts, f, p[:1, 0] - dp[0], p[:1, 1] - dp[1], p[:1, 2] - dp[2], h, nclose, order, phs)["out"]
            fd = (hi[:, 0] - lo[:, 0]) / (2 * eps)
            assert np.allclose(a["tensor"][ax][:, 0], fd, rtol=2e-4, atol=2e-4 * np.abs(a["tensor"]).max()), (phs, ax)
    # polynomial reproduction with an even spline
    x, y, z = (pts - o).T * 10.0
    poly = np.stack([1.0 + x - 2 * y + 0.5 * z, x * y * z - y ** 2])
    X, Y, Z = (p - o).T * 10.0
    got = oracle.rbf_interp(nodes, o, poly, p[:, 0], p[:, 1], p[:, 2], nclose=120, order=3, phs=4)["out"]
    assert np.allclose(got, np.stack([1.0 + X - 2 * Y + 0.5 * Z, X * Y * Z - Y ** 2]), rtol=0, atol=1e-9)


def test_rbf_reproduces_polynomials_up_to_its_order_and_their_gradients():
    nodes, o, pts = lattice()
    x, y, z = (pts - o).T * 10.0
    f = np.stack([1.0 + x - 2 * y + 0.5 * z, x * y * z - y ** 2, x ** 5 - 3 * x * y ** 3 * z + z ** 4, x ** 2 * z ** 3])
    p = o + np.array([[0.003, -0.004, 0.0012], [0.011, 0.002, -0.007]])
    res = oracle.rbf_interp(nodes, o, f, p[:, 0], p[:, 1], p[:, 2], want_tensor=True)
    X, Y, Z = (p - o).T * 10.0
    exact = np.stack([1.0 + X - 2 * Y + 0.5 * Z, X * Y * Z - Y ** 2, X ** 5 - 3 * X * Y ** 3 * Z + Z ** 4, X ** 2 * Z ** 3])
    assert np.allclose(res["out"], exact, rtol=0, atol=1e-9)
    gx = 10.0 * np.stack([np.ones_like(X), Y * Z, 5 * X ** 4 - 3 * Y ** 3 * Z, 2 * X * Z ** 3])
    gz = 10.0 * np.stack([0.5 * np.ones_like(X), X * Y, -3 * X * Y ** 3 + 4 * Z ** 3, 3 * X ** 2 * Z ** 2])
    assert np.allclose(res["tensor"][0], gx, rtol=0, atol=1e-6)
    assert np.allclose(res["tensor"][2], gz, rtol=0, atol=1e-6)


def test_rbf_neighbour_order_and_ties():
    nodes, o, pts = lattice(n=8, half=0.035)
    f = smooth_fields(pts)[:1]
    # a star exactly on the grid origin: the origin row is its nearest neighbour (distance 0), then shells of equidistant
    # lattice nodes, ordered by point index inside a shell
    res = oracle.rbf_interp(nodes, o, f, [o[0]], [o[1]], [o[2]], nclose=100, want_neighbors=True)
    nb = res["neighbors"][:, 0]
    assert nb[0] == len(pts) - 1
    d = np.linalg.norm(pts[nb] - o, axis=1)
    assert np.all(np.diff(d) >= -1e-15)
    shell = nb[1:9]
    assert np.all(np.diff(shell) > 0) and np.allclose(d[1:9], d[1])
    assert res["out"][0, 0] == pytest.approx(f[0, -1], abs=1e-12)  # interpolation: exact at a data point


def _golden_cases():
    import os
    from oc_nbody_b200.grid_cartesian import grid
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbf_reference.npz"))
    origin = z["origin"]
    for k in range(int(z["n_cases"])):
        g = grid(*z["case%d_coarse_args" % k])
        fargs = z["case%d_fine_args" % k]
        if fargs[3] > 0:
            g.add_fine_grid(*fargs)
        g.gen_evolved_grid(origin)
        yield g, origin, z["case%d_fields" % k], z["case%d_stars" % k], z["case%d_neighbors" % k], z["case%d_values" % k]


def test_rbf_oracle_matches_the_golden_fixture():
    """tests/golden/rbf_reference.npz (made by make_golden.py from the REAL grid class, cKDTree and scipy's RBFInterpolator):
    single lattice and the nested grid."""
    n = 0
    for g, origin, fields, stars, nbr, vals in _golden_cases():
        h = (g.fine_nodes if g.has_fine_grid else g.nodes)[0]
        res = oracle.rbf_interp_points(g.evolved_grid, fields, stars[:, 0], stars[:, 1], stars[:, 2], h[1] - h[0], want_neighbors=True)
        assert np.array_equal(np.sort(res["neighbors"].T, axis=1), nbr)
        assert np.max(np.abs(res["out"] - vals) / np.abs(vals).max(axis=1, keepdims=True)) <= 1e-9
        n += 1
    assert n == 3

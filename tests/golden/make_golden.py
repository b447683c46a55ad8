"""Generates tests/golden/*.npz from the REAL reference, imported in the build container.

Run here (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference: tests only read the committed .npz files.

What can be generated: only `grid_cartesian.py` is importable (numpy only).  Everything else on the hot path
needs amuse / pykdgrav / rbf / gizmo_analysis, which are absent, and gizmo_interface.py cannot be imported at all
(analysis.py:573 SyntaxError) — SURVEY.md §0.3.  The reference's time interpolation is scipy's splrep/splev
(gizmo_interface.py:587-597,38-44), which IS runnable: a fixture of it pins the cubic-in-time "next" row.  So is its
neighbour search (scipy's cKDTree over the real grid class's point list); the RBF interpolant itself comes from scipy's
RBFInterpolator, the `rbf` author's own port: rbf_reference.npz pins the reference-algorithm kick (K7).
"""
import importlib.util
import os

import numpy as np
from scipy import interpolate

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/grid_cartesian.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_grid_cartesian", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = [(0.6, 0.6, 0.6, 0.6 / 16), (0.6, 0.6, 0.6, 0.6 / 4), (0.6, 0.3, 0.45, 0.011), (1.0, 1.0, 1.0, 0.13),
             (0.6, 0.6, 0.6, 0.005 * 6)]
    out = {}
    for k, (lx, ly, lz, res) in enumerate(cases):
        g = ref.grid(lx, ly, lz, res)
        g.gen_evolved_grid(np.array([8.0, -0.25, 0.125]))
        out["case%d_args" % k] = np.array([lx, ly, lz, res])
        out["case%d_n" % k] = np.array([g.x_n, g.y_n, g.z_n])
        out["case%d_x_grid" % k] = g.x_grid
        out["case%d_y_grid" % k] = g.y_grid
        out["case%d_z_grid" % k] = g.z_grid
        out["case%d_init_grid" % k] = g.init_grid
        out["case%d_evolved_grid" % k] = g.evolved_grid
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "grid_reference.npz"), **out)

    # the reference's whole-grid cache (gizmo_interface.py:510: pickle.dump(self.grid, ..., protocol=4)) as its own class
    # writes it: the real class registered under its real module path, a nested grid with stacked fields of 3 snapshots
    import pickle
    import sys
    import types
    spec2 = importlib.util.spec_from_file_location("oceanic.grid_cartesian", REF)
    mod = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(mod)
    pkg = types.ModuleType("oceanic")
    pkg.grid_cartesian = mod
    sys.modules["oceanic"], sys.modules["oceanic.grid_cartesian"] = pkg, mod
    g = mod.grid(0.06, 0.06, 0.06, 0.012)
    g.add_fine_grid(0.02, 0.02, 0.02, 0.005)
    g.gen_evolved_grid(np.array([8.0, -0.25, 0.125]))
    rng = np.random.default_rng(1776)
    for c in "xyz":
        setattr(g, "snapshot_acceleration_" + c, rng.normal(0.0, 1e-2, (3, len(g.init_grid))))
    with open(os.path.join(HERE, "grid_cache_reference.pickle"), "wb") as fh:
        pickle.dump(g, fh, protocol=4)
    del sys.modules["oceanic"], sys.modules["oceanic.grid_cartesian"]

    # nested fine grid (grid.add_fine_grid, grid_cartesian.py:34-53,71-91): coarse args, fine args
    nested = [((0.6, 0.6, 0.6, 0.05), (0.1, 0.1, 0.1, 0.01)), ((0.6, 0.3, 0.45, 0.03), (0.2, 0.1, 0.12, 0.013)),
              ((0.6, 0.6, 0.6, 0.02), (0.02, 0.02, 0.02, 0.0021)), ((0.6, 0.6, 0.6, 0.6 / 16), (0.25, 0.25, 0.25, 0.25 / 9))]
    out = {}
    for k, (cargs, fargs) in enumerate(nested):
        g = ref.grid(*cargs)
        g.add_fine_grid(*fargs)
        g.gen_evolved_grid(np.array([8.0, -0.25, 0.125]))
        out["case%d_coarse_args" % k] = np.array(cargs)
        out["case%d_fine_args" % k] = np.array(fargs)
        out["case%d_n" % k] = np.array([g.x_n, g.y_n, g.z_n])  # overwritten with the fine counts by the reference
        out["case%d_x_fine_grid" % k] = g.x_fine_grid
        out["case%d_y_fine_grid" % k] = g.y_fine_grid
        out["case%d_z_fine_grid" % k] = g.z_fine_grid
        out["case%d_init_grid" % k] = g.init_grid
        out["case%d_evolved_grid" % k] = g.evolved_grid
    out["n_cases"] = np.array(len(nested))
    np.savez_compressed(os.path.join(HERE, "grid_nested_reference.npz"), **out)

    # the reference's time interpolation mechanism, as-is: splrep per grid point, splev at t
    rng = np.random.default_rng(1776)
    times = np.array([0.0, 22.0, 45.5, 68.0, 91.2, 113.9])
    series = rng.normal(0.0, 1e-2, (len(times), 40)) + np.sin(times[:, None] / 35.0)
    t_eval = np.array([0.0, 3.3, 22.0, 50.1, 100.0, 113.9])
    vals = np.empty((len(t_eval), series.shape[1]))
    for i in range(series.shape[1]):
        tck = interpolate.splrep(times, series[:, i])
        vals[:, i] = [float(interpolate.splev(t, tck)) for t in t_eval]
    np.savez_compressed(os.path.join(HERE, "time_spline_reference.npz"), times=times, series=series, t_eval=t_eval,
                        values=vals)
    # the reference's spatial interpolation (gizmo_interface.py:651-717) on the REAL grid class: cKDTree.query(150) over
    # grid.evolved_grid — the reference's own neighbour search — and the RBF interpolant through the neighbours.  The `rbf`
    # package is absent; scipy.interpolate.RBFInterpolator(kernel="cubic", degree=5) is the same author's port of
    # rbf.interpolate.RBFInterpolant(phs3, order=5) into scipy and evaluates the same interpolant.
    from scipy.interpolate import RBFInterpolator
    from scipy.spatial import cKDTree
    out = {}
    rbf_cases = [((0.05, 0.05, 0.05, 0.05 / 16), None), ((0.06, 0.06, 0.06, 0.06 / 8), (0.012, 0.012, 0.012, 0.012 / 14)),
                 # coarse spacing ~2x the fine one: stars near the fine-box surface get well-conditioned MIXED-LEVEL stencils
                 ((0.06, 0.06, 0.06, 0.06 / 24), (0.0135, 0.0135, 0.0135, 0.0135 / 12))]
    origin = np.array([8.0, -0.25, 0.125])
    for k, (cargs, fargs) in enumerate(rbf_cases):
        g = ref.grid(*cargs)
        if fargs is not None:
            g.add_fine_grid(*fargs)
        g.gen_evolved_grid(origin)
        pts = g.evolved_grid
        x, y, z = ((pts - origin) * 40.0).T
        fields = np.stack([np.sin(x) + y * z, np.cos(y) * x - 0.2 * z, x * x - z + 0.3 * y ** 3])   # ax, ay, az stand-ins
        rng = np.random.default_rng(1776 + k)
        half = 0.012 if fargs is None else (0.0035 if k == 1 else 0.006)
        stars = origin + rng.uniform(-half, half, (24, 3))
        if k == 2:
            # MIXED-LEVEL stencils: stars within a few fine cells of the fine-box surface (inside and outside it), whose 150
            # nearest points come from both the fine lattice and the kept coarse points (grid_cartesian.py:71-91)
            L, hf = fargs[0], 2 * fargs[0] / (g.x_n - 1)
            stars = stars[:8]
            near = np.array([[L - 0.4 * hf, 0.3 * hf, -0.2 * hf], [-(L - 1.7 * hf), L - 0.9 * hf, 0.0], [L + 0.6 * hf, 0.1 * hf, 0.2 * hf],
                             [L - 0.2 * hf, L - 0.3 * hf, L - 0.1 * hf], [0.5 * hf, -(L + 1.5 * hf), 2.2 * hf],
                             [L - 2.5 * hf, -3.1 * hf, L - 0.5 * hf], [-(L + 0.2 * hf), -(L + 0.3 * hf), 0.4 * hf],
                             [L + 2.0 * hf, L + 1.0 * hf, -(L + 0.5 * hf)]])
            stars = np.concatenate([stars, origin + near])
        tree = cKDTree(pts)                                    # gizmo_interface.py:642
        _, ids = tree.query(stars, 150)                        # :654,664
        vals = np.empty((3, len(stars)))
        for s in range(len(stars)):
            for c in range(3):                                 # :666-674, 698-704
                vals[c, s] = RBFInterpolator(pts[ids[s]], fields[c][ids[s]], kernel="cubic", degree=5)(stars[s][None])[0]
        out["case%d_coarse_args" % k] = np.array(cargs)
        out["case%d_fine_args" % k] = np.array(fargs if fargs is not None else [0.0, 0.0, 0.0, 0.0])
        out["case%d_fields" % k] = fields
        out["case%d_stars" % k] = stars
        out["case%d_neighbors" % k] = np.sort(ids, axis=1)
        out["case%d_values" % k] = vals
    out["n_cases"] = np.array(len(rbf_cases))
    out["origin"] = origin
    np.savez_compressed(os.path.join(HERE, "rbf_reference.npz"), **out)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()

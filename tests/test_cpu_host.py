"""CPU: host-side logic of the plugin mirror and the C-ABI surface (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import oc_nbody_b200
    from oc_nbody_b200._lib import ABI_SYMBOLS
    L = oc_nbody_b200.load_library()
    header = open(os.path.join(ROOT, "include", "ocg.h")).read()
    declared = sorted(set(re.findall(r"\b(ocg_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ocg_version() == 200


def test_library_exports_nothing_but_the_declared_c_abi():
    """Built with -fvisibility=hidden: the dynamic symbol table holds exactly include/ocg.h + include/ocg_debug.h — no
    mangled internals, no undeclared hooks."""
    import subprocess
    from oc_nbody_b200._lib import ABI_SYMBOLS, DEBUG_SYMBOLS, LIB_PATH
    dbg = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ocg_debug.h")).read(), flags=re.S)
    declared_dbg = sorted(set(re.findall(r"\b(ocg_[a-z0-9_]+)\s*\(", dbg)))
    assert declared_dbg == sorted(DEBUG_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if " T " in ln or " W " in ln or " B " in ln or " D " in ln)
    assert exported == sorted(ABI_SYMBOLS + DEBUG_SYMBOLS)


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run instead of silently computing on the host."""
    import torch
    import oc_nbody_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(oc_nbody_b200.OcgError, match="no CPU fallback"):
        oc_nbody_b200.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oc_nbody_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "ocg_oracle" not in text.replace("oracle/ocg_oracle.c", ""), f


def test_units_shim():
    from oc_nbody_b200.units import G_KPC_KMS_MYR, G_PC_KMS2, Quantity, to_value, units
    q = np.array([1.0, 2.0]) | units.parsec
    assert isinstance(q, Quantity) and np.allclose(q.value_in(units.kpc), [1e-3, 2e-3])
    assert np.isclose(((0.01 | units.parsec) ** 2).value_in(units.kpc ** 2), 1e-10)
    assert np.isclose((1.0 | units.kms).value_in(units.kpc / units.Myr), 1.022712165045695e-3)
    assert [3.0, 4.0] | units.kms  # a plain list, as the reference returns (gizmo_interface.py:706)
    with pytest.raises(ValueError):
        q.value_in(units.Myr)
    assert to_value(3.5, units.kpc) == 3.5
    # G in the unit of gizmo_interface.py:70, and in pc (km/s)^2/Msun
    assert np.isclose(G_KPC_KMS_MYR, 4.3986004e-09, rtol=1e-7) and np.isclose(G_PC_KMS2, 4.30091727e-3, rtol=1e-7)


def test_synthetic_inputs_are_seeded_and_shaped_like_gizmo():
    from oc_nbody_b200.synthetic import advance_snapshot, make_plummer_cluster, make_snapshot
    a, b = make_snapshot(5000), make_snapshot(5000)
    for sp in ("star", "dark", "gas"):
        assert np.array_equal(a[sp]["position"], b[sp]["position"]) and np.array_equal(a[sp]["mass"], b[sp]["mass"])
        assert a[sp].prop("host.distance.principal").shape[1] == 3
    assert len(a["star"]["mass"]) == 1000 and len(a["gas"]["mass"]) == 1500 and len(a["dark"]["mass"]) == 2500
    assert "smooth.length" in a["gas"] and a.snapshot["index"] == 577
    assert np.max(np.linalg.norm(a["dark"]["position"], axis=1)) <= 50.0
    c = advance_snapshot(a, 23.0)
    assert np.allclose(np.hypot(*c["star"]["position"][:, :2].T), np.hypot(*a["star"]["position"][:, :2].T))
    assert np.isclose(c.snapshot["time"] - a.snapshot["time"], 0.023)
    pos, vel, m = make_plummer_cluster(2000)
    assert pos.shape == (3, 2000) and np.allclose(pos.mean(axis=1), 0, atol=1e-12)
    r = np.sqrt((pos ** 2).sum(0))
    assert 0.6 < np.median(r) / 0.8 < 1.7  # half-mass radius of a Plummer sphere ~ 1.3 a


def _field(opts=None, nsnap=3):
    from oc_nbody_b200.gizmo_field import gizmo_field
    from oc_nbody_b200.synthetic import advance_snapshot, make_snapshot
    snaps = [make_snapshot(2000)]
    for _ in range(nsnap - 1):
        snaps.append(advance_snapshot(snaps[-1], 23.0))
    return gizmo_field(opts or {}, snaps, chosen_id=7, build=False), snaps


def test_source_assembly_follows_the_reference_rules():
    """gizmo_interface.py:515-558: star(-chosen)+dark+gas order, softening rules, the Q4 fix."""
    f, snaps = _field()
    r, m, h = f._source_arrays_(snaps[0])
    ns, nd, ng = 400 - 1, 1000, 600
    assert r.shape == (ns + nd + ng, 3) and m.shape == h.shape == (ns + nd + ng,)
    assert r.dtype == m.dtype == h.dtype == np.float64
    assert np.array_equal(r[ns], snaps[0]["dark"]["position"][0])
    assert np.allclose(h[:ns], 11.2e-3) and np.allclose(h[ns:ns + nd], 112e-3)
    assert np.allclose(h[ns + nd:], 2.8e-3 * snaps[0]["gas"]["smooth.length"])
    assert 7 not in snaps[0]["star"]["id"][np.where(snaps[0]["star"]["id"] != 7)[0]]
    f2, _ = _field(dict(star_char_mass=7100.0, dark_char_mass=35000.0, softening_kernel="plummer"))
    r2, m2, h2 = f2._source_arrays_(snaps[0])
    assert h2.shape == m2.shape  # Q4: the reference's un-indexed branch is one entry too long
    assert np.allclose(h2[:ns], np.cbrt(m2[:ns] / 7100.0) / 1000.0 / 2.8)


def test_evolve_model_brackets_and_follows_the_grid():
    from oc_nbody_b200.units import units
    f, _ = _field()
    assert np.allclose(f.time_in_Myr, [0.0, 23.0, 46.0])
    f.evolve_model(34.5 | units.Myr)
    assert f._bracket[:2] == (1, 2) and np.isclose(f._bracket[2], 0.5)
    f.evolve_model(1e3 | units.Myr)
    assert f._bracket == (1, 2, 1.0)
    f.evolve_model(11.5)  # bare float = Myr
    assert f._bracket[:2] == (0, 1) and np.isclose(f._bracket[2], 0.5)
    with pytest.raises(ValueError):
        _field(dict(softening_kernel="cubic"))
    with pytest.raises(ValueError):
        _field(dict(fine_grid=True))  # options.py:113-118: fine_grid comes with its four grid_fine_* values
    f, _ = _field(dict(fine_grid=True, grid_fine_x_size_in_kpc=0.1, grid_fine_y_size_in_kpc=0.1, grid_fine_z_size_in_kpc=0.1,
                    grid_fine_resolution=0.02))
    g = f._make_grid_()
    assert g.has_fine_grid and g.fine_shape == (5, 5, 5) and g.origin_row == len(g) - 1


def test_bridge_orders_kicks_and_drifts_like_amuse():
    """K(dt/2) D(dt) K(dt/2) with the field's time advanced inside the drift; first call at t = 0 is a no-op
    (oc_nbody.py:55-56)."""
    from oc_nbody_b200.bridge import Bridge
    from oc_nbody_b200.units import units
    log = []

    class Cluster(object):
        class P(object):
            x = y = z = np.zeros(3) | units.kpc
        particles = P()

        def kick_velocities(self, ax, ay, az, dt):
            log.append(("kick", round(dt, 6), float(ax.value_in(units.kms / units.Myr)[0])))

        def evolve_model(self, t):
            log.append(("drift-cluster", round(t.value_in(units.Myr), 6)))

    class Field(object):
        t = 0.0

        def get_gravity_at_point(self, eps, x, y, z):
            a = np.full(3, self.t) | units.kms / units.Myr
            return a, a, a

        def evolve_model(self, t):
            self.t = t.value_in(units.Myr)
            log.append(("drift-field", round(self.t, 6)))

    c, f = Cluster(), Field()
    b = Bridge(timestep=0.1 | units.Myr, use_threading=False)
    b.add_system(c, (f,))
    b.add_system(f)
    b.evolve_model(0.0 | units.Myr, timestep=0.1 | units.Myr)
    assert log == []
    b.evolve_model(0.1 | units.Myr, timestep=0.1 | units.Myr)
    assert log == [("kick", 0.05, 0.0), ("drift-cluster", 0.1), ("drift-field", 0.1), ("kick", 0.05, 0.1)]
    with pytest.raises(NotImplementedError):
        Bridge(use_threading=True)


def test_shard_bounds():
    from oc_nbody_b200.distributed import shard_bounds, shard_range
    b = shard_bounds(10, 4)
    assert list(b) == [0, 3, 6, 8, 10]
    assert shard_range(10_000_000, 7, 8) == (8750000, 10000000)
    assert list(shard_bounds(3, 5)) == [0, 1, 2, 3, 3, 3]


def test_snapshot_cache_names_and_format_follow_the_reference(tmp_path):
    """gizmo_interface.py:373-391 (name), :441-446 ('snapshot' -> 'snapshot_x/_y/_z'), :454-459 (pickle protocol 4 of one
    FP64 array per file): files written the reference's way are found and read; ours are readable the reference's way."""
    import pickle
    opts = dict(cache_directory=str(tmp_path), sim_name="m12i_res7100", grid_seed=1776, Rmax=50.0, theta=0.5, startnum=577,
                endnum=585, num_prior=3, grid_x_size_in_kpc=0.6, grid_y_size_in_kpc=0.6, grid_z_size_in_kpc=0.6)
    f, _ = _field(opts)
    name, path = f._grid_cache_name_()
    assert name == ("grid_m12i_res7100_ssid7_gridseed1776_Rmax50.0_theta0.5_grid_x_size0.6_grid_y_size0.6_grid_z_size0.6"
                    "_start577_end585_numprior3") and path == str(tmp_path) + "/" + name
    sname, _ = f._grid_cache_name_(580)
    assert sname.startswith("grid_snapshot580_m12i_res7100_ssid7_")
    files = f._snapshot_cache_files_(580)
    assert [os.path.basename(x)[:len("grid_snapshot_x580")] for x in files[:3]] == ["grid_snapshot_x580", "grid_snapshot_y580",
                                                                                     "grid_snapshot_z580"]
    ff, _ = _field(dict(opts, fine_grid=True, grid_fine_x_size_in_kpc=0.02, grid_fine_y_size_in_kpc=0.02,
                        grid_fine_z_size_in_kpc=0.02, grid_fine_resolution=0.0005))
    assert "_fine_grid_x_size0.02_fine_grid_y_size0.02_fine_grid_z_size0.02_fine_grid_resolution0.0005_start577" in \
        ff._grid_cache_name_()[0]
    # a cache written the reference's way (three pickles, protocol 4) is a hit; wrong length or missing file is a miss
    rng = np.random.default_rng(0)
    arrs = [rng.normal(size=4097) for _ in range(3)]
    for path, a in zip(files[:3], arrs):
        pickle.dump(a, open(path, "wb"), protocol=4)
    with pytest.warns(UserWarning, match="no provenance sidecar"):   # the reference writes none: accepted, but flagged
        got = f._load_snapshot_cache_(580, 4097, want_pot=False)
    assert got is not None and all(np.array_equal(g, a) for g, a in zip(got, arrs)) and len(f.cache_unverified) == 1
    assert f._load_snapshot_cache_(580, 4096, want_pot=False) is None
    assert f._load_snapshot_cache_(580, 4097, want_pot=True) is None      # no potential file: recompute
    assert f._load_snapshot_cache_(581, 4097, want_pot=False) is None
    # and what we write is what the reference reads back with pickle.load
    f._dump_snapshot_cache_(581, arrs + [arrs[0] * 2])
    for path, a in zip(f._snapshot_cache_files_(581), arrs + [arrs[0] * 2]):
        assert np.array_equal(pickle.load(open(path, "rb")), a)
    # ... plus a sidecar naming what the file name does not: a later run with another softening kernel rebuilds
    assert f._load_snapshot_cache_(581, 4097, want_pot=True) is not None
    f.softening_kernel = "plummer" if f.softening_kernel != "plummer" else "spline"
    with pytest.warns(UserWarning, match="other settings"):
        assert f._load_snapshot_cache_(581, 4097, want_pot=True) is None
    f.cache_directory = None
    assert f._load_snapshot_cache_(580, 4097, want_pot=False) is None


def test_reference_algorithm_options_are_validated_on_the_host():
    """space_interpolation / basis / integrator (SURVEY §8f rank 5 modes) are checked before any GPU work."""
    import pytest
    from oc_nbody_b200.cluster import cluster_code
    from oc_nbody_b200.gizmo_field import gizmo_field

    class _Snap(object):
        snapshot = {"index": 0, "time": 0.0}
    base = dict(grid_x_size_in_kpc=0.05, grid_y_size_in_kpc=0.05, grid_z_size_in_kpc=0.05, grid_resolution=0.05 / 8)
    f = gizmo_field(dict(base, space_interpolation="rbf", nclose="150", order="5", basis="phs3"), [_Snap()], build=False)
    assert f.space_interpolation == "rbf" and f.nclose == 150 and f.order == 5 and f._rbf_phs == 3

    class phs5(object):  # options.py:178-246 hands over rbf.basis objects; their name is what matters
        pass
    phs5.__name__ = "phs5"
    assert gizmo_field(dict(base, space_interpolation="rbf", basis=phs5), [_Snap()], build=False)._rbf_phs == 5
    assert gizmo_field(dict(base, space_interpolation="rbf", basis="phs8"), [_Snap()], build=False)._rbf_phs == 8
    with pytest.raises(ValueError):          # phs8 needs order >= 4
        gizmo_field(dict(base, space_interpolation="rbf", basis="phs8", order=3), [_Snap()], build=False)
    with pytest.raises(NotImplementedError):  # the shape-parameter bases of options.py:203-246
        gizmo_field(dict(base, space_interpolation="rbf", basis="mq"), [_Snap()], build=False)
    with pytest.raises(ValueError):
        gizmo_field(dict(base, space_interpolation="cubic"), [_Snap()], build=False)
    with pytest.raises(NotImplementedError):
        gizmo_field(dict(base, space_interpolation="rbf", basis="ga"), [_Snap()], build=False)
    with pytest.raises(ValueError):
        cluster_code(np.ones(4), np.zeros((3, 4)), np.zeros((3, 4)), integrator="rk4")


def test_clean_Rmag_follows_the_reference_rule():
    """gizmo_interface.py:297-304: keep rmag < Rmax (strict), every per-particle array of every species."""
    from oc_nbody_b200.gizmo_field import clean_Rmag
    from oc_nbody_b200.synthetic import make_snapshot
    snap = make_snapshot(5000, seed=3)
    before = {k: (snap[k]["position"].copy(), snap[k]["mass"].copy()) for k in ("star", "dark", "gas")}
    n_gas_before = len(snap["gas"]["smooth.length"])
    clean_Rmag(snap, 12.5)
    for k in ("star", "dark", "gas"):
        pos0, m0 = before[k]
        keep = np.sqrt((pos0 * pos0).sum(axis=1)) < 12.5
        assert np.array_equal(snap[k]["position"], pos0[keep]) and np.array_equal(snap[k]["mass"], m0[keep])
        assert len(snap[k]["id"]) == keep.sum()
    assert len(snap["gas"]["smooth.length"]) == len(snap["gas"]["mass"]) < n_gas_before
    assert snap.snapshot["index"] == 577


def test_whole_grid_pickle_written_by_the_reference_loads(tmp_path):
    """gizmo_interface.py:395-398,510: the whole-grid cache is a pickle of the reference's own class.  The committed fixture
    was written by the REAL class (tests/golden/make_golden.py); it loads into our grid with the lattice bookkeeping
    re-derived, and what we write back is the same pickle stream apart from our extra attributes."""
    import pickle
    import pickletools
    from oc_nbody_b200 import cache_compat
    path = os.path.join(ROOT, "tests", "golden", "grid_cache_reference.pickle")
    g = cache_compat.load_grid_pickle(path)
    assert type(g).__module__ == "oc_nbody_b200.grid_cartesian"
    assert g.has_fine_grid and g.coarse_shape == (5, 5, 5) and g.fine_shape == (4, 4, 4)
    assert g.snapshot_acceleration_x.shape == (3, len(g)) and g.snapshot_potential is None
    assert np.array_equal(g.evolved_grid, g.init_grid + np.array([8.0, -0.25, 0.125]))
    rng = np.random.default_rng(1776)
    assert np.array_equal(g.snapshot_acceleration_x, rng.normal(0.0, 1e-2, (3, len(g))))
    # write it back in the reference's format: the class global is the reference's module path, every attribute the
    # reference's own pickle holds is there with equal values, and nothing of our bookkeeping leaks into it
    out = tmp_path / "grid_rewritten"
    cache_compat.dump_grid_pickle(g, str(out))
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(out.read_bytes())]
    assert ("GLOBAL", "oceanic.grid_cartesian grid") in ops
    assert not any(op in ("GLOBAL", "STACK_GLOBAL") and arg and "oc_nbody_b200" in str(arg) for op, arg in ops)

    class _Probe(pickle.Unpickler):  # read both streams as plain attribute dicts, whatever the class
        def find_class(self, module, name):
            if name == "grid":
                return type("grid", (), {})
            return super().find_class(module, name)
    ref_state = _Probe(open(path, "rb")).load().__dict__
    our_state = _Probe(open(str(out), "rb")).load().__dict__
    assert set(ref_state) <= set(our_state) | {"x_n", "y_n", "z_n"}
    for k, v in ref_state.items():
        if k in our_state:
            assert np.array_equal(np.asarray(v), np.asarray(our_state[k])), k
    assert not set(our_state) & {"coarse_shape", "fine_shape", "coarse_keep_index", "coarse_hole_index", "fine_row0"}
    # and our own file reads back
    back = cache_compat.load_grid_pickle(str(out))
    assert np.array_equal(back.snapshot_acceleration_z, g.snapshot_acceleration_z) and back.fine_row0 == g.fine_row0
    # a pickle whose point list is not what its sizes generate is refused
    bad = pickle.loads(pickle.dumps(back))
    bad.init_grid = bad.init_grid + 1e-9
    with open(str(tmp_path / "bad"), "wb") as fh:
        pickle.dump(bad, fh, protocol=4)
    with pytest.raises(ValueError, match="point list"):
        cache_compat.load_grid_pickle(str(tmp_path / "bad"))


def test_whole_grid_pickle_is_read_by_the_real_reference_class(tmp_path):
    """Where the reference is present (the build container): its own plain pickle.load on our file yields ITS class."""
    ref = "/root/reference/grid_cartesian.py"
    if not os.path.exists(ref):
        pytest.skip("/root/reference is not on this box")
    import importlib.util
    import pickle
    import sys
    import types
    from oc_nbody_b200 import cache_compat
    from oc_nbody_b200.grid_cartesian import grid
    g = grid(0.06, 0.06, 0.06, 0.012)
    g.gen_evolved_grid(np.array([8.0, 0.0, 0.0]))
    g.snapshot_acceleration_x = g.snapshot_acceleration_y = g.snapshot_acceleration_z = np.ones((2, len(g)))
    cache_compat.dump_grid_pickle(g, str(tmp_path / "grid_x"))
    spec = importlib.util.spec_from_file_location("oceanic.grid_cartesian", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pkg = types.ModuleType("oceanic")
    pkg.grid_cartesian = mod
    sys.modules["oceanic"], sys.modules["oceanic.grid_cartesian"] = pkg, mod
    try:
        with open(str(tmp_path / "grid_x"), "rb") as fh:
            r = pickle.load(fh)
        assert type(r) is mod.grid and np.array_equal(r.init_grid, g.init_grid)
        r.gen_evolved_grid(np.array([1.0, 2.0, 3.0]))   # the reference's own method works on the loaded object
        assert np.array_equal(r.evolved_grid, g.init_grid + np.array([1.0, 2.0, 3.0]))
    finally:
        del sys.modules["oceanic"], sys.modules["oceanic.grid_cartesian"]


def test_cache_compat_rejects_what_it_cannot_vouch_for(tmp_path):
    """cache_compat's failure modes are loud: a pickle that is not a grid, a grid whose point list is not the one its sizes
    generate, an `interface` file that is not our settings dict."""
    import pickle
    from oc_nbody_b200 import cache_compat
    from oc_nbody_b200.grid_cartesian import grid
    not_a_grid = tmp_path / "plain.pickle"
    pickle.dump({"x": 1}, open(not_a_grid, "wb"), protocol=4)
    with pytest.raises(ValueError, match="does not hold a grid"):
        cache_compat.load_grid_pickle(str(not_a_grid))
    g = grid(0.1, 0.1, 0.1, 0.05)
    g.snapshot_acceleration_x = g.snapshot_acceleration_y = g.snapshot_acceleration_z = np.zeros((2, len(g)))
    good = tmp_path / "grid.pickle"
    cache_compat.dump_grid_pickle(g, str(good))
    back = cache_compat.load_grid_pickle(str(good))
    assert np.array_equal(back.init_grid, g.init_grid) and back.snapshot_acceleration_x.shape == (2, len(g))
    g.init_grid = g.init_grid + 1e-3            # a point list its sizes do not generate
    bad = tmp_path / "tampered.pickle"
    cache_compat.dump_grid_pickle(g, str(bad))
    with pytest.raises(ValueError, match="point list"):
        cache_compat.load_grid_pickle(str(bad))
    (tmp_path / "iface").mkdir()
    pickle.dump([1, 2, 3], open(tmp_path / "iface" / "interface", "wb"), protocol=4)
    with pytest.raises(ValueError, match="settings dict"):
        cache_compat.load_interface(str(tmp_path / "iface"))


def test_iterated_bound_subset_logic_matches_the_oracle():
    """cluster.iterate_bound_subset above a numpy stand-in for the one-pass reduction kernel (same contract as
    ocg_bound_com: frame = weighted mean velocity of all stars, mask over all stars, COM weighted by the weights) against
    oracle.bound_com(iterations=...) on a cluster with a one-sided tail of escapers: the tail drags the one-pass frame, the
    iteration recovers the core's."""
    import torch
    import oracle
    from oc_nbody_b200.cluster import iterate_bound_subset
    from oc_nbody_b200.synthetic import make_plummer_cluster
    pos_pc, vel, mass = make_plummer_cluster(1500, seed=3)
    pos = pos_pc * 1e-3
    tail = np.arange(0, 1500, 4)                       # a quarter of the stars stream away in +x
    vel[0, tail] += 6.0
    pos[0, tail] += 0.004
    G = 4.3009125e-6                                    # kpc (km/s)^2 / Msun
    _, phi = oracle.self_gravity(pos, mass, (0.01e-3) ** 2, G, want_pot=True)
    t_pos, t_vel, t_mass, t_phi = (torch.from_numpy(np.ascontiguousarray(a)) for a in (pos, vel, mass, phi))
    out = torch.empty((1, 8), dtype=torch.float64)
    mask = torch.empty(1500, dtype=torch.uint8)
    calls = []

    def one_pass(weights):
        w = weights.numpy()
        vc = (vel * w).sum(axis=1) / w.sum()
        e = 0.5 * ((vel - vc[:, None]) ** 2).sum(axis=0) + phi
        bd = e < 0.0
        if not bd.any():
            bd[:] = True
        mask.copy_(torch.from_numpy(bd.astype(np.uint8)))
        wb = w * bd
        row = np.concatenate([(pos * wb).sum(axis=1) / wb.sum() if wb.sum() > 0 else (pos * w).sum(axis=1) / w.sum(),
                              [wb.sum(), bd.sum()], vc])
        out.copy_(torch.from_numpy(row[None]))
        calls.append(1)
        return mask
    one = iterate_bound_subset(one_pass, t_pos, t_mass, out, 1)
    com1, m1, _ = oracle.bound_com(pos, vel, mass, phi, iterations=1)
    assert len(calls) == 1 and np.allclose(one[:3], com1, rtol=1e-12, atol=1e-15) and int(one[4]) == m1.sum()
    it = iterate_bound_subset(one_pass, t_pos, t_mass, out, 8)
    com8, m8, _ = oracle.bound_com(pos, vel, mass, phi, iterations=8)
    assert np.array_equal(mask.numpy().astype(bool), m8) and int(it[4]) == m8.sum()
    assert np.allclose(it[:3], com8, rtol=1e-12, atol=1e-15) and np.isclose(it[3], mass[m8].sum())
    assert m8.sum() != m1.sum()                        # the frame matters on this configuration
    assert not m8[tail].any()                          # the escapers are out, whatever the one-pass frame said of them

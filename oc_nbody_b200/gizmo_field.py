"""External-field code: the B200 drop-in for the hot path of the reference's ``gizmo_interface``.

Duck type preserved (what amuse.couple.bridge.Bridge and oc_nbody.py call):
  ``get_gravity_at_point(eps, x, y, z)``   gizmo_interface.py:677-717
  ``get_potential_at_point(eps, x, y, z)`` new per BASELINE.json north_star (no reference body exists)
  ``evolve_model(time, timestep=None)``    gizmo_interface.py:622-638
  ``evolve_grid(pos)``                     gizmo_interface.py:640-642
plus the attributes other reference modules read (``grid``, ``time_in_Myr``, ``chosen_id``,
``chosen_evolved_position`` ...).

What runs where
  * field build  : per snapshot, K1 direct sum on the GPU for every grid point (+ appended origin), frame
                   acceleration subtracted (gizmo_interface.py:512-573), stacked to
                   ``grid.snapshot_acceleration_{x,y,z}[Nsnap, Ngrid+1]`` (gizmo_interface.py:501-506);
  * evolve_model : O(1) — picks the bracketing snapshots and the linear weight; the blend is fused into
                   the interpolation kernel instead of 3*(Ngrid+1) Python-level splev calls through a
                   process pool (gizmo_interface.py:607-620);
  * kick         : K3 trilinear gather at the star positions (replaces the per-star kNN + 3 RBF solves of
                   gizmo_interface.py:661-675,696-704).

Snapshot ingestion (gizmo_analysis/h5py), starting-star selection and the axisymmetric/agama mode are out
of scope: snapshots are handed in as dict-likes with the fields the reference reads.
"""
import os

import numpy as np

from . import _lib
from .grid_cartesian import grid
from .units import G_KPC_KMS_MYR, KMS_TO_PC_PER_MYR, to_value, units

KMS_TO_KPC_PER_MYR = KMS_TO_PC_PER_MYR * 1e-3

_DEFAULTS = dict(
    grid_x_size_in_kpc=0.6, grid_y_size_in_kpc=0.6, grid_z_size_in_kpc=0.6, grid_resolution=0.6 / 16,
    star_softening_in_pc=11.2, dark_softening_in_pc=112.0, star_char_mass=None, dark_char_mass=None,
    softening_kernel="spline", plummer_eps_over_h=1.0 / 2.8, theta=0.5, fine_grid=False, with_potential=True,
    grid_fine_x_size_in_kpc=None, grid_fine_y_size_in_kpc=None, grid_fine_z_size_in_kpc=None, grid_fine_resolution=None,
    # "linear" (BASELINE.json north_star) or "cubic" (the reference's own splrep/splev, gizmo_interface.py:587-620)
    time_interpolation="linear",
    space_interpolation="trilinear",  # "rbf": the reference's own kNN + RBF-PHS interpolant (options nclose/basis/order)
    nclose=150, basis="phs3", order=5,  # options.py:43-45
    # per-snapshot field caches in the reference's own file names and format (gizmo_interface.py:373-391,454-459,
    # 470-495): None = no caching.  The remaining keys only enter the cache file name.
    cache_directory=None, sim_name="synthetic", grid_seed=1776, Rmax=50.0, startnum=None, endnum=None, num_prior=0,
    # source assembly (gizmo_interface.py:515-558): "device" = ocg_assemble_sources, the raw species arrays are uploaded once
    # and cut / softened / concatenated / recentred / rounded on the GPU; "host" = numpy + ocg_field_build_host.
    # clean_Rmax=True applies the reference's ingestion-time cut |pos| < Rmax (_clean_Rmag_, :297-304) during assembly.
    source_assembly="device", clean_Rmax=False,
)


def clean_Rmag(snap, Rmax):
    """Drop every particle farther than ``Rmax`` kpc from the galactic centre, species by species, all per-particle
    arrays alike — ``gizmo_interface._clean_Rmag_`` (gizmo_interface.py:297-304), applied by the reference right after
    reading the snapshots (:254-255).  In place; returns the snapshot."""
    for key in snap.keys():
        part = snap[key]
        keep = np.where(part.prop("host.distance.total") < Rmax)[0]
        n = len(part.prop("host.distance.total"))
        for name in list(part.keys()):
            arr = part[name]
            if hasattr(arr, "shape") and len(getattr(arr, "shape", ())) >= 1 and arr.shape[0] == n:
                part[name] = arr[keep]
    return snap


class gizmo_field(object):
    """Time-varying tidal field on a Cartesian grid that rides with the cluster.

    Parameters
    ----------
    options : dict or an ``options_reader``-like object with ``set_options(obj)`` (options.py:151-153)
    snapshots : sequence of dict-likes with 'star'/'dark'/'gas' species (synthetic.make_snapshot)
    chosen_positions : [Nsnap, 3] kpc, centre of the grid in each snapshot (the tracked star's position,
        gizmo_interface.py:464-468); default: ``grid_center`` for all snapshots
    time_in_Myr : [Nsnap] snapshot times relative to the start (gizmo_interface.py:329-335)
    chosen_id : id of the tracked star, excluded from the sources (gizmo_interface.py:515)
    """

    def __init__(self, options, snapshots, chosen_positions=None, time_in_Myr=None, chosen_id=-1, grid_center=(8.0, 0.0, 0.0),
                 ctx=None, build=True):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        if hasattr(options, "set_options"):
            options.set_options(self)
        else:
            for k, v in dict(options or {}).items():
                setattr(self, k, v)
        if self.softening_kernel not in _lib.KERNELS:
            raise ValueError("softening_kernel must be one of %r" % (sorted(_lib.KERNELS),))
        if self.fine_grid:
            missing = [k for k in ("grid_fine_x_size_in_kpc", "grid_fine_y_size_in_kpc", "grid_fine_z_size_in_kpc",
                                   "grid_fine_resolution") if getattr(self, k, None) is None]
            if missing:
                raise ValueError("fine_grid needs %s (options.py:108-118)" % ", ".join(missing))
        if self.time_interpolation not in ("linear", "cubic"):
            raise ValueError("time_interpolation must be 'linear' or 'cubic'")
        if self.space_interpolation not in ("trilinear", "rbf"):
            raise ValueError("space_interpolation must be 'trilinear' or 'rbf'")
        if self.space_interpolation == "rbf":
            basis = getattr(self.basis, "__name__", None) or str(self.basis)  # a name or an rbf.basis object (options.py:178-246)
            if basis not in ("phs1", "phs2", "phs3", "phs4", "phs5", "phs6", "phs7", "phs8"):
                raise NotImplementedError("space_interpolation='rbf' implements the polyharmonic splines phs1 .. phs8 of "
                                          "options.py:178-202, not the shape-parameter bases (%r)" % (basis,))
            self._rbf_phs = int(basis[3:])
            self.nclose, self.order = int(self.nclose), int(self.order)
            if self.order < self._rbf_phs // 2:
                raise ValueError("basis %s needs order >= %d (the interpolant is not well posed below)" % (basis, self._rbf_phs // 2))
        self.G = G_KPC_KMS_MYR  # kpc^2 km/s /Myr /Msun, the unit of gizmo_interface.py:70
        self._ctx = ctx  # created lazily: host-side logic (source assembly, time bracketing) needs no GPU
        self.snapshots = list(snapshots)
        nsnap = len(self.snapshots)
        if nsnap < 1:
            raise ValueError("need at least one snapshot")
        self.chosen_id = chosen_id
        if chosen_positions is None:
            chosen_positions = np.tile(np.asarray(grid_center, np.float64), (nsnap, 1))
        self.chosen_snapshot_positions = np.asarray(chosen_positions, np.float64).reshape(nsnap, 3)
        if time_in_Myr is None:
            t0 = self.snapshots[0].snapshot["time"]
            time_in_Myr = [1000.0 * (s.snapshot["time"] - t0) for s in self.snapshots]
        self.time_in_Myr = np.asarray(time_in_Myr, np.float64)
        if nsnap > 1 and not np.all(np.diff(self.time_in_Myr) > 0):
            raise ValueError("snapshot times must be strictly increasing")
        self.chosen_evolved_position = self.chosen_snapshot_positions[0].copy()
        self.chosen_evolved_velocity = np.zeros(3)
        self._origin = np.zeros(3)
        self._bracket = (0, min(1, nsnap - 1), 0.0)
        self._blend_cache = None
        self.cache_unverified = []   # cache files loaded without a provenance sidecar (_check_provenance_)
        self._dev = None
        if build:
            self._init_grid_()
            self.evolve_model(0.0 | units.Myr)

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    # ------------------------------------------------------------------ field build (init time) ----
    def _source_arrays_(self, snap):
        """Source assembly (gizmo_interface.py:515-558): star(-chosen) + dark + gas, per-species softening in kpc."""
        star, dark, gas = snap["star"], snap["dark"], snap["gas"]
        ss_key = np.where(star["id"] != self.chosen_id)[0]
        pos = np.concatenate((star.prop("host.distance.principal")[ss_key], dark.prop("host.distance.principal"),
                              gas.prop("host.distance.principal")))
        mass = np.concatenate((star["mass"][ss_key], dark["mass"], gas["mass"]))
        if self.star_char_mass is not None:
            # indexed with ss_key — the evident intent of gizmo_interface.py:531-534
            s_soft = np.power(star["mass"][ss_key] / float(self.star_char_mass), 1.0 / 3.0) / 1000.0
        else:
            s_soft = np.full(len(ss_key), float(self.star_softening_in_pc) / 1000.0)
        if self.dark_char_mass is not None:
            d_soft = np.power(dark["mass"] / float(self.dark_char_mass), 1.0 / 3.0) / 1000.0
        else:
            d_soft = np.full(len(dark["mass"]), float(self.dark_softening_in_pc) / 1000.0)
        g_soft = 2.8 * gas["smooth.length"] / 1000.0
        soft = np.concatenate((s_soft, d_soft, g_soft))
        if self.softening_kernel == "plummer":
            soft = soft * float(self.plummer_eps_over_h)
        return np.float64(pos), np.float64(mass), np.float64(soft)

    def _populate_grid_acceleration_(self, snap, grid, want_pot=False):
        """One snapshot's field on the (shifted) grid, frame acceleration removed
        (gizmo_interface.py:512-573). Returns acc_x, acc_y, acc_z (, pot) as FP64 [Ngrid+1]."""
        if self.source_assembly == "device":
            acc, pot = self._populate_device_(snap, grid, want_pot)
        else:
            if self.clean_Rmax:
                raise ValueError("clean_Rmax needs source_assembly='device' (or call clean_Rmag(snap, Rmax) at ingestion)")
            r, m, soft = self._source_arrays_(snap)
            out = self.ctx.field_build_host(r, m, soft, grid.evolved_grid, grid.ss_evolved_position, grid.origin_row,
                                            _lib.KERNELS[self.softening_kernel], self.G, want_pot=want_pot)
            acc, pot = out if want_pot else (out, None)
        if want_pot:
            return acc[0], acc[1], acc[2], pot
        return acc[0], acc[1], acc[2]

    def _species_rules_(self):
        """(species, softening rule, parameter) in the reference's concatenation order (gizmo_interface.py:518-549)."""
        star = (("mass_cube_root", float(self.star_char_mass)) if self.star_char_mass is not None
                else ("constant", float(self.star_softening_in_pc) / 1000.0))
        dark = (("mass_cube_root", float(self.dark_char_mass)) if self.dark_char_mass is not None
                else ("constant", float(self.dark_softening_in_pc) / 1000.0))
        return (("star",) + star, ("dark",) + dark, ("gas", "gas_smoothing", 0.0))

    def _assemble_device_(self, snap, center):
        """Source records on the device: FP32 (x, y, z, m) recentred on `center` + softening, star | dark | gas
        (ocg_assemble_sources: tracked star dropped, optional Rmax cut, per-species softening).  Returns (xyzm, soft)."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        n_max = sum(len(snap[sp]["mass"]) for sp in ("star", "dark", "gas"))
        xyzm = torch.empty((n_max, 4), dtype=torch.float32, device=dev)
        soft = torch.empty(n_max, dtype=torch.float32, device=dev)
        scale = float(self.plummer_eps_over_h) if self.softening_kernel == "plummer" else 1.0
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
        off = 0
        for sp, rule, param in self._species_rules_():
            part = snap[sp]
            ids = up(part["id"], np.int64) if sp == "star" else None
            hsml = up(part["smooth.length"], np.float64) if sp == "gas" else None
            off += self.ctx.assemble_sources(up(part.prop("host.distance.principal"), np.float64), up(part["mass"], np.float64), ids,
                                             hsml, self.chosen_id, float(self.Rmax) if self.clean_Rmax else 0.0, rule, param, scale,
                                             center, xyzm, soft, off)
        return xyzm[:off], soft[:off]

    def _populate_device_(self, snap, grid, want_pot):
        import torch
        dev = torch.device("cuda", self.ctx.device)
        center = np.asarray(grid.ss_evolved_position, np.float64)
        xyzm, soft = self._assemble_device_(snap, center)
        n_tgt = len(grid)
        tgt = torch.empty((n_tgt, 4), dtype=torch.float32, device=dev)
        self.ctx.recentre_f64(torch.from_numpy(np.ascontiguousarray(grid.evolved_grid, dtype=np.float64)).to(dev), None, center, tgt)
        acc = torch.empty((3, n_tgt), dtype=torch.float64, device=dev)
        pot = torch.empty(n_tgt, dtype=torch.float64, device=dev) if want_pot else None
        self.ctx.field_direct(xyzm, soft, tgt, _lib.KERNELS[self.softening_kernel], self.G, acc, pot)
        self.ctx.frame_subtract(acc, grid.origin_row)
        return acc.cpu().numpy(), (pot.cpu().numpy() if want_pot else None)

    def _make_grid_(self):
        """grid(...) [+ add_fine_grid(...)] exactly as gizmo_interface.py:412-419."""
        g = grid(self.grid_x_size_in_kpc, self.grid_y_size_in_kpc, self.grid_z_size_in_kpc, self.grid_resolution)
        if self.fine_grid:
            g.add_fine_grid(self.grid_fine_x_size_in_kpc, self.grid_fine_y_size_in_kpc, self.grid_fine_z_size_in_kpc,
                            self.grid_fine_resolution)
        return g

    # ---- the reference's per-snapshot field caches (gizmo_interface.py:373-391,454-459,470-495) ----
    def _grid_cache_name_(self, snapshot_index=None):
        """Same file name, character for character, as gizmo_interface.py:373-391."""
        if snapshot_index is not None:
            cache_name = 'grid_snapshot' + str(snapshot_index) + '_'
        else:
            cache_name = 'grid_'
        cache_name += self.sim_name
        cache_name += '_ssid' + str(self.chosen_id)
        cache_name += '_gridseed' + str(self.grid_seed) + '_Rmax' + str(self.Rmax)
        cache_name += '_theta' + str(self.theta) + '_grid_x_size' + str(self.grid_x_size_in_kpc)
        cache_name += '_grid_y_size' + str(self.grid_y_size_in_kpc)
        cache_name += '_grid_z_size' + str(self.grid_z_size_in_kpc)
        if self.fine_grid:
            cache_name += '_fine_grid_x_size' + str(self.grid_fine_x_size_in_kpc)
            cache_name += '_fine_grid_y_size' + str(self.grid_fine_y_size_in_kpc)
            cache_name += '_fine_grid_z_size' + str(self.grid_fine_z_size_in_kpc)
            cache_name += '_fine_grid_resolution' + str(self.grid_fine_resolution)
        cache_name += '_start' + str(self.startnum)
        cache_name += '_end' + str(self.endnum) + '_numprior' + str(self.num_prior)
        return cache_name, str(self.cache_directory) + '/' + cache_name

    def _snapshot_cache_files_(self, snapshot_index):
        """x, y, z files as the reference derives them (gizmo_interface.py:441-446: 'snapshot' -> 'snapshot_x' ...)
        + our potential file.  The reference runs str.replace over the whole PATH, so a cache directory whose name
        contains 'snapshot' would be rewritten too; here only the file name is touched."""
        name, _ = self._grid_cache_name_(snapshot_index)
        return tuple(str(self.cache_directory) + '/' + name.replace('snapshot', 'snapshot_' + c) for c in ('x', 'y', 'z', 'pot'))

    # ---- provenance sidecars: the reference's cache names do not carry everything that changes the cached numbers ----
    def _cache_provenance_(self):
        """What produced a cached field beyond what the reference's file name encodes (round-1 advice): the softening
        kernel and lengths and the summation method.  Written as <cache file name>.provenance.json beside the cache."""
        def num(v):
            return None if v is None else float(v)
        return {"producer": "oc_nbody_b200 direct sum (FP32 pair arithmetic, FP64 accumulation)",
                "softening_kernel": str(self.softening_kernel), "plummer_eps_over_h": num(self.plummer_eps_over_h),
                "star_softening_in_pc": num(self.star_softening_in_pc), "dark_softening_in_pc": num(self.dark_softening_in_pc),
                "star_char_mass": num(self.star_char_mass), "dark_char_mass": num(self.dark_char_mass)}

    def _write_provenance_(self, cache_path):
        import json
        with open(cache_path + ".provenance.json", "w") as fh:
            json.dump(self._cache_provenance_(), fh, indent=1, sort_keys=True)

    def _check_provenance_(self, cache_path):
        """True: the sidecar matches these settings.  False: it does not — the cache is ignored and rebuilt.  None: no
        sidecar, i.e. a cache the reference wrote (or an older run): accepted as the reference accepts it, but recorded in
        ``cache_unverified`` with a warning — its numbers come from the theta = 0.5 tree (~1e-3 relative) with the
        reference's softening, whatever this run's options say."""
        import json
        import warnings
        try:
            with open(cache_path + ".provenance.json") as fh:
                have = json.load(fh)
        except (OSError, ValueError):
            self.cache_unverified.append(cache_path)
            if len(self.cache_unverified) == 1:
                warnings.warn("field cache %s has no provenance sidecar (written by the reference?): its values are used as they "
                              "are — tree-code accuracy, the writer's softening settings" % cache_path)
            return None
        want = self._cache_provenance_()
        if have != want:
            diff = sorted(k for k in set(have) | set(want) if have.get(k) != want.get(k))
            warnings.warn("field cache %s was written with other settings (%s): rebuilding" % (cache_path, ", ".join(diff)))
            return False
        return True

    def _load_snapshot_cache_(self, snapshot_index, n_points, want_pot):
        """Pickled FP64 [Ngrid+1] arrays the reference (or this code) wrote; None on any miss, shape mismatch, or a
        provenance sidecar that names other softening settings."""
        import pickle
        if self.cache_directory is None:
            return None
        files = self._snapshot_cache_files_(snapshot_index)
        if not all(os.path.exists(f) for f in files[:4 if want_pot else 3]):
            return None
        if self._check_provenance_(self._grid_cache_name_(snapshot_index)[1]) is False:
            return None
        try:
            out = []
            for f in files[:4 if want_pot else 3]:
                with open(f, 'rb') as fh:
                    a = np.asarray(pickle.load(fh), np.float64)
                if a.shape != (n_points,):
                    return None
                out.append(a)
            return tuple(out)
        except (OSError, pickle.UnpicklingError, EOFError, ValueError):
            return None

    def _dump_snapshot_cache_(self, snapshot_index, arrays):
        """pickle protocol 4 of each array into its own file (gizmo_interface.py:454-459,490-495)."""
        import pickle
        if self.cache_directory is None:
            return
        os.makedirs(str(self.cache_directory), exist_ok=True)
        for f, a in zip(self._snapshot_cache_files_(snapshot_index), arrays):
            with open(f, 'wb') as fh:
                pickle.dump(np.asarray(a, np.float64), fh, protocol=4)
        self._write_provenance_(self._grid_cache_name_(snapshot_index)[1])

    def _init_grid_(self):
        """Per-snapshot grid loop (gizmo_interface.py:393-510): the whole-grid cache first (:395-398), else per snapshot load
        the cached field or build it on the GPU; both caches are written in the reference's own formats (cache_compat)."""
        from . import cache_compat
        self.grid = self._make_grid_()
        if self.startnum is None:
            self.startnum = self.snapshots[0].snapshot["index"]
        if self.endnum is None:
            self.endnum = self.snapshots[-1].snapshot["index"]
        self.cache_hits = 0
        self.grid_cache_hit = False
        self.cache_unverified = []
        grid_cache_file = self._grid_cache_name_()[1] if self.cache_directory is not None else None
        if grid_cache_file is not None and os.path.exists(grid_cache_file) and self._check_provenance_(grid_cache_file) is not False:
            try:
                cached = cache_compat.load_grid_pickle(grid_cache_file)
                shape = (len(self.snapshots), len(self.grid))
                pot_ok = (not self.with_potential) or (cached.snapshot_potential is not None and cached.snapshot_potential.shape == shape)
                if np.array_equal(cached.init_grid, self.grid.init_grid) and cached.snapshot_acceleration_x.shape == shape and pot_ok:
                    for k in ("snapshot_acceleration_x", "snapshot_acceleration_y", "snapshot_acceleration_z"):
                        setattr(self.grid, k, getattr(cached, k))
                    self.grid.snapshot_potential = cached.snapshot_potential if self.with_potential else None
                    self.grid.gen_evolved_grid(self._origin)
                    self.grid_cache_hit = True
                    self._upload_planes_()
                    return
            except (OSError, ValueError, AttributeError, EOFError, ImportError, IndexError, KeyError, TypeError):
                pass  # "couldnt find cached grid" (gizmo_interface.py:399-400): build it
        ax, ay, az, ph = [], [], [], []
        for i, snap in enumerate(self.snapshots):
            self.grid.gen_evolved_grid(self.chosen_snapshot_positions[i])
            res = self._load_snapshot_cache_(snap.snapshot["index"], len(self.grid), self.with_potential)
            if res is not None:
                self.cache_hits += 1
            else:
                res = self._populate_grid_acceleration_(snap, self.grid, want_pot=self.with_potential)
                self._dump_snapshot_cache_(snap.snapshot["index"], res)
            ax.append(res[0]), ay.append(res[1]), az.append(res[2])
            if self.with_potential:
                ph.append(res[3])
        self.grid.snapshot_acceleration_x = np.array(ax)
        self.grid.snapshot_acceleration_y = np.array(ay)
        self.grid.snapshot_acceleration_z = np.array(az)
        self.grid.snapshot_potential = np.array(ph) if self.with_potential else None
        self.grid.gen_evolved_grid(self._origin)
        if grid_cache_file is not None:
            cache_compat.dump_grid_pickle(self.grid, grid_cache_file)   # gizmo_interface.py:510
            self._write_provenance_(grid_cache_file)
        self._upload_planes_()

    def set_snapshot_fields(self, acc_x, acc_y, acc_z, pot=None):
        """Install pre-computed fields ([Nsnap, Ngrid+1] each, e.g. loaded from the reference's caches)."""
        if not hasattr(self, "grid"):
            self.grid = self._make_grid_()
        self.grid.snapshot_acceleration_x = np.asarray(acc_x, np.float64)
        self.grid.snapshot_acceleration_y = np.asarray(acc_y, np.float64)
        self.grid.snapshot_acceleration_z = np.asarray(acc_z, np.float64)
        self.grid.snapshot_potential = None if pot is None else np.asarray(pot, np.float64)
        self.with_potential = pot is not None
        self._upload_planes_()

    def _layout_planes_(self, planes):
        """FP64 [P, 4, Npoints] (ax, ay, az, phi per point of grid.init_grid) -> float4 node records in HBM.

        Single lattice: one record array [P, Npoints, 4] in point order (= lattice order + origin row).
        Nested grid (grid_cartesian.py:34-53,71-91): two arrays, ``coarse`` [P, ncoarse+1, 4] on the FULL coarse
        lattice and ``fine`` [P, nfine+1, 4].  Kept coarse points are scattered to their lattice slots; the coarse
        points the reference dropped (strictly inside the fine box) are filled by interpolating the fine lattice at
        their positions (K3 itself), so that coarse cells straddling the fine box have all eight corners."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        g = self.grid
        P, _, npts = planes.shape
        if npts != len(g):
            raise ValueError("field arrays have %d points, grid has %d" % (npts, len(g)))

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        if not g.has_fine_grid:
            rec = torch.empty((P, npts, 4), dtype=torch.float32, device=dev)
            for i in range(P):
                self.ctx.pack_planes(up(planes[i, :3]), up(planes[i, 3]), rec[i])
            return dict(coarse=rec, fine=None)
        nc, nf, f0 = g.n_lattice, int(np.prod(g.fine_shape)), g.fine_row0
        coarse = torch.zeros((P, nc + 1, 4), dtype=torch.float32, device=dev)
        fine = torch.empty((P, nf + 1, 4), dtype=torch.float32, device=dev)
        keep_idx = up(np.concatenate([g.coarse_keep_index, [nc]]).astype(np.int64))  # + the origin row
        hole_idx = up(g.coarse_hole_index)
        hole_pts = g.coarse_hole_points()
        hx, hy, hz = (up(hole_pts[:, d]) for d in range(3))
        fnodes = [up(a) for a in g.fine_nodes]
        zero_origin = torch.zeros((1, 3), dtype=torch.float64, device=dev)
        n_hole = hole_pts.shape[0]
        h_acc = torch.empty((3, n_hole), dtype=torch.float64, device=dev)
        h_pot = torch.empty(n_hole, dtype=torch.float64, device=dev)
        for i in range(P):
            self.ctx.pack_planes(up(planes[i, :3, f0:]), up(planes[i, 3, f0:]), fine[i])
            rows = np.concatenate([planes[i, :, :f0], planes[i, :, -1:]], axis=1)
            self.ctx.pack_planes_indexed(up(rows[:3]), up(rows[3]), keep_idx, coarse[i])
            if n_hole:
                self.ctx.grid_interp(g.fine_shape, fnodes, zero_origin, fine[i], None, 0.0, hx, hy, hz, None, h_acc, h_pot)
                self.ctx.pack_planes_indexed(h_acc, h_pot, hole_idx, coarse[i])
        return dict(coarse=coarse, fine=fine)

    def _planes_(self):
        g = self.grid
        stacks = [g.snapshot_acceleration_x, g.snapshot_acceleration_y, g.snapshot_acceleration_z]
        stacks.append(g.snapshot_potential if g.snapshot_potential is not None else np.zeros_like(stacks[0]))
        return np.stack([np.asarray(a, np.float64) for a in stacks], axis=1)  # [Nsnap, 4, Npoints]

    def _upload_planes_(self):
        """FP64 [Nsnap, Npoints] x3 (+ potential) -> float4 node records per snapshot in HBM (K2 pack)."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        g = self.grid
        planes = self._planes_()
        lay = self._layout_planes_(planes)
        self._dev = dict(rec=lay["coarse"], rec_fine=lay["fine"],
                         nodes=[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in g.nodes],
                         fine_nodes=[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in g.fine_nodes] if g.has_fine_grid else None,
                         origin=torch.zeros((1, 3), dtype=torch.float64, device=dev), device=dev)
        if self.time_interpolation == "cubic":
            # one vectorised spline fit over all grid points (= splrep per point, gizmo_interface.py:591-597);
            # the coefficient planes replace the snapshot planes as what K3 gathers from
            from . import time_spline
            self._knots, coef = time_spline.fit(self.time_in_Myr, planes)  # [Nsnap, 4, Npoints]
            self._coef = coef
            clay = self._layout_planes_(coef)
            self._dev["coef_rec"], self._dev["coef_rec_fine"] = clay["coarse"], clay["fine"]
        self._set_origin_(self._origin)

    # ---------------------------------------------------------------------- per-step state ----
    def _set_origin_(self, pos):
        import torch
        self._origin = np.asarray(pos, np.float64).reshape(3).copy()
        if self._dev is not None:
            self._dev["origin"].copy_(torch.from_numpy(self._origin.reshape(1, 3)))
        self._blend_cache = None

    def evolve_model(self, time, timestep=None):
        """Advance the field to `time`: choose the bracketing snapshots and the linear weight.
        (The reference evaluates 3*(Ngrid+1) splines here, gizmo_interface.py:607-638.)  The grid origin set
        by evolve_grid is kept — the reference resets the coordinates but not its KD-tree."""
        t = float(to_value(time, units.Myr))
        self.time = t
        times = self.time_in_Myr
        if len(times) == 1:
            self._bracket = (0, 0, 0.0)
        else:
            i = int(np.searchsorted(times, t, side="right")) - 1
            i = min(max(i, 0), len(times) - 2)
            w = (t - times[i]) / (times[i + 1] - times[i])
            self._bracket = (i, i + 1, float(min(max(w, 0.0), 1.0)))
        if len(times) > 1:
            a, b, w = self._bracket
            p = self.chosen_snapshot_positions
            self.chosen_evolved_position = (1.0 - w) * p[a] + w * p[b]
        if self.time_interpolation == "cubic" and getattr(self, "_knots", None) is not None:
            from . import time_spline
            self._spline_first, self._spline_w = time_spline.basis(self._knots, t)
        self._blend_cache = None

    def evolve_grid(self, pos):
        """Re-origin the grid at the cluster's bound centre of mass (gizmo_interface.py:640-642)."""
        self.grid.gen_evolved_grid(np.asarray(pos, np.float64))
        self._set_origin_(pos)

    # materialised blend, for code that reads grid.evolved_acceleration_* (gizmo_interface.py:618-620)
    def _blend_(self):
        if self._blend_cache is None and self.time_interpolation == "cubic":
            f, w = self._spline_first, self._spline_w
            v = np.tensordot(w, self._coef[f:f + 4], axes=1)  # FP64 on the host: [4, n_node]
            self._blend_cache = (v[:3], v[3])
        if self._blend_cache is None and self.grid.has_fine_grid:
            # records are in lattice order for the two-level grid; blend the FP64 point-list stacks on the host
            a, b, w = self._bracket
            pl = self._planes_()
            v = (1.0 - w) * pl[a] + w * pl[b]
            self._blend_cache = (v[:3], v[3])
        if self._blend_cache is None:
            import torch
            a, b, w = self._bracket
            rec = self._dev["rec"]
            n = rec.shape[1]
            acc = torch.empty((3, n), dtype=torch.float64, device=self._dev["device"])
            pot = torch.empty(n, dtype=torch.float64, device=self._dev["device"])
            self.ctx.grid_time_blend(rec[a], rec[b] if b != a else None, w, acc, pot)
            self._blend_cache = (acc.cpu().numpy(), pot.cpu().numpy())
        return self._blend_cache

    @property
    def evolved_acceleration(self):
        return self._blend_()[0]

    @property
    def evolved_potential(self):
        return self._blend_()[1]

    # ---------------------------------------------------------------------------- the kick ----
    def _time_planes_(self):
        """(coarse planes, fine planes or None, weights) of the current model time."""
        d = self._dev
        if self.time_interpolation == "cubic":
            f = self._spline_first
            fine = None if d["coef_rec_fine"] is None else [d["coef_rec_fine"][f + j] for j in range(4)]
            return [d["coef_rec"][f + j] for j in range(4)], fine, list(self._spline_w)
        a, b, w = self._bracket
        if b == a:
            return [d["rec"][a]], None if d["rec_fine"] is None else [d["rec_fine"][a]], [1.0]
        # FP32 weights exactly as ocg_grid_interp forms them: wb = fl32(w), wa = fl32(1 - wb)
        wb = np.float32(w)
        wa = np.float32(np.float32(1.0) - wb)
        fine = None if d["rec_fine"] is None else [d["rec_fine"][a], d["rec_fine"][b]]
        return [d["rec"][a], d["rec"][b]], fine, [float(wa), float(wb)]

    def _rbf_field_(self):
        """FP64 [4, n_node] device array of the time-evaluated grid fields (grid.evolved_acceleration_x/y/z and the
        potential, gizmo_interface.py:618-620) that the RBF interpolant reads; refreshed when the model time changes.
        Single lattice, linear in time: blended on the device (K2) with no host round trip.  Cubic in time or nested
        grid: the host-side blend, uploaded (nested: over the reference's whole point list, see _interp_rbf_)."""
        import torch
        d, g = self._dev, self.grid
        if self.time_interpolation == "linear" and not g.has_fine_grid:
            a, b, w = self._bracket
            key = (a, b, w, d["rec"].data_ptr())
            if getattr(self, "_rbf_cache", None) is None or self._rbf_cache[0] != key:
                n = d["rec"].shape[1]
                f = torch.empty((4, n), dtype=torch.float64, device=d["device"])
                self.ctx.grid_time_blend(d["rec"][a], d["rec"][b] if b != a else None, w, f[:3], f[3])
                self._rbf_cache = (key, f)
            return self._rbf_cache[1]
        blend = self._blend_()
        if getattr(self, "_rbf_cache", None) is None or self._rbf_cache[0] is not blend:
            f = np.concatenate([blend[0], blend[1][None]], axis=0)   # nested grid: the WHOLE point list (both levels are searched)
            self._rbf_cache = (blend, torch.from_numpy(np.ascontiguousarray(f)).to(d["device"]))
        return self._rbf_cache[1]

    def _interp_rbf_(self, sx, sy, sz, want_pot, want_tensor):
        """K7: kNN(nclose) + RBF-PHS on the device; the last status vector is kept in self.rbf_status (status & 0xff
        == 0: good).  On the nested grid the search runs over BOTH levels of the reference's point list (kept coarse points,
        fine lattice, origin row): a star within a few fine cells of the fine-box surface gets the mixed-level stencil
        cKDTree.query would give it (gizmo_interface.py:661-675 over grid_cartesian.py:71-91)."""
        import torch
        d, g = self._dev, self.grid
        n = sx.shape[0]
        f = self._rbf_field_()
        out = torch.empty((4, n), dtype=torch.float64, device=d["device"])
        tensor = torch.empty((3, 4, n), dtype=torch.float64, device=d["device"]) if want_tensor else None
        self.rbf_status = torch.empty(n, dtype=torch.int32, device=d["device"])
        shape, nodes_h, nodes_d = (g.fine_shape, g.fine_nodes, d["fine_nodes"]) if g.has_fine_grid else (g.shape, g.nodes, d["nodes"])
        dup = all(len(a) % 2 == 1 and a[len(a) // 2] == 0.0 for a in nodes_h)  # the origin row duplicates a lattice node
        if g.has_fine_grid:
            if "coarse_row" not in d:
                row = np.full(int(np.prod(g.coarse_shape)), -1, np.int32)
                row[g.coarse_keep_index] = np.arange(len(g.coarse_keep_index), dtype=np.int32)
                d["coarse_row"] = torch.from_numpy(row).to(d["device"])
            self.ctx.grid_interp_rbf_nested(g.coarse_shape, d["nodes"], g.fine_shape, d["fine_nodes"], d["origin"], d["coarse_row"],
                                            g.fine_row0, f, sx, sy, sz, None, out, nclose=self.nclose, order=self.order,
                                            phs=self._rbf_phs, include_origin=not dup, tensor_out=tensor, status_out=self.rbf_status)
        else:
            self.ctx.grid_interp_rbf(shape, nodes_d, d["origin"], f, sx, sy, sz, None, out, nclose=self.nclose,
                                     order=self.order, phs=self._rbf_phs, include_origin=not dup, tensor_out=tensor,
                                     status_out=self.rbf_status)
        # stars without a good stencil (status & 0xff != 0: too far outside the grid, clipped / ill-conditioned stencil) are
        # counted on the device; rbf_bad_count reads the count back on demand and get_gravity_at_point warns about it
        self._rbf_bad = (self.rbf_status & 0xff).ne(0).sum()
        acc, pot = out[:3], (out[3] if want_pot else None)
        if want_tensor:
            # [3 (d/dx_i), 3 (a_j), n] -> [9, n] with row 3*i + j, as K3's tensor output
            return acc, pot, tensor[:, :3, :].reshape(9, n)
        return acc, pot

    @property
    def rbf_bad_count(self):
        """Stars of the last RBF evaluation whose stencil was truncated, ill-conditioned or singular (their values are not
        to be trusted; NaN when no stencil existed at all)."""
        b = getattr(self, "_rbf_bad", None)
        return 0 if b is None else int(b.item())

    def _interp_device_(self, sx, sy, sz, want_pot, want_tensor=False):
        """K3 on device tensors (FP64 kpc). Returns acc [3,n], pot [n] or None (, tensor [9,n]) device tensors."""
        import torch
        if self.space_interpolation == "rbf":
            return self._interp_rbf_(sx, sy, sz, want_pot, want_tensor)
        d = self._dev
        n = sx.shape[0]
        acc = torch.empty((3, n), dtype=torch.float64, device=d["device"])
        pot = torch.empty(n, dtype=torch.float64, device=d["device"]) if want_pot else None
        g = self.grid
        if g.has_fine_grid or want_tensor:
            recs, recs_fine, w = self._time_planes_()
            tensor = torch.empty((9, n), dtype=torch.float64, device=d["device"]) if want_tensor else None
            self.ctx.grid_interp_nested(g.shape, d["nodes"], d["origin"], recs, w, sx, sy, sz, None, acc, pot,
                                        fine_n=g.fine_shape, fine_nodes=d["fine_nodes"], recs_fine=recs_fine,
                                        tensor_out=tensor)
            return (acc, pot, tensor) if want_tensor else (acc, pot)
        if self.time_interpolation == "cubic":
            f = self._spline_first
            self.ctx.grid_interp_multi(g.shape, d["nodes"], d["origin"], [d["coef_rec"][f + j] for j in range(4)],
                                       self._spline_w, sx, sy, sz, None, acc, pot)
            return acc, pot
        a, b, w = self._bracket
        rec = d["rec"]
        self.ctx.grid_interp(g.shape, d["nodes"], d["origin"], rec[a], rec[b] if b != a else None, w, sx, sy, sz,
                             None, acc, pot)
        return acc, pot

    def _interp_host_(self, x, y, z, want_pot):
        import torch
        scalar = not hasattr(x, "__iter__") and np.ndim(x) == 0
        xs = [torch.from_numpy(np.atleast_1d(np.asarray(v, np.float64)).copy()).to(self._dev["device"]) for v in (x, y, z)]
        if not (xs[0].shape == xs[1].shape == xs[2].shape):
            raise ValueError("x, y, z must have the same length")
        acc, pot = self._interp_device_(xs[0], xs[1], xs[2], want_pot)
        if self.space_interpolation == "rbf" and self.rbf_bad_count:
            import warnings
            warnings.warn("%d of %d points have no trustworthy RBF stencil (rbf_status & 0xff != 0)" % (self.rbf_bad_count, xs[0].shape[0]))
        acc = acc.cpu().numpy()
        pot = pot.cpu().numpy() if want_pot else None
        return scalar, acc, pot

    def get_gravity_at_point(self, eps, xlist, ylist, zlist):
        """Acceleration at the given points (kpc) in km/s/Myr; `eps` is accepted and ignored, as in the reference.
        Scalar and 1-D forms (gizmo_interface.py:696-717)."""
        x, y, z = (to_value(v, units.kpc) for v in (xlist, ylist, zlist))
        scalar, acc, _ = self._interp_host_(x, y, z, False)
        u = units.kms / units.Myr
        if scalar:
            return float(acc[0, 0]) | u, float(acc[1, 0]) | u, float(acc[2, 0]) | u
        return acc[0] | u, acc[1] | u, acc[2] | u

    def get_potential_at_point(self, eps, xlist, ylist, zlist):
        """Potential of the snapshot particles at the given points in (km/s)^2 (AMUSE's convention for
        get_potential_at_point); the uniform frame term removed from the accelerations is not integrated in."""
        if not self.with_potential:
            raise RuntimeError("field was built with with_potential=False")
        x, y, z = (to_value(v, units.kpc) for v in (xlist, ylist, zlist))
        scalar, _, pot = self._interp_host_(x, y, z, True)
        pot = pot / KMS_TO_KPC_PER_MYR  # kpc km/s /Myr -> (km/s)^2
        u = units.kms ** 2
        return (float(pot[0]) | u) if scalar else (pot | u)

    def get_tidal_tensor_at_point(self, eps, xlist, ylist, zlist):
        """Tidal tensor T[i][j] = d a_j / d x_i in km/s/Myr/kpc (gizmo_interface.py:719-756): the analytic gradient
        of the trilinear interpolant in place of nine RBF derivative evaluations per star.  Scalar form returns a
        3x3 quantity, 1-D form an [n, 3, 3] one."""
        import torch
        x, y, z = (to_value(v, units.kpc) for v in (xlist, ylist, zlist))
        scalar = not hasattr(x, "__iter__") and np.ndim(x) == 0
        xs = [torch.from_numpy(np.atleast_1d(np.asarray(v, np.float64)).copy()).to(self._dev["device"]) for v in (x, y, z)]
        _, _, tensor = self._interp_device_(xs[0], xs[1], xs[2], False, want_tensor=True)
        T = tensor.cpu().numpy().T.reshape(-1, 3, 3)
        u = units.kms / units.Myr / units.kpc
        return (T[0] | u) if scalar else (T | u)

    def kick_device(self, pos_kpc, vel_kms, dt_myr, planes=None, w_slot=None):
        """Fused BRIDGE half-kick on device state (FP64 [3,n] tensors): v += dt * a_tidal(x). No host copies.
        With `planes` = (coarse planes, fine planes or None) and `w_slot` the time-blend weights are read from the
        constant-memory slot (ocg_set_interp_weight_slots) — the form a captured CUDA graph replays."""
        if w_slot is None or self.space_interpolation == "rbf":
            acc = self._interp_device_(pos_kpc[0], pos_kpc[1], pos_kpc[2], False)[0]
        else:
            import torch
            d, g = self._dev, self.grid
            acc = torch.empty((3, pos_kpc.shape[1]), dtype=torch.float64, device=d["device"])
            self.ctx.grid_interp_slot(g.shape, d["nodes"], d["origin"], planes[0], w_slot, pos_kpc[0], pos_kpc[1], pos_kpc[2],
                                      None, acc, None, fine_n=g.fine_shape, fine_nodes=d["fine_nodes"], recs_fine=planes[1])
        self.ctx.kick(vel_kms, acc, dt_myr)

    def stop(self):
        pass


# the name the reference's driver imports (oc_nbody.py:11)
gizmo_interface = gizmo_field

"""The reference's remaining cache / wire formats, readable and writable without the reference:

* the whole-``grid`` pickle of gizmo_interface.py:395-398,510 — ``pickle.dump(self.grid, ..., protocol=4)`` of an
  ``oceanic.grid_cartesian.grid`` instance holding the lattice, the point list and the stacked per-snapshot fields;
* the piecemeal interface dump of oceanic_io.py:72-145 (``dump_interface`` / ``load_interface``): ``evolved_grid``,
  ``init_grid``, ``snapshot_acceleration_{x,y,z}.npz`` (pickles in spite of the suffix), ``grid_acc{x,y,z}_interpolators``
  and ``interface``.

A pickle names its class by module path.  The reference's is ``oceanic.grid_cartesian.grid`` (gizmo_interface.py:8); ours
is ``oc_nbody_b200.grid_cartesian.grid`` with the same attribute names (the preserved plugin surface).  Writing therefore
emits the reference's module path for the class global, with only the attributes the reference's class has, so that the
reference's plain ``pickle.load`` rebuilds ITS class; reading maps that global (also the bare ``grid_cartesian.grid``) onto
our class and re-derives our extra lattice bookkeeping from the stored sizes.

The ``interface`` file of ``dump_interface`` is a dill of the reference's own ``gizmo_interface`` object and cannot exist
without that class; here it holds a plain dict of the field code's settings (loadable by ``dill.load`` / ``pickle.load``
as data), and ``load_interface`` rebuilds a ``gizmo_field`` from it.
"""
import io
import os
import pickle

import numpy as np

from .grid_cartesian import grid as _grid

REFERENCE_GRID_MODULE = "oceanic.grid_cartesian"
_REF_ATTRS = ("x_size_in_kpc", "y_size_in_kpc", "z_size_in_kpc", "resolution", "x_n", "y_n", "z_n", "x_grid", "y_grid", "z_grid",
              "fine_x_size_in_kpc", "fine_y_size_in_kpc", "fine_z_size_in_kpc", "fine_resolution", "x_fine_grid", "y_fine_grid",
              "z_fine_grid", "init_grid", "evolved_grid", "ss_evolved_position", "snapshot_acceleration_x",
              "snapshot_acceleration_y", "snapshot_acceleration_z", "snapshot_potential", "evolved_acceleration_x",
              "evolved_acceleration_y", "evolved_acceleration_z")


class _RefGridStandIn(object):
    """Pickled in place of our grid: its class global is written as <module>.grid."""


class _RefPickler(pickle._Pickler):
    def __init__(self, fh, module):
        super().__init__(fh, protocol=4)
        self._module = module

    def save_global(self, obj, name=None):
        if obj is _RefGridStandIn:
            self.write(pickle.GLOBAL + self._module.encode("ascii") + b"\n" + b"grid\n")
            self.memoize(obj)
            return
        super().save_global(obj, name)


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == "grid" and module in (REFERENCE_GRID_MODULE, "grid_cartesian", "oc_nbody_b200.grid_cartesian"):
            return _grid
        return super().find_class(module, name)


def dump_grid_pickle(grid_obj, path, module=REFERENCE_GRID_MODULE):
    """Write ``grid_obj`` as the reference writes its whole-grid cache (gizmo_interface.py:510): protocol 4, class global
    ``<module>.grid``, the reference class's attributes only."""
    stand_in = _RefGridStandIn()
    for k in _REF_ATTRS:
        if hasattr(grid_obj, k) and getattr(grid_obj, k) is not None:
            stand_in.__dict__[k] = getattr(grid_obj, k)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fh:
        _RefPickler(fh, module).dump(stand_in)


def _rederive_(g):
    """Our grid class carries lattice bookkeeping the reference's does not (shapes, kept / dropped coarse rows, first fine
    row).  Re-run the construction from the stored sizes, check the point list is the pickled one, adopt the fields."""
    fresh = _grid(g.x_size_in_kpc, g.y_size_in_kpc, g.z_size_in_kpc, g.resolution)
    if hasattr(g, "fine_resolution"):
        fresh.add_fine_grid(g.fine_x_size_in_kpc, g.fine_y_size_in_kpc, g.fine_z_size_in_kpc, g.fine_resolution)
    if not np.array_equal(np.asarray(g.init_grid), fresh.init_grid):
        raise ValueError("pickled grid: init_grid is not the point list its sizes generate")
    for k in ("snapshot_acceleration_x", "snapshot_acceleration_y", "snapshot_acceleration_z", "snapshot_potential",
              "ss_evolved_position", "evolved_grid"):
        if hasattr(g, k):
            setattr(fresh, k, np.asarray(getattr(g, k), np.float64) if getattr(g, k) is not None else None)
    if not hasattr(fresh, "snapshot_potential"):
        fresh.snapshot_potential = None
    return fresh


def load_grid_pickle(path):
    """Read a whole-grid cache written by the reference (gizmo_interface.py:395-398) or by dump_grid_pickle -> our grid."""
    with open(path, "rb") as fh:
        g = _RefUnpickler(fh).load()
    if not isinstance(g, _grid):
        raise ValueError("%s does not hold a grid_cartesian.grid" % path)
    return _rederive_(g)


# ------------------------------------------------------------------ oceanic_io.dump_interface / load_interface ----
_INTERFACE_KEYS = ("grid_x_size_in_kpc", "grid_y_size_in_kpc", "grid_z_size_in_kpc", "grid_resolution", "fine_grid",
                   "grid_fine_x_size_in_kpc", "grid_fine_y_size_in_kpc", "grid_fine_z_size_in_kpc", "grid_fine_resolution",
                   "star_softening_in_pc", "dark_softening_in_pc", "star_char_mass", "dark_char_mass", "softening_kernel",
                   "plummer_eps_over_h", "theta", "time_interpolation", "space_interpolation", "nclose", "order", "sim_name",
                   "grid_seed", "Rmax", "startnum", "endnum", "num_prior", "chosen_id")


def _dump_(obj, path):
    with open(path, "wb") as fh:
        pickle.dump(obj, fh, protocol=4)   # what dill.dump writes for plain arrays / lists / dicts


def dump_interface(field, directory_out="interface"):
    """The file set of oceanic_io.dump_interface (oceanic_io.py:72-113) for a gizmo_field.  Array files are byte-compatible
    with what the reference reads back (oceanic_io.py:116-125); ``interface`` holds our settings as a dict."""
    out = os.path.join(getattr(field, "output_directory", "."), directory_out)
    os.makedirs(out, exist_ok=True)
    g = field.grid
    _dump_(np.asarray(g.evolved_grid), os.path.join(out, "evolved_grid"))
    _dump_(np.asarray(g.init_grid), os.path.join(out, "init_grid"))
    for c in "xyz":
        _dump_(np.asarray(getattr(g, "snapshot_acceleration_" + c)), os.path.join(out, "snapshot_acceleration_%s.npz" % c))
        # the reference stores one splrep tck per grid point here; all points share the knot vector, so the equivalent is the
        # knot vector + the coefficient planes (time_spline.fit), which load_interface does not need (it refits: "skinny")
        _dump_([], os.path.join(out, "grid_acc%s_interpolators" % c))
    if getattr(g, "snapshot_potential", None) is not None:
        _dump_(np.asarray(g.snapshot_potential), os.path.join(out, "snapshot_potential.npz"))
    state = {k: getattr(field, k, None) for k in _INTERFACE_KEYS}
    state["basis"] = getattr(field.basis, "__name__", None) or str(field.basis)
    state["time_in_Myr"] = np.asarray(field.time_in_Myr)
    state["chosen_snapshot_positions"] = np.asarray(field.chosen_snapshot_positions)
    state["chosen_evolved_position"] = np.asarray(field.chosen_evolved_position)
    state["chosen_evolved_velocity"] = np.asarray(field.chosen_evolved_velocity)
    state["format"] = "oc_nbody_b200 interface v1 (settings dict; the reference stores a dill of its gizmo_interface object here)"
    _dump_(state, os.path.join(out, "interface"))
    return out


def _load_(path):
    with open(path, "rb") as fh:
        return _RefUnpickler(io.BytesIO(fh.read())).load()


def load_interface(directory="interface", skinny=True, ctx=None):
    """oceanic_io.load_interface (oceanic_io.py:115-145): rebuild the field code from the dumped files and leave it at
    t = 0.  Reads the array files whether this package or the reference wrote them (an ``interface`` file written by the
    reference is a dill of its own class and cannot be read without it: pass the settings through `dump_interface`)."""
    from .gizmo_field import gizmo_field
    from .units import units
    state = _load_(os.path.join(directory, "interface"))
    if not isinstance(state, dict) or "time_in_Myr" not in state:
        raise ValueError("%s/interface is not a settings dict written by oc_nbody_b200.cache_compat.dump_interface" % directory)
    opts = {k: state[k] for k in _INTERFACE_KEYS if k != "chosen_id" and state.get(k) is not None}
    opts["basis"] = state["basis"]

    class _Snap(object):
        def __init__(self, i, t):
            self.snapshot = {"index": i, "time": t}
    times = np.asarray(state["time_in_Myr"], np.float64)
    snaps = [_Snap(i, 1e-3 * t) for i, t in enumerate(times)]
    field = gizmo_field(opts, snaps, chosen_positions=state["chosen_snapshot_positions"], time_in_Myr=times,
                        chosen_id=state.get("chosen_id", -1), ctx=ctx, build=False)
    acc = [np.asarray(_load_(os.path.join(directory, "snapshot_acceleration_%s.npz" % c)), np.float64) for c in "xyz"]
    pot_file = os.path.join(directory, "snapshot_potential.npz")
    pot = np.asarray(_load_(pot_file), np.float64) if os.path.exists(pot_file) else None
    field.set_snapshot_fields(acc[0], acc[1], acc[2], pot)
    init = np.asarray(_load_(os.path.join(directory, "init_grid")), np.float64)
    if not np.array_equal(init, field.grid.init_grid):
        raise ValueError("%s/init_grid is not the point list the dumped grid sizes generate" % directory)
    field.chosen_evolved_position = np.asarray(state["chosen_evolved_position"], np.float64)
    field.chosen_evolved_velocity = np.asarray(state["chosen_evolved_velocity"], np.float64)
    field.evolve_grid(field.chosen_evolved_position)
    field.evolve_model(0 | units.Myr)
    return field

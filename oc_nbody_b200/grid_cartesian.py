"""Target lattice that rides with the cluster — host-side mirror of the reference's ``grid`` class.

Same public surface as grid_cartesian.py:15-69 of gusbeane/oc_nbody (attribute and method names, node
count rule, point ordering) so that code written against the reference keeps working, but built
vectorised and with the lattice description the CUDA interpolation kernel needs (``nodes``,
``shape``, ``origin_row``) instead of only the flattened point list.

Layout contract (bit-exact with the reference, pinned by tests/golden/grid_*.npz):
  * nodes per axis  n_d = int(L_d / resolution)                       (grid_cartesian.py:25-27)
  * node coordinates np.linspace(-L_d, L_d, num=n_d)                  (grid_cartesian.py:29-31)
    -> the true spacing is 2 L_d / (n_d - 1), NOT `resolution`
  * init_grid[(i*ny + j)*nz + k] = (x_grid[i], y_grid[j], z_grid[k])  (grid_cartesian.py:59-65)
  * one extra row (0, 0, 0) appended at index nx*ny*nz                (grid_cartesian.py:66-67)
    so that after the frame subtraction its acceleration is exactly 0.

Nested fine grid (``add_fine_grid``, grid_cartesian.py:34-53,71-91; the reference's default configuration,
test_options:93-100): every coarse point strictly inside the fine box (|x_d| < L_fine_d on all axes — this
includes the coarse origin row) is dropped, the fine lattice is appended in C order, then ONE origin row:
    init_grid = [ kept coarse points | fine lattice | (0,0,0) ]
As in the reference, ``x_n, y_n, z_n`` are overwritten with the FINE node counts; the coarse counts stay
available as ``coarse_shape``.  The CUDA interpolation kernel works on two regular lattices, so the class also
exposes ``coarse_keep_index`` / ``coarse_hole_index`` (positions of the kept / dropped points in the full coarse
lattice) and ``fine_row0``.
"""
import numpy as np


class grid(object):
    def __init__(self, x_size_in_kpc, y_size_in_kpc, z_size_in_kpc, resolution):
        self.x_size_in_kpc = x_size_in_kpc
        self.y_size_in_kpc = y_size_in_kpc
        self.z_size_in_kpc = z_size_in_kpc
        self.resolution = resolution

        half = (x_size_in_kpc, y_size_in_kpc, z_size_in_kpc)
        counts = tuple(int(L / resolution) for L in half)
        if min(counts) < 2:
            raise ValueError("grid needs at least 2 nodes per axis for interpolation; int(L/res) = %r" % (counts,))
        self.x_n, self.y_n, self.z_n = counts
        self.coarse_shape = counts
        self.fine_shape = None
        self.x_grid, self.y_grid, self.z_grid = (np.linspace(-L, L, num=n) for L, n in zip(half, counts))
        self._gen_init_grid_()
        self.gen_evolved_grid(np.zeros(3))

    # -- lattice description consumed by the CUDA kernels --
    @property
    def shape(self):
        """Node counts of the (coarse) lattice."""
        return self.coarse_shape

    @property
    def nodes(self):
        return (self.x_grid, self.y_grid, self.z_grid)

    @property
    def fine_nodes(self):
        return (self.x_fine_grid, self.y_fine_grid, self.z_fine_grid)

    @property
    def has_fine_grid(self):
        return self.fine_shape is not None

    @property
    def n_lattice(self):
        """Points of the full coarse lattice (before any are dropped for a fine grid)."""
        return int(np.prod(self.coarse_shape))

    @property
    def origin_row(self):
        """Row of init_grid holding the appended (0,0,0) point (always the last one)."""
        return self.init_grid.shape[0] - 1

    def __len__(self):
        return self.init_grid.shape[0]

    @staticmethod
    def _lattice_points_(xg, yg, zg):
        pts = np.empty((len(xg), len(yg), len(zg), 3), dtype=np.float64)
        pts[..., 0] = xg[:, None, None]
        pts[..., 1] = yg[None, :, None]
        pts[..., 2] = zg[None, None, :]
        return pts.reshape(-1, 3)

    def _gen_init_grid_(self):
        pts = np.empty((self.n_lattice + 1, 3), dtype=np.float64)
        pts[:-1] = self._lattice_points_(self.x_grid, self.y_grid, self.z_grid)
        pts[-1] = 0.0  # origin keeps the total acceleration on the cluster zero
        self.init_grid = pts

    def add_fine_grid(self, x_size_in_kpc, y_size_in_kpc, z_size_in_kpc, resolution):
        """Nested fine lattice around the origin (grid_cartesian.py:34-53): drop the coarse points strictly inside
        the fine box (``_remove_coarse_points_``, :71-81), append the fine lattice and one origin row
        (``_add_fine_grid_``, :83-91)."""
        if self.has_fine_grid:
            raise ValueError("a fine grid has already been added")
        self.fine_x_size_in_kpc = x_size_in_kpc
        self.fine_y_size_in_kpc = y_size_in_kpc
        self.fine_z_size_in_kpc = z_size_in_kpc
        self.fine_resolution = resolution
        half = (x_size_in_kpc, y_size_in_kpc, z_size_in_kpc)
        counts = tuple(int(L / resolution) for L in half)
        if min(counts) < 2:
            raise ValueError("fine grid needs at least 2 nodes per axis; int(L/res) = %r" % (counts,))
        for L, Lc in zip(half, (self.x_size_in_kpc, self.y_size_in_kpc, self.z_size_in_kpc)):
            if not L < Lc:
                raise ValueError("the fine box must lie inside the coarse one")
        # the reference overwrites the coarse counts here (grid_cartesian.py:41-43); mirrored on purpose
        self.x_n, self.y_n, self.z_n = counts
        self.fine_shape = counts
        self.x_fine_grid, self.y_fine_grid, self.z_fine_grid = (np.linspace(-L, L, num=n) for L, n in zip(half, counts))
        # _remove_coarse_points_: the appended coarse origin row satisfies |0| < L and goes too
        inside = np.ones(self.init_grid.shape[0], bool)
        for d, L in enumerate(half):
            inside &= np.abs(self.init_grid[:, d]) < L
        keep = np.where(np.logical_not(inside))[0]
        lattice_inside = inside[:-1]
        self.coarse_keep_index = np.where(np.logical_not(lattice_inside))[0].astype(np.int64)
        self.coarse_hole_index = np.where(lattice_inside)[0].astype(np.int64)
        fine = self._lattice_points_(self.x_fine_grid, self.y_fine_grid, self.z_fine_grid)
        self.fine_row0 = len(keep)
        self.init_grid = np.ascontiguousarray(np.concatenate([self.init_grid[keep], fine, np.zeros((1, 3))]))
        self.gen_evolved_grid(getattr(self, "ss_evolved_position", np.zeros(3)))

    def gen_evolved_grid(self, position):
        """evolved_grid = init_grid + position (grid_cartesian.py:55-57)."""
        self.ss_evolved_position = position
        self.evolved_grid = np.add(self.init_grid, position)

    # -- point list <-> lattice records (what K3 gathers from) --
    def coarse_hole_points(self):
        """Coordinates (un-shifted) of the coarse lattice points the fine grid replaced, in lattice order."""
        full = self._lattice_points_(self.x_grid, self.y_grid, self.z_grid)
        return full[self.coarse_hole_index]

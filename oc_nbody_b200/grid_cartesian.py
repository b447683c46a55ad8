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
        self.x_grid, self.y_grid, self.z_grid = (np.linspace(-L, L, num=n) for L, n in zip(half, counts))
        self._gen_init_grid_()
        self.gen_evolved_grid(np.zeros(3))

    # -- lattice description consumed by the CUDA kernels --
    @property
    def shape(self):
        return (self.x_n, self.y_n, self.z_n)

    @property
    def nodes(self):
        return (self.x_grid, self.y_grid, self.z_grid)

    @property
    def n_lattice(self):
        return self.x_n * self.y_n * self.z_n

    @property
    def origin_row(self):
        """Row of init_grid holding the appended (0,0,0) point."""
        return self.n_lattice

    def __len__(self):
        return self.n_lattice + 1

    def _gen_init_grid_(self):
        pts = np.empty((self.n_lattice + 1, 3), dtype=np.float64)
        lattice = pts[:-1].reshape(self.x_n, self.y_n, self.z_n, 3)
        lattice[..., 0] = self.x_grid[:, None, None]
        lattice[..., 1] = self.y_grid[None, :, None]
        lattice[..., 2] = self.z_grid[None, None, :]
        pts[-1] = 0.0  # origin keeps the total acceleration on the cluster zero
        self.init_grid = pts

    def gen_evolved_grid(self, position):
        """evolved_grid = init_grid + position (grid_cartesian.py:55-57)."""
        self.ss_evolved_position = position
        self.evolved_grid = np.add(self.init_grid, position)

    def add_fine_grid(self, *args, **kwargs):
        raise NotImplementedError(
            "nested fine grid (grid_cartesian.py:34-53,71-91) is a SURVEY §8(f) 'next' row; the CUDA "
            "interpolation kernel works on one regular lattice")

"""Drop-in for the two pykdgrav calls of the reference's field build.

The reference does (gizmo_interface.py:15,561,564,566)::

    from pykdgrav import ConstructKDTree, GetAccelParallel
    tree = ConstructKDTree(np.float64(r), np.float64(m), np.float64(soft))
    accel = GetAccelParallel(grid.evolved_grid, tree, self.G, self.theta)

Changing the import to ``from oc_nbody_b200.pykdgrav_compat import ConstructKDTree, GetAccelParallel`` keeps
that code as it is and runs the sum on a B200: "the tree" becomes the device-resident source set and the
walk becomes the exact (theta -> 0) direct sum of liboc_nbody_b200 (K1).  ``theta`` is accepted and
ignored.  pykdgrav's softening argument is the compact-support radius of its cubic-spline kernel, so the
default ``kernel`` here is ``"spline"``; ``"plummer"`` uses the same lengths as Plummer epsilons.
"""
import numpy as np

from . import _lib


class DeviceSources:
    """What ConstructKDTree returns: FP64 sources resident in HBM (torch tensors as plain buffers)."""

    def __init__(self, x, m, softening, ctx=None):
        import torch
        self.ctx = ctx or _lib.default_context()
        dev = torch.device("cuda", self.ctx.device)
        x = np.ascontiguousarray(x, np.float64).reshape(-1, 3)
        m = np.ascontiguousarray(m, np.float64)
        if m.shape[0] != x.shape[0]:
            raise ValueError("positions and masses differ in length: %d vs %d" % (x.shape[0], m.shape[0]))
        self.n = x.shape[0]
        self.pos = torch.from_numpy(x).to(dev)
        self.mass = torch.from_numpy(m).to(dev)
        if softening is None:
            self.soft32 = None
        else:
            s = np.ascontiguousarray(softening, np.float64)
            if s.shape[0] != self.n:
                raise ValueError("softening has %d entries for %d particles" % (s.shape[0], self.n))
            self.soft32 = torch.empty(self.n, dtype=torch.float32, device=dev)
            self.ctx.cast_f64_f32(torch.from_numpy(s).to(dev), self.soft32)
        self._xyzm32 = torch.empty((self.n, 4), dtype=torch.float32, device=dev)
        self._center = None

    def recentred(self, center):
        """FP32 (x - c, m) records, recentred in FP64 first; cached per centre."""
        c = tuple(float(v) for v in center)
        if self._center != c:
            self.ctx.recentre_f64(self.pos, self.mass, c, self._xyzm32)
            self._center = c
        return self._xyzm32


def ConstructKDTree(x, m, softening=None, ctx=None):
    """pykdgrav.ConstructKDTree(x, m, softening) -> handle to the device-resident source set."""
    return DeviceSources(x, m, softening, ctx)


def _accel(x_target, tree, G, kernel, center, want_pot):
    import torch
    ctx = tree.ctx
    dev = tree.pos.device
    tgt = np.ascontiguousarray(x_target, np.float64).reshape(-1, 3)
    n = tgt.shape[0]
    if center is None:
        center = 0.5 * (tgt.min(axis=0) + tgt.max(axis=0)) if n else np.zeros(3)
    src32 = tree.recentred(center)
    tgt64 = torch.from_numpy(tgt).to(dev)
    tgt32 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ctx.recentre_f64(tgt64, None, center, tgt32)
    acc = torch.empty((3, n), dtype=torch.float64, device=dev)
    pot = torch.empty(n, dtype=torch.float64, device=dev) if want_pot else None
    ctx.field_direct(src32, tree.soft32, tgt32, _lib.KERNELS[kernel], G, acc, pot)
    return acc, pot


def GetAccelParallel(x_target, tree, G=1.0, theta=0.7, kernel="spline", center=None):
    """pykdgrav.GetAccelParallel(x_target, tree, G, theta) -> [n_target, 3] FP64 accelerations."""
    acc, _ = _accel(x_target, tree, G, kernel, center, False)
    return np.ascontiguousarray(acc.cpu().numpy().T)


def GetPotentialParallel(x_target, tree, G=1.0, theta=0.7, kernel="spline", center=None):
    """pykdgrav.GetPotentialParallel equivalent -> [n_target] FP64 potentials."""
    _, pot = _accel(x_target, tree, G, kernel, center, True)
    return pot.cpu().numpy()

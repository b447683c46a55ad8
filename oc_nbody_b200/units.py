"""Minimal AMUSE-style units so the plugin API keeps the reference's calling convention.

The reference passes AMUSE quantities across its plugin boundary: ``x.value_in(units.kpc)``
(gizmo_interface.py:679-681), ``time.value_in(units.Myr)`` (gizmo_interface.py:623) and returns
``list | units.kms/units.Myr`` (gizmo_interface.py:706-708).  AMUSE is not installable here, so this shim
provides the same ``value | unit`` / ``.value_in(unit)`` surface for the handful of units the hot path
touches.  If ``amuse`` is importable its own ``units`` module is used instead and this file is bypassed.
"""
import numpy as np

try:  # pragma: no cover - AMUSE absent in this image
    from amuse.units import units as _amuse_units
    HAVE_AMUSE = True
except Exception:  # noqa: BLE001
    _amuse_units = None
    HAVE_AMUSE = False

KM_PER_PC = 3.0856775814913673e13
SEC_PER_MYR = 3.15576e13
# pc travelled per Myr at 1 km/s
KMS_TO_PC_PER_MYR = SEC_PER_MYR / KM_PER_PC
# G in kpc^2 (km/s) / Myr / Msun — the unit of gizmo_interface.py:70 — from GM_sun = 1.32712440018e20 m^3/s^2
G_KPC_KMS_MYR = 1.32712440018e20 / (KM_PER_PC * 1e3 * 1e3) ** 2 * 1e-3 * SEC_PER_MYR
# G in pc (km/s)^2 / Msun
G_PC_KMS2 = 1.32712440018e20 / (KM_PER_PC * 1e3) * 1e-6


class Unit:
    """A named unit with dimension exponents over (length[kpc], time[Myr], mass[MSun]) and a scale."""

    def __init__(self, name, scale, dims):
        self.name, self.scale, self.dims = name, float(scale), tuple(dims)

    def __mul__(self, o):
        return Unit("%s*%s" % (self.name, o.name), self.scale * o.scale, [a + b for a, b in zip(self.dims, o.dims)])

    def __truediv__(self, o):
        return Unit("%s/%s" % (self.name, o.name), self.scale / o.scale, [a - b for a, b in zip(self.dims, o.dims)])

    def __pow__(self, k):
        return Unit("%s**%d" % (self.name, k), self.scale ** k, [a * k for a in self.dims])

    # make `ndarray | unit` defer to __ror__ instead of numpy's bitwise_or
    __array_ufunc__ = None

    def __ror__(self, value):  # value | unit
        return Quantity(value, self)

    def __repr__(self):
        return self.name


class Quantity:
    def __init__(self, value, unit):
        self.number = np.asarray(value, dtype=np.float64) if np.ndim(value) else float(value)
        self.unit = unit

    def value_in(self, unit):
        if tuple(unit.dims) != tuple(self.unit.dims):
            raise ValueError("cannot convert %r to %r" % (self.unit, unit))
        return self.number * (self.unit.scale / unit.scale)

    def __len__(self):
        return len(self.number)

    def __getitem__(self, i):
        return Quantity(self.number[i], self.unit)

    def __add__(self, o):
        return Quantity(self.number + o.value_in(self.unit), self.unit)

    def __sub__(self, o):
        return Quantity(self.number - o.value_in(self.unit), self.unit)

    def __mul__(self, k):
        if isinstance(k, Quantity):
            return Quantity(self.number * k.number, self.unit * k.unit)
        return Quantity(self.number * k, self.unit)

    __rmul__ = __mul__

    def __pow__(self, k):
        return Quantity(self.number ** k, self.unit ** k)

    def __repr__(self):
        return "%r | %r" % (self.number, self.unit)


class _Units:
    kpc = Unit("kpc", 1.0, (1, 0, 0))
    parsec = Unit("parsec", 1e-3, (1, 0, 0))
    Myr = Unit("Myr", 1.0, (0, 1, 0))
    MSun = Unit("MSun", 1.0, (0, 0, 1))
    # 1 km/s in kpc/Myr
    kms = Unit("kms", KMS_TO_PC_PER_MYR * 1e-3, (1, -1, 0))


units = _amuse_units if HAVE_AMUSE else _Units


def to_value(q, unit):
    """Unwrap an AMUSE/shim quantity, or pass a bare float/array through (assumed already in `unit`)."""
    if hasattr(q, "value_in"):
        return q.value_in(unit)
    return q

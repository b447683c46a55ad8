"""oc_nbody_b200 — B200-native (sm_100a) gravity hot path for oceanic (gusbeane/oc_nbody).

Host-side mirror of the reference's plugin interfaces over the C ABI of ``lib/liboc_nbody_b200.so``
(include/ocg.h).  There is no CPU fallback: importing is cheap, calling needs the built library and a B200.
"""
from ._lib import KERNEL_PLUMMER, KERNEL_SPLINE, Context, OcgError, default_context, load_library  # noqa: F401
from .grid_cartesian import grid  # noqa: F401

__all__ = ["Context", "OcgError", "default_context", "load_library", "grid", "KERNEL_PLUMMER", "KERNEL_SPLINE"]

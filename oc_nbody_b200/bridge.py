"""Minimal BRIDGE (Fujii et al. 2007) with the call pattern of ``amuse.couple.bridge.Bridge`` as the reference
uses it (oc_nbody.py:49-56): ``Bridge(timestep=..., use_threading=False)``, ``add_system(code, partners)``,
``evolve_model(tend, timestep=...)``, ``.particles``.

One step is K(dt/2) D(dt) K(dt/2): every system with partners is kicked by its partners' gravity at its
particles' positions, then every system drifts (``evolve_model(time + dt)``), then the second half kick.
AMUSE itself is not installable in this image; this file exists so the driver loop of
oc_nbody.py runs unchanged against the B200 codes, and so the oracle has the same step order to follow.

When both the kicked system and its partner are this package's GPU codes the kick stays on the device
(``partner.kick_device``): K3 gather + K5 update, no host copies.  Any other partner is kicked through the
public ``get_gravity_at_point(eps, x, y, z)``.
"""
from .units import to_value, units


class Bridge(object):
    def __init__(self, timestep=None, use_threading=False, verbose=False):
        if use_threading:
            raise NotImplementedError("the reference runs the bridge single-threaded (oc_nbody.py:49)")
        self.timestep = None if timestep is None else float(to_value(timestep, units.Myr))
        self.systems = []
        self.partners = {}
        self.time = 0.0

    def add_system(self, system, partners=()):
        self.systems.append(system)
        self.partners[id(system)] = tuple(partners)

    @property
    def particles(self):
        for s in self.systems:
            if hasattr(s, "particles"):
                return s.particles
        raise AttributeError("no system with particles")

    def kick_systems(self, dt):
        for s in self.systems:
            for p in self.partners[id(s)]:
                if hasattr(p, "kick_device") and hasattr(s, "pos") and hasattr(s, "vel"):
                    p.kick_device(s.pos, s.vel, dt)
                else:
                    parts = s.particles
                    ax, ay, az = p.get_gravity_at_point(0.0 | units.kpc, parts.x, parts.y, parts.z)
                    s.kick_velocities(ax, ay, az, dt)

    def drift_systems(self, tend):
        for s in self.systems:
            s.evolve_model(tend | units.Myr)

    def evolve_model(self, tend, timestep=None):
        tend = float(to_value(tend, units.Myr))
        dt = self.timestep if timestep is None else float(to_value(timestep, units.Myr))
        if dt is None:
            dt = tend - self.time
        while self.time < tend - 0.5 * dt:
            self.kick_systems(0.5 * dt)
            self.drift_systems(self.time + dt)
            self.kick_systems(0.5 * dt)
            self.time += dt

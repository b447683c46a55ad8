"""The reference's driver loop (oc_nbody.py:15-70, ``evolve_cluster_in_galaxy``) on top of the B200 codes.

The reference's own driver "stays as it is" (BASELINE.json north_star): with the three import changes of
INTEGRATION.md it runs unchanged.  It cannot run HERE — AMUSE, gizmo_analysis and the FIRE snapshots are absent — so
this module restates the same loop over this package's codes and in-memory snapshots, step for step:

    system.evolve_model(t, timestep)                 oc_nbody.py:56   K(dt/2) D(dt) K(dt/2)
    cluster.clean_ejections(system)                  oc_nbody.py:58   median-centred radius cut (oc_code.py:231-246)
    bound = particles.bound_subset(); bound.center_of_mass()   :60-61
    galaxy_code.evolve_grid(bound_com)               oc_nbody.py:64   the grid follows the cluster
    snap_reader.process_snapshot(...)                oc_nbody.py:66   one frame per step (oceanic_io.py:42-70)

It is the end-to-end example of the package and what the configs[0] integration test drives.
"""
import numpy as np

from .bridge import Bridge
from .cluster import cluster_code
from .units import units


class frame_recorder(object):
    """Frames with the keys of the reference's ``snapshot_reader._grab_frame_`` (oceanic_io.py:42-70): time [Myr],
    position [n,3] pc, velocity [n,3] km/s, mass Msun, com kpc, chosen_position pc, chosen_velocity km/s.  Kept in memory;
    ``write_frequency`` / dill output stay with the reference's oceanic_io."""

    def __init__(self, galaxy_code, every=1):
        self.frames = []
        self.every = max(1, int(every))
        self.meta = dict(ss_id=getattr(galaxy_code, "chosen_id", None), snapshot_times=getattr(galaxy_code, "time_in_Myr", None),
                         sim_name=getattr(galaxy_code, "sim_name", None))

    def process_snapshot(self, system, galaxy_code, com, i, time):
        if i % self.every:
            return
        p = system.particles
        self.frames.append({
            "time": float(time.value_in(units.Myr)) if hasattr(time, "value_in") else float(time),
            "position": np.transpose([p.x.value_in(units.parsec), p.y.value_in(units.parsec), p.z.value_in(units.parsec)]),
            "velocity": np.transpose([p.vx.value_in(units.kms), p.vy.value_in(units.kms), p.vz.value_in(units.kms)]),
            "mass": p.mass.value_in(units.MSun),
            "com": np.asarray(com, np.float64).copy(),
            "chosen_position": np.asarray(galaxy_code.chosen_evolved_position) * 1000.0,
            "chosen_velocity": np.asarray(galaxy_code.chosen_evolved_velocity),
        })


def evolve_cluster_in_galaxy(galaxy_code, mass, pos_kpc, vel_kms, timestep, tend, softening_pc=0.01, eject_cut=None,
                             substeps=1, use_cuda_graph=True, record_every=1, ctx=None, integrator="leapfrog"):
    """oc_nbody.py:15-70 with the cluster given as arrays (the reference builds it from a King model + Kroupa IMF through
    AMUSE, oc_code.py:197-216) and ``galaxy_code`` an already built ``gizmo_field``.  timestep / tend in Myr
    (options 'timestep', 'tend', oc_nbody.py:18-20); integrator "hermite" drifts with ph4's own scheme (K6).
    Returns (cluster_code, frame_recorder)."""
    cl = cluster_code(mass, pos_kpc, vel_kms, softening_pc=softening_pc, substeps=substeps, eject_cut=eject_cut,
                      ctx=ctx or galaxy_code.ctx, integrator=integrator)
    times = np.arange(0.0, float(tend), float(timestep))            # oc_nbody.py:20
    system = Bridge(timestep=timestep | units.Myr, use_threading=False, use_cuda_graph=use_cuda_graph)  # :49
    system.add_system(cl, (galaxy_code,))                           # :50  the cluster is kicked by the galaxy
    system.add_system(galaxy_code)                                  # :51  the galaxy only drifts
    rec = frame_recorder(galaxy_code, every=record_every)
    for i, t in enumerate(times):                                   # :55
        system.evolve_model(t | units.Myr, timestep=timestep | units.Myr)   # :56 (the first call, t = 0, is a no-op)
        cl.clean_ejections(system)                                  # :58
        bound_com = cl.bound_center_of_mass()                       # :60-61
        galaxy_code.evolve_grid(bound_com)                          # :64
        rec.process_snapshot(system, galaxy_code, bound_com, i, t | units.Myr)   # :66
    return cl, rec

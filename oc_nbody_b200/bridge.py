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

``use_cuda_graph=True``: for that device-resident pair the whole step (K3, K5, K4 = pack + stream-K kernel,
K5 ... — 9 launches) is captured once into a CUDA graph and replayed with ONE launch per step.  What changes from
step to step — the time-blend weights of the two half-kicks — lives in constant memory and is refreshed before each
replay (ocg_set_interp_weight_slots); the grid origin is a device buffer.  The graph is re-captured when the
bracketing snapshots, the particle count, the timestep or the capture epoch of a context (scratch reallocated, another
work plan uploaded: include/ocg.h ocg_capture_epoch) change.  Results equal the eager path's to FP64 rounding
(the eager drift integrates over (t + dt) - model_time, which can differ from dt in the last bit).
The star-sharded cluster code is captured the same way, its NCCL all-gathers included (every rank takes the same
eager / capture / replay decisions, so the collectives stay matched).
"""
from .units import to_value, units


class Bridge(object):
    def __init__(self, timestep=None, use_threading=False, verbose=False, use_cuda_graph=False):
        if use_threading:
            raise NotImplementedError("the reference runs the bridge single-threaded (oc_nbody.py:49)")
        self.timestep = None if timestep is None else float(to_value(timestep, units.Myr))
        self.systems = []
        self.partners = {}
        self.time = 0.0
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph = None
        self._graph_key = None
        self._seen_key = None
        self.graph_replays = 0
        self.graph_captures = 0

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

    # ---- one-launch step for the device-resident cluster + field pair ----
    def _device_pair_(self):
        """(cluster, field) when the bridge holds exactly this package's cluster code kicked by its field code."""
        if len(self.systems) != 2:
            return None
        for cl, fld in (self.systems, self.systems[::-1]):
            if (self.partners[id(cl)] == (fld,) and self.partners[id(fld)] == () and hasattr(fld, "kick_device")
                    and hasattr(fld, "_time_planes_") and getattr(fld, "space_interpolation", "trilinear") == "trilinear"
                    and type(cl).__name__ in ("cluster_code", "sharded_cluster_code") and hasattr(cl, "_evolve_device_")
                    and not getattr(cl, "block_steps", False)):  # block steps read back the schedule: not capturable
                return cl, fld
        return None

    def _graph_step_(self, cl, fld, dt):
        """K(dt/2) D(dt) K(dt/2) as one CUDA-graph launch.  The first step with a new configuration runs eagerly (it
        also sizes every scratch buffer), the second captures, later ones replay."""
        import torch
        t0 = self.time
        if abs(getattr(fld, "time", t0) - t0) > 1e-9 * max(1.0, abs(t0)) or not cl._acc_valid:
            return False
        if getattr(cl, "auto_substeps", False) and not cl._auto_started:
            cl._first_substeps_(dt)
        rec0, fine0, w0 = fld._time_planes_()
        fld.evolve_model((t0 + dt) | units.Myr)          # host only: bracket + weights of the second half-kick
        rec1, fine1, w1 = fld._time_planes_()
        # what a captured graph froze: record planes, particle buffers, step parameters — and, through the capture epochs,
        # the scratch addresses and resident work plans of the two contexts (another call on a shared ctx may have grown a
        # buffer or uploaded another cluster's plan since: then this step runs eagerly and the next one re-captures)
        key = (tuple(r.data_ptr() for r in rec0), tuple(r.data_ptr() for r in rec1), cl.n, cl.pos.data_ptr(), cl.substeps,
               dt, cl.parameters._eps2_kpc2, fld._dev["origin"].data_ptr(), cl.ctx.capture_epoch(), fld.ctx.capture_epoch())

        def body():
            fld.kick_device(cl.pos, cl.vel, 0.5 * dt, planes=(rec0, fine0), w_slot=0)
            cl._evolve_device_(dt)
            fld.kick_device(cl.pos, cl.vel, 0.5 * dt, planes=(rec1, fine1), w_slot=1)

        cl.ctx.set_interp_weight_slots([w0, w1])
        if key == self._graph_key:
            self._graph.replay()
            self.graph_replays += 1
        elif key == self._seen_key:
            # capture.  On a ctx shared with another particle set the resident work plan may be the other one's, and a
            # plan upload is not capturable: re-evaluate the current force eagerly first (same inputs, same values), which
            # leaves this cluster's plan resident, and take the key with the epoch that results
            cl._prepare_capture_()
            key = key[:-2] + (cl.ctx.capture_epoch(), fld.ctx.capture_epoch())
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                body()
            self._graph_key = key
            self.graph_captures += 1
            self._graph.replay()
            self.graph_replays += 1
        else:
            self._seen_key = key
            body()
        cl.model_time = t0 + dt
        if getattr(cl, "auto_substeps", False):
            cl._update_substeps_(dt)   # a changed count changes the key: the next step runs eagerly, the one after re-captures
        return True

    def evolve_model(self, tend, timestep=None):
        tend = float(to_value(tend, units.Myr))
        dt = self.timestep if timestep is None else float(to_value(timestep, units.Myr))
        if dt is None:
            dt = tend - self.time
        pair = self._device_pair_() if self.use_cuda_graph else None
        while self.time < tend - 0.5 * dt:
            if pair is None or not self._graph_step_(pair[0], pair[1], dt):
                self.kick_systems(0.5 * dt)
                self.drift_systems(self.time + dt)
                self.kick_systems(0.5 * dt)
            self.time += dt

"""Cluster N-body code: the B200 drop-in for the AMUSE ``ph4`` worker the reference instantiates in
oc_code.py:218-229, behind the same gravity-code duck type the driver uses (oc_nbody.py:22-23,28,46-53,70):
``.particles``, ``.parameters.epsilon_squared``, ``.evolve_model(t)``, ``.stop()``.

Algorithm (BASELINE.json north_star): Plummer-softened direct-sum self-gravity (K4) + kick-drift-kick
leapfrog (K5) in place of ph4's 4th-order Hermite with block time steps.  ``integrator="hermite"`` selects ph4's
own scheme instead: acceleration + jerk direct sum (K6), 4th-order Hermite predictor /
corrector with one shared step, Aarseth step criterion evaluated on the device.  State lives in HBM as FP64
component-major arrays in the field code's unit system (kpc, km/s, Myr, Msun) so that the BRIDGE kick (K3)
consumes the positions without conversion or host copies.

Also here: ``clean_ejections`` (oc_code.py:231-246) and the bound-subset centre of mass the driver feeds to
``evolve_grid`` (oc_nbody.py:60-64).
"""
import numpy as np

from . import _lib
from .units import G_KPC_KMS_MYR, KMS_TO_PC_PER_MYR, to_value, units

KMS_TO_KPC_PER_MYR = KMS_TO_PC_PER_MYR * 1e-3


class _Parameters(object):
    """``code.parameters.epsilon_squared = (softening | units.parsec)**2`` (oc_code.py:225-228)."""

    def __init__(self):
        self._eps2_kpc2 = 0.0

    @property
    def epsilon_squared(self):
        return self._eps2_kpc2 | units.kpc ** 2

    @epsilon_squared.setter
    def epsilon_squared(self, q):
        v = to_value(q, units.kpc ** 2)
        if not v >= 0.0:
            raise ValueError("epsilon_squared must be >= 0")
        self._eps2_kpc2 = float(v)


class ParticleView(object):
    """Numpy-side view of the particle attributes with AMUSE-like unit-carrying accessors."""

    def __init__(self, mass, pos_kpc, vel_kms, key=None):
        self._m = np.asarray(mass, np.float64)
        self._x = np.asarray(pos_kpc, np.float64)
        self._v = np.asarray(vel_kms, np.float64)
        self.key = np.arange(self._m.shape[0]) if key is None else np.asarray(key)

    def __len__(self):
        return self._m.shape[0]

    mass = property(lambda s: s._m | units.MSun)
    x = property(lambda s: s._x[0] | units.kpc)
    y = property(lambda s: s._x[1] | units.kpc)
    z = property(lambda s: s._x[2] | units.kpc)
    vx = property(lambda s: s._v[0] | units.kms)
    vy = property(lambda s: s._v[1] | units.kms)
    vz = property(lambda s: s._v[2] | units.kms)
    position = property(lambda s: s._x.T | units.kpc)
    velocity = property(lambda s: s._v.T | units.kms)

    def __getitem__(self, idx):
        return ParticleView(self._m[idx], self._x[:, idx], self._v[:, idx], self.key[idx])

    def copy(self):
        return ParticleView(self._m.copy(), self._x.copy(), self._v.copy(), self.key.copy())

    def center_of_mass(self):
        return (self._x * self._m).sum(axis=1) / self._m.sum() | units.kpc

    def center_of_mass_velocity(self):
        return (self._v * self._m).sum(axis=1) / self._m.sum() | units.kms


class cluster_code(object):
    """Direct-sum leapfrog cluster integrator on one GPU (or a target shard of one, see distributed.py).

    mass [n] Msun, pos [3,n] kpc, vel [3,n] km/s (floats or unit-carrying); ``softening_pc`` as in
    oc_code.py:225; ``substeps`` leapfrog steps per evolve_model call (ph4 chooses its own block steps;
    here the BRIDGE timestep is subdivided evenly).  With ``integrator="hermite"``, ``substeps="auto"`` lets the Aarseth
    criterion choose: the count (a power of two, at most ``max_substeps``) follows the minimum step measured on the
    device during the previous call, and eta * min |a| / |j| before the first one — ph4's accuracy control applied to
    one shared step."""

    def __init__(self, mass, pos, vel, softening_pc=0.01, substeps=1, eject_cut=None, ctx=None, integrator="leapfrog",
                 eta=0.14, block_steps=False, max_level=12):
        import torch
        if integrator not in ("leapfrog", "hermite"):
            raise ValueError("integrator must be 'leapfrog' or 'hermite', not %r" % (integrator,))
        self.integrator = integrator
        self.eta = float(eta)  # Aarseth accuracy parameter (ph4's timestep_parameter, default 0.14)
        # ph4's individual block time steps (integrator="hermite" only): every star on its own step span / 2^k, k <= max_level
        self.block_steps = bool(block_steps)
        self.max_level = int(max_level)
        if self.block_steps and integrator != "hermite":
            raise ValueError("block_steps needs integrator='hermite'")
        self.block_step_count = self.star_step_count = 0
        # a ctx of its own unless the caller shares one: a ctx is only scratch + settings, and a BRIDGE step captured as a
        # CUDA graph freezes the scratch addresses and the resident work plan (ocg_capture_epoch, include/ocg.h)
        self.ctx = ctx or _lib.private_context()
        self._dev = torch.device("cuda", self.ctx.device)
        self.parameters = _Parameters()
        self.parameters.epsilon_squared = (softening_pc | units.parsec) ** 2
        self.auto_substeps = substeps == "auto"
        if self.auto_substeps and integrator != "hermite":
            raise ValueError("substeps='auto' needs integrator='hermite' (the step criterion uses the jerk)")
        self.substeps = 1 if self.auto_substeps else int(substeps)
        self.max_substeps = 1 << 14
        self._auto_started = False
        self.eject_cut = eject_cut  # pc, oc_code.py:241
        self.G = G_KPC_KMS_MYR
        self.model_time = 0.0
        self.key = None
        def unwrap(v, unit):
            # [3, n] array / quantity, or a sequence of three per-axis vectors (AMUSE's particles.x, .y, .z)
            if isinstance(v, (list, tuple)) and len(v) == 3:
                return np.stack([np.asarray(to_value(c, unit), np.float64) for c in v])
            return np.asarray(to_value(v, unit), np.float64)

        self._set_state_(np.asarray(to_value(mass, units.MSun), np.float64), unwrap(pos, units.kpc), unwrap(vel, units.kms))

    # ---- state ----
    def _set_state_(self, mass, pos, vel, key=None):
        import torch
        n = mass.shape[0]
        if pos.shape != (3, n) or vel.shape != (3, n):
            raise ValueError("pos and vel must be [3, n] with n = len(mass) = %d" % n)
        self.n = n
        self.mass = torch.from_numpy(np.ascontiguousarray(mass)).to(self._dev)
        self.pos = torch.from_numpy(np.ascontiguousarray(pos)).to(self._dev)
        self.vel = torch.from_numpy(np.ascontiguousarray(vel)).to(self._dev)
        self.acc = torch.empty((3, n), dtype=torch.float64, device=self._dev)
        self.pot = torch.empty(n, dtype=torch.float64, device=self._dev)
        self.key = np.arange(n) if key is None else np.asarray(key)
        self._acc_valid = False
        self._alloc_hermite_()

    def _alloc_hermite_(self):
        import torch
        if self.integrator != "hermite":
            return
        n = self.n
        mk = lambda: torch.empty((3, n), dtype=torch.float64, device=self._dev)
        self.jerk, self.acc1, self.jerk1, self.pos_p, self.vel_p = mk(), mk(), mk(), mk(), mk()
        self.dt_min = torch.full((1,), float("inf"), dtype=torch.float64, device=self._dev)

    @property
    def particles(self):
        return ParticleView(self.mass.cpu().numpy(), self.pos.cpu().numpy(), self.vel.cpu().numpy(), self.key)

    def remove_particles(self, idx):
        """Drop particles by index (system.particles.remove_particles, oc_code.py:241-242)."""
        import torch
        keep = torch.ones(self.n, dtype=torch.uint8, device=self._dev)
        keep[torch.as_tensor(np.asarray(idx, np.int64), device=self._dev)] = 0
        self._compact_(keep)

    def _compact_(self, keep):
        """Stable on-device compaction of mass / pos / vel by a byte mask (ocg_compact_rows); returns the removed indices."""
        import torch
        rows = torch.cat([self.mass[None], self.pos, self.vel], dim=0).contiguous()  # [7, n]
        n_keep = torch.zeros(1, dtype=torch.int64, device=self._dev)
        self.ctx.compact_rows(rows, keep, None, n_keep)
        nk = int(n_keep.item())
        removed = torch.nonzero(keep == 0).flatten().cpu().numpy()
        if nk == self.n:
            return removed
        out = torch.empty((7, nk), dtype=torch.float64, device=self._dev)
        self.ctx.compact_rows(rows, keep, out, n_keep)
        self.n = nk
        self.mass, self.pos, self.vel = out[0].contiguous(), out[1:4].contiguous(), out[4:7].contiguous()
        self.acc = torch.empty((3, nk), dtype=torch.float64, device=self._dev)
        self.pot = torch.empty(nk, dtype=torch.float64, device=self._dev)
        self.key = self.key[np.setdiff1d(np.arange(len(self.key)), removed)]
        self._acc_valid = False
        self._alloc_hermite_()
        return removed

    # ---- gravity ----
    def compute_self_gravity(self, want_pot=False):
        """K4 on the current positions -> self.acc (km/s/Myr) and optionally self.pot (kpc km/s/Myr)."""
        self.ctx.self_gravity(self.pos, self.mass, self.parameters._eps2_kpc2, self.G, self.acc,
                              self.pot if want_pot else None)
        self._acc_valid = True
        return self.acc

    def _force_hermite_(self, pos, vel, acc, jerk):
        """K6 at the state (pos, vel) -> acc (km/s/Myr), jerk (km/s/Myr^2)."""
        self.ctx.self_gravity_hermite(pos, vel, self.mass, self.parameters._eps2_kpc2, self.G, KMS_TO_KPC_PER_MYR, acc, jerk)

    def _evolve_hermite_(self, span):
        """`substeps` shared 4th-order Hermite steps over `span` Myr (ph4's scheme without its block steps): force at the
        current state (the BRIDGE kick has just changed the velocities, so the jerk is always re-evaluated), then
        predict -> K6 at the predicted state -> correct.  Leaves the Aarseth step of the last step in self.dt_min."""
        if self.block_steps:
            bs, ss = self.ctx.hermite_block_evolve(self.pos, self.vel, self.mass, self.acc, self.jerk, self.parameters._eps2_kpc2,
                                                   self.G, KMS_TO_KPC_PER_MYR, span, self.eta, self.max_level)
            self.block_step_count += bs
            self.star_step_count += ss
            self._acc_valid = True
            return
        h = span / self.substeps
        self._force_hermite_(self.pos, self.vel, self.acc, self.jerk)
        for _ in range(self.substeps):
            self.ctx.hermite_predict(self.pos, self.vel, self.acc, self.jerk, h, KMS_TO_KPC_PER_MYR, self.pos_p, self.vel_p)
            self._force_hermite_(self.pos_p, self.vel_p, self.acc1, self.jerk1)
            self.ctx.hermite_correct(self.pos, self.vel, self.acc, self.jerk, self.pos_p, self.vel_p, self.acc1, self.jerk1,
                                     h, KMS_TO_KPC_PER_MYR, self.eta, self.dt_min)
        self._acc_valid = True

    def _update_substeps_(self, span):
        """substeps="auto": size the next call from the Aarseth minimum of the last step (one 8-byte read-back)."""
        if self.auto_substeps:
            self.substeps = min(self.max_substeps, self.suggested_substeps(span))

    def _first_substeps_(self, span):
        """substeps="auto", before any step exists: the start-up criterion dt = eta * min |a| / |j| (Aarseth 1985)."""
        import torch
        self._force_hermite_(self.pos, self.vel, self.acc, self.jerk)
        a2, j2 = (self.acc * self.acc).sum(dim=0), (self.jerk * self.jerk).sum(dim=0)
        ok = j2 > 0
        dt = float((self.eta * torch.sqrt(a2[ok] / j2[ok])).min().item()) if bool(ok.any()) else span
        if hasattr(self, "group") and getattr(self, "world", 1) > 1:
            import torch.distributed as dist
            t = torch.tensor([dt], dtype=torch.float64, device=self._dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            dt = float(t.item())
        self.substeps = min(self.max_substeps, int(2 ** max(0, int(np.ceil(np.log2(span / dt))))) if dt > 0 else 1)
        self._auto_started = True

    def suggested_substeps(self, span):
        """Sub-steps (a power of two) that bring span / substeps under the Aarseth step of the last Hermite step."""
        dt = float(self.dt_min.item())
        if not (dt > 0.0) or not np.isfinite(dt):
            return self.substeps
        return int(2 ** max(0, int(np.ceil(np.log2(span / dt)))))

    def _evolve_device_(self, span):
        """The device work of evolve_model: `substeps` kick-drift-kick leapfrog (or Hermite) steps over `span` Myr.  No host
        bookkeeping, no allocation, no synchronisation once the scratch buffers exist — capturable in a CUDA graph."""
        if self.integrator == "hermite":
            return self._evolve_hermite_(span)
        h = span / self.substeps
        for _ in range(self.substeps):
            self.ctx.kick(self.vel, self.acc, 0.5 * h)
            self.ctx.drift(self.pos, self.vel, h, KMS_TO_KPC_PER_MYR)
            self.compute_self_gravity()
            self.ctx.kick(self.vel, self.acc, 0.5 * h)

    def evolve_model(self, t_end, timestep=None):
        """Kick-drift-kick leapfrog under self-gravity from model_time to t_end (the BRIDGE drift,
        oc_nbody.py:56 -> cluster_code.evolve_model)."""
        t_end = float(to_value(t_end, units.Myr))
        span = t_end - self.model_time
        if span <= 0.0:
            return
        if not self._acc_valid and self.integrator != "hermite":
            self.compute_self_gravity()
        if self.auto_substeps and not self._auto_started:
            self._first_substeps_(span)
        self._evolve_device_(span)
        self._update_substeps_(span)
        self.model_time = t_end

    def kick_velocities(self, ax, ay, az, dt_myr):
        """BRIDGE kick from host-side accelerations (generic partner): v += dt * a."""
        import torch
        a = np.stack([np.asarray(to_value(c, units.kms / units.Myr), np.float64) for c in (ax, ay, az)])
        self.ctx.kick(self.vel, torch.from_numpy(np.ascontiguousarray(a)).to(self._dev), dt_myr)

    # ---- driver helpers ----
    def clean_ejections(self, system=None):
        """Remove stars farther than `eject_cut` pc from the median position (oc_code.py:231-246): median select,
        distance test and compaction all on the device."""
        import torch
        if self.eject_cut is None:
            return None
        keep = torch.empty(self.n, dtype=torch.uint8, device=self._dev)
        self.ctx.eject_mask(self.pos, 1000.0, float(self.eject_cut), keep)
        return self._compact_(keep)

    def bound_center_of_mass(self, return_mask=False, iterations=1):
        """Centre of mass (kpc) of the bound subset: E_i = |v_i - v_com|^2/2 + phi_i < 0
        (oc_nbody.py:60-61: particles.bound_subset().center_of_mass()).  K4 with the potential, then one reduction
        kernel; only the 8-double result row comes back to the host.

        iterations = 1: v_com is the centre-of-mass velocity of ALL stars (one pass).  iterations > 1: v_com is re-taken
        from the stars found bound and the test repeated until the bound set stops changing (at most `iterations` passes) —
        once a tidal tail has developed, the escapers inside the ejection cut otherwise drag the frame away from the core
        (AMUSE's own bound_subset works in the core's frame).  The potential is that of all stars either way."""
        import torch
        self.compute_self_gravity(want_pot=True)
        out = torch.empty((1, 8), dtype=torch.float64, device=self._dev)
        want_mask = return_mask or iterations > 1
        mask = torch.empty(self.n, dtype=torch.uint8, device=self._dev) if want_mask else None

        def one_pass(weights):
            # pot is in kpc km/s /Myr; 1/KMS_TO_KPC_PER_MYR turns it into (km/s)^2
            self.ctx.bound_com(self.pos, self.vel, weights, self.pot, 1.0 / KMS_TO_KPC_PER_MYR, out, None, mask)
            return mask
        res = iterate_bound_subset(one_pass, self.pos, self.mass, out, iterations)
        self.n_bound, self.bound_mass = int(res[4]), float(res[3])
        return (res[:3], mask.cpu().numpy().astype(bool)) if return_mask else res[:3]

    def _prepare_capture_(self):
        """Called by Bridge right before it captures a step as a CUDA graph: recompute the current force eagerly (same
        positions, so the same values) so that everything the force call uploads on demand — the work plan — is resident
        and the capture itself contains kernel launches only."""
        if self.integrator == "hermite":
            self._force_hermite_(self.pos, self.vel, self.acc, self.jerk)
        else:
            self.compute_self_gravity()

    def stop(self):
        pass


def iterate_bound_subset(one_pass, pos, mass, out, iterations):
    """The bound-subset iteration above the one-pass reduction kernel.  one_pass(weights) runs ocg_bound_com with `weights`
    in place of the masses — the frame velocity is the weighted mean over ALL stars, so weights = mass * (previous mask)
    makes it the bound set's — fills `out` ([1, 8]) and returns the device mask of ALL stars that pass the energy test.
    Returns the 8-number row (numpy) for the final bound set: COM (3), bound mass, bound count, frame velocity (3)."""
    import torch
    mask = one_pass(mass)
    if iterations <= 1:
        return out.cpu().numpy()[0]
    prev = mask.clone()
    for _ in range(iterations - 1):
        mask = one_pass(mass * prev.to(mass.dtype))
        if torch.equal(mask, prev):
            break
        prev.copy_(mask)
    res = out.cpu().numpy()[0].copy()
    # the kernel's COM row weighs with mass * (previous mask): exact once the set has converged, else redo it for the
    # final mask with the true masses
    w = mass * mask.to(mass.dtype)
    res[3], res[4] = float(w.sum().item()), float(mask.sum().item())
    res[:3] = ((pos * w[None, :]).sum(dim=1) / w.sum()).cpu().numpy()
    return res


class sharded_cluster_code(cluster_code):
    """The cluster code with its stars block-sharded over the ranks of a torch.distributed group (one process per
    GPU): every rank owns stars [a, b) = shard_range(n_total, rank, world) — positions, velocities,
    masses of that block only.  Each self-gravity evaluation all-gathers the positions over NCCL (1.5 MB at
    N = 65 536) and runs K4 for the rank's target range; kicks, drifts and the tidal K3 gather are local.

    mass / pos / vel passed in are the FULL arrays (identical on every rank); the constructor keeps the block."""

    def __init__(self, mass, pos, vel, softening_pc=0.01, substeps=1, eject_cut=None, ctx=None, group=None,
                 integrator="leapfrog", eta=0.14, exchange="peer"):
        """exchange="peer": the gather of the force evaluation runs over peer memory, fused into the tile pack
        (ocg_self_gravity_sharded; ocg_self_gravity_hermite_sharded for positions + velocities: 2 launches per evaluation);
        "nccl": torch.distributed all-gather + copy + the force kernel with a target range."""
        import torch.distributed as dist
        from .distributed import shard_range
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl', not %r" % (exchange,))
        self.exchange = exchange
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        mass = np.asarray(to_value(mass, units.MSun), np.float64)
        self.n_total = mass.shape[0]
        self.a, self.b = shard_range(self.n_total, self.rank, self.world)
        pos = np.asarray(to_value(pos, units.kpc), np.float64)
        vel = np.asarray(to_value(vel, units.kms), np.float64)
        super().__init__(mass[self.a:self.b], pos[:, self.a:self.b], vel[:, self.a:self.b], softening_pc=softening_pc,
                         substeps=substeps, eject_cut=eject_cut, ctx=ctx, integrator=integrator, eta=eta)
        import torch
        self.mass_all = torch.from_numpy(np.ascontiguousarray(mass)).to(self._dev)
        self._acc_all = torch.empty((3, self.n_total), dtype=torch.float64, device=self._dev)
        self._pot_all = torch.empty(self.n_total, dtype=torch.float64, device=self._dev)
        if integrator == "hermite":
            self._jerk_all = torch.empty((3, self.n_total), dtype=torch.float64, device=self._dev)
        self.key = np.arange(self.a, self.b)
        self._peer = False
        if exchange == "peer" and self.parameters._eps2_kpc2 > 0.0 and self.n_total // self.world >= 2048:
            from .distributed import connect_comm
            # first half of the window: 2 x 3 doubles per star of a block for the K4 exchange + 2 x 6 behind them for the
            # Hermite one; the second half belongs to the all-reduce
            connect_comm(self.ctx, group, window_bytes=max(1 << 20, 320 * (self.n_total // self.world + 1)))
            self._peer = True

    def _force_hermite_(self, pos, vel, acc, jerk):
        """K6 for the rank's target block: positions AND velocities are all-gathered (the jerk needs both)."""
        from .distributed import allgather_particles
        if self._peer:
            # one kernel publishes this block's positions and velocities, reads the peers' over NVLink and packs the tiles
            self.ctx.self_gravity_hermite_sharded(pos, vel, self.mass_all, self.parameters._eps2_kpc2, self.G, KMS_TO_KPC_PER_MYR,
                                                  self._acc_all, self._jerk_all)
            acc.copy_(self._acc_all[:, self.a:self.b])
            jerk.copy_(self._jerk_all[:, self.a:self.b])
            return
        pos_all = allgather_particles(pos, self.n_total, self.group).contiguous()
        vel_all = allgather_particles(vel, self.n_total, self.group).contiguous()
        self.ctx.self_gravity_hermite(pos_all, vel_all, self.mass_all, self.parameters._eps2_kpc2, self.G, KMS_TO_KPC_PER_MYR,
                                      self._acc_all, self._jerk_all, tgt_begin=self.a, tgt_end=self.b)
        acc.copy_(self._acc_all[:, self.a:self.b])
        jerk.copy_(self._jerk_all[:, self.a:self.b])

    def _evolve_hermite_(self, span):
        import torch.distributed as dist
        super()._evolve_hermite_(span)
        if dist.is_initialized() and self.world > 1:
            dist.all_reduce(self.dt_min, op=dist.ReduceOp.MIN, group=self.group)

    def compute_self_gravity(self, want_pot=False):
        from .distributed import allgather_particles
        if self._peer:
            # one kernel publishes this block, reads the peers' over NVLink and packs the FP32 tiles; one force kernel
            self.ctx.self_gravity_sharded(self.pos, self.mass_all, self.parameters._eps2_kpc2, self.G, self.acc,
                                          self.pot if want_pot else None)
            self._acc_valid = True
            return self.acc
        pos_all = allgather_particles(self.pos, self.n_total, self.group).contiguous()
        self.ctx.self_gravity(pos_all, self.mass_all, self.parameters._eps2_kpc2, self.G, self._acc_all,
                              self._pot_all if want_pot else None, tgt_begin=self.a, tgt_end=self.b)
        self.acc.copy_(self._acc_all[:, self.a:self.b])
        if want_pot:
            self.pot.copy_(self._pot_all[self.a:self.b])
        self._acc_valid = True
        return self.acc

    def gather_state(self):
        """(pos [3, n_total], vel [3, n_total]) on every rank (for output / verification)."""
        from .distributed import allgather_particles
        return (allgather_particles(self.pos, self.n_total, self.group), allgather_particles(self.vel, self.n_total, self.group))

    def clean_ejections(self, system=None):
        raise NotImplementedError("ejections change the block partition; gather, clean on one rank, and re-shard")

    def bound_center_of_mass(self, return_mask=False):
        raise NotImplementedError("use gather_state() and a single-rank cluster_code for the bound-subset reduction")

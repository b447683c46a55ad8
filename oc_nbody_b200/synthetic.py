"""Seeded synthetic inputs shaped like what the reference reads.

"GIZMO-format" here means the fields gizmo_interface.py:246-251,518-547 touches per species:
``snap[species]['mass']``, ``snap[species].prop('host.distance.principal')``, ``snap['gas']['smooth.length']``,
``snap['star']['id']``, ``snap.snapshot['index'|'time']``.  No file I/O: gizmo_analysis/h5py are absent
and snapshot ingestion is out of scope.
"""
import numpy as np

SEED_SNAPSHOT = 1776  # the reference's default seed (options.py:82,107)
SEED_CLUSTER = 1777


class _Species(dict):
    """dict of arrays with the gizmo_analysis `.prop()` accessor the reference calls."""

    def prop(self, name):
        if name == "host.distance.principal":
            return self["position"]
        if name == "host.velocity.principal":
            return self["velocity"]
        if name == "host.distance.total":
            p = self["position"]
            return np.sqrt((p * p).sum(axis=1))
        if name == "host.distance.principal.cylindrical":
            p = self["position"]
            return np.stack([np.hypot(p[:, 0], p[:, 1]), np.arctan2(p[:, 1], p[:, 0]), p[:, 2]], axis=1)
        raise KeyError(name)


class Snapshot(dict):
    def __init__(self, index, time_gyr):
        super().__init__()
        self.snapshot = {"index": int(index), "time": float(time_gyr)}


def _hernquist_r(rng, n, a, rmax):
    # M(<r)/M = r^2/(r+a)^2, truncated at rmax
    umax = (rmax / (rmax + a)) ** 2
    s = np.sqrt(rng.random(n) * umax)
    return a * s / (1.0 - s)


def _iso(rng, n):
    mu = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    st = np.sqrt(1.0 - mu * mu)
    return np.stack([st * np.cos(ph), st * np.sin(ph), mu], axis=1)


def _disk(rng, n, rd, hz, rmax):
    # exponential surface density: R ~ Gamma(2, rd); sech^2 vertical profile
    R = rng.gamma(2.0, rd, n)
    R = np.where(R > rmax, rng.uniform(0, rmax, n), R)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    z = hz * np.arctanh(rng.uniform(-0.999999, 0.999999, n))
    return np.stack([R * np.cos(ph), R * np.sin(ph), z], axis=1)


def make_snapshot(n_total, seed=SEED_SNAPSHOT, index=577, time_gyr=10.0, rmax=50.0, dtype=np.float64):
    """Milky-Way-like particle set: star:dark:gas = 2:5:3 by number; dark = Hernquist(a=20 kpc) truncated at
    Rmax = 50 kpc (test_options:40); star and gas = exponential disk (R_d = 3 kpc, sech^2 h_z = 0.3 kpc).
    Masses log-uniform within +-10% of 7100 Msun (star, gas) and 35000 Msun (dark) at n_total = 1e8 and scaled by
    1e8/n_total so the total mass is resolution independent."""
    rng = np.random.default_rng(seed)
    n_star = (2 * n_total) // 10
    n_gas = (3 * n_total) // 10
    n_dark = n_total - n_star - n_gas
    scale = 1.0e8 / float(n_total)
    snap = Snapshot(index, time_gyr)

    def masses(n, m0):
        return (m0 * scale * np.exp(rng.uniform(np.log(0.9), np.log(1.1), n))).astype(dtype)

    star = _Species(position=_disk(rng, n_star, 3.0, 0.3, rmax).astype(dtype), mass=masses(n_star, 7100.0),
                    id=np.arange(n_star, dtype=np.int64))
    dark = _Species(position=(_hernquist_r(rng, n_dark, 20.0, rmax)[:, None] * _iso(rng, n_dark)).astype(dtype),
                    mass=masses(n_dark, 35000.0), id=np.arange(n_star, n_star + n_dark, dtype=np.int64))
    gas = _Species(position=_disk(rng, n_gas, 3.0, 0.3, rmax).astype(dtype), mass=masses(n_gas, 7100.0),
                   id=np.arange(n_star + n_dark, n_total, dtype=np.int64))
    gas["smooth.length"] = np.exp(rng.uniform(np.log(1.0), np.log(100.0), n_gas)).astype(dtype)  # pc
    snap["star"], snap["dark"], snap["gas"] = star, dark, gas
    return snap


def advance_snapshot(snap, dt_myr, index=None, vc_kms=220.0):
    """Second snapshot = first one rotated on circular orbits at v_c = 220 km/s for dt (FIRE cadence ~23 Myr)."""
    from .units import KMS_TO_PC_PER_MYR
    out = Snapshot(snap.snapshot["index"] + 1 if index is None else index, snap.snapshot["time"] + dt_myr * 1e-3)
    for sp in ("star", "dark", "gas"):
        src = snap[sp]
        p = src["position"]
        R = np.maximum(np.hypot(p[:, 0], p[:, 1]), 1e-3)
        dphi = vc_kms * KMS_TO_PC_PER_MYR * 1e-3 * dt_myr / R
        c, s = np.cos(dphi), np.sin(dphi)
        q = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1], p[:, 2]], axis=1).astype(p.dtype)
        new = _Species({k: v for k, v in src.items()})
        new["position"] = q
        out[sp] = new
    return out


def make_plummer_cluster(n, a_pc=0.8, m_each=1.0, seed=SEED_CLUSTER, G_pc=None):
    """Equal-mass Plummer sphere in virial equilibrium (a = Rcluster = 0.8 pc, test_options:62).
    Returns pos [3,n] pc, vel [3,n] km/s, mass [n] Msun (King/Kroupa ICs need AMUSE, oc_code.py:197-216)."""
    from .units import G_PC_KMS2
    G = G_PC_KMS2 if G_pc is None else G_pc
    rng = np.random.default_rng(seed)
    M = n * m_each
    u = rng.uniform(1e-10, 1.0, n)
    r = a_pc / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    r = np.minimum(r, 20.0 * a_pc)
    pos = (r[:, None] * _iso(rng, n)).T
    # velocities: rejection sampling of q in g(q) = q^2 (1-q^2)^3.5 (Aarseth, Henon & Wielen 1974)
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.uniform(0.0, 1.0, todo.size)
        y = rng.uniform(0.0, 0.1, todo.size)
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    vesc = np.sqrt(2.0 * G * M / np.sqrt(r * r + a_pc * a_pc))
    vel = ((q * vesc)[:, None] * _iso(rng, n)).T
    mass = np.full(n, m_each)
    pos = pos - pos.mean(axis=1, keepdims=True)
    vel = vel - vel.mean(axis=1, keepdims=True)
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), mass

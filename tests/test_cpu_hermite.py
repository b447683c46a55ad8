"""CPU known-answer tests that pin the oracle's Hermite pieces (the ph4 force loop and step,
oc_code.py:218-229).  ph4 itself is absent from /root/reference and this image (parity unpinned by the reference);
the scheme is restated from Makino & Aarseth (1992) and pinned here by independent checks."""
import numpy as np

import oracle


def _f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def _plummer(n, seed=5):
    rng = np.random.default_rng(seed)
    pos = _f32(rng.normal(0.0, 1.0, (3, n)))
    vel = _f32(rng.normal(0.0, 0.3, (3, n)))
    pos[:, 0] = 0.0
    vel[:, 0] = 0.0  # the first particle is the recentring origin: keep the inputs exactly representable
    mass = _f32(rng.uniform(0.5, 1.5, n))
    return pos, vel, mass


def test_hermite_force_equals_the_closed_form_in_numpy():
    pos, vel, mass = _plummer(150)
    eps2, G, vtl = float(np.float32(1e-3)), 0.7, 1.25
    acc, jerk, pot = oracle.self_gravity_hermite(pos, vel, mass, eps2, G, vtl, want_pot=True)
    d = pos[:, None, :] - pos[:, :, None]          # d[c, i, j] = x_j - x_i
    w = vel[:, None, :] - vel[:, :, None]
    r2 = (d * d).sum(axis=0) + eps2
    np.fill_diagonal(r2, np.inf)
    ri = 1.0 / np.sqrt(r2)
    f = mass[None, :] * ri ** 3
    rv = (d * w).sum(axis=0)
    a_ref = G * (f[None] * d).sum(axis=2)
    j_ref = G * vtl * (f[None] * (w - 3.0 * rv[None] * d * (ri * ri)[None])).sum(axis=2)
    p_ref = -G * (mass[None, :] * ri).sum(axis=1)
    assert np.allclose(acc, a_ref, rtol=1e-12, atol=0)
    assert np.allclose(jerk, j_ref, rtol=1e-10, atol=1e-12 * np.abs(j_ref).max())
    assert np.allclose(pot, p_ref, rtol=1e-12)
    # and the acceleration is the one of the plain force loop, bit for bit
    assert np.array_equal(acc, oracle.self_gravity(pos, mass, eps2, G))


def test_jerk_is_the_time_derivative_of_the_acceleration_along_the_flow():
    # positions on a 2^-6 lattice and velocities on a 2^-4 lattice: x +- h v stays exactly representable in FP32
    rng = np.random.default_rng(11)
    n = 60
    pos = rng.integers(-64, 64, (3, n)) / 64.0
    vel = rng.integers(-8, 8, (3, n)) / 16.0
    pos[:, 0], vel[:, 0] = 0.0, 0.0
    mass = np.ones(n)
    eps2 = 0.25
    h = 2.0 ** -10
    _, jerk = oracle.self_gravity_hermite(pos, vel, mass, eps2, 1.0)
    fd = (oracle.self_gravity(pos + h * vel, mass, eps2, 1.0) - oracle.self_gravity(pos - h * vel, mass, eps2, 1.0)) / (2 * h)
    assert np.max(np.abs(fd - jerk)) <= 1e-5 * np.max(np.abs(jerk))


def test_segments_are_independent_and_shards_are_slices():
    pos, vel, mass = _plummer(90, seed=2)
    seg = np.array([0, 40, 40, 90])
    a, j = oracle.self_gravity_hermite(pos, vel, mass, 1e-3, 1.0, seg_offsets=seg)
    a0, j0 = oracle.self_gravity_hermite(pos[:, :40], vel[:, :40], mass[:40], 1e-3, 1.0)
    assert np.array_equal(a[:, :40], a0) and np.array_equal(j[:, :40], j0)
    a1, j1 = oracle.self_gravity_hermite(pos, vel, mass, 1e-3, 1.0, seg_offsets=seg, t0=30, t1=70)
    assert np.array_equal(a1[:, 30:70], a[:, 30:70]) and np.all(a1[:, :30] == 0) and np.all(j1[:, 70:] == 0)


def _kepler(e=0.5):
    # equal masses, G = 1, total mass 1, semi-major axis 1: period 2 pi; start at apocentre
    m = np.array([0.5, 0.5])
    r_ap = 1.0 + e
    v_ap = np.sqrt((1.0 - e) / (1.0 + e))
    pos = np.array([[-0.5 * r_ap, 0.5 * r_ap], [0.0, 0.0], [0.0, 0.0]])
    vel = np.array([[0.0, 0.0], [-0.5 * v_ap, 0.5 * v_ap], [0.0, 0.0]])
    return pos, vel, m


def _energy(pos, vel, m):
    d = pos[:, 1] - pos[:, 0]
    return 0.5 * (m * (vel * vel).sum(axis=0)).sum() - m[0] * m[1] / np.sqrt((d * d).sum())


def _pure_hermite(pos, vel, m, span, steps):
    """The same predictor / corrector driven by an unrounded FP64 force (the oracle's force loop rounds its inputs to
    FP32 as the GPU does, which would mask the integrator's order)."""
    def force(x, v):
        d, w = x[:, ::-1] - x, v[:, ::-1] - v
        r2 = (d * d).sum(axis=0)
        f = m[::-1] / r2 ** 1.5
        return f * d, f * (w - 3.0 * (d * w).sum(axis=0) / r2 * d)
    h = span / steps
    a, j = force(pos, vel)
    for _ in range(steps):
        xp, vp = oracle.hermite_predict(pos, vel, a, j, h)
        a1, j1 = force(xp, vp)
        pos, vel, dtm = oracle.hermite_correct(xp, vp, a, j, a1, j1, h)
        a, j = a1, j1
    return pos, vel, dtm


def test_hermite_step_is_fourth_order_on_a_kepler_orbit():
    pos, vel, m = _kepler(0.5)
    # global position error after a fraction of an orbit, against a 16x finer run: O(h^4)
    xr, vr, _ = _pure_hermite(pos, vel, m, 2.0, 4096)
    errs = []
    for steps in (64, 128, 256):
        x, v, _ = _pure_hermite(pos, vel, m, 2.0, steps)
        errs.append(np.max(np.abs(x - xr)))
    assert errs[0] < 1e-5
    assert 12.0 < errs[0] / errs[1] < 20.0 and 12.0 < errs[1] / errs[2] < 20.0
    # energy is conserved and the orbit closes after one period
    e0 = _energy(pos, vel, m)
    x, v, _ = _pure_hermite(pos, vel, m, 2 * np.pi, 2048)
    assert abs(_energy(x, v, m) - e0) < 1e-8 * abs(e0)
    assert np.max(np.abs(x - pos)) < 1e-5


def test_aarseth_step_scales_with_the_orbital_time():
    pos, vel, m = _kepler(0.0)  # circular: |a|/|j| = 1/Omega everywhere, a2 = -Omega^2 a, a3 = -Omega^2 j
    _, _, dtm = _pure_hermite(pos, vel, m, 0.5, 64)
    # eta (|a||a2| + |j|^2) / (|j||a3| + |a2|^2) = eta / Omega^2 with Omega = 1
    assert abs(dtm - np.sqrt(0.14)) < 2e-3


def test_oracle_hermite_evolve_conserves_energy_of_a_small_cluster():
    pos, vel, mass = _plummer(64, seed=9)
    eps2 = 4e-2

    def energy(x, v):
        _, _, pot = oracle.self_gravity_hermite(x, v, mass, eps2, 1.0, want_pot=True)
        return 0.5 * (mass * (v * v).sum(axis=0)).sum() + 0.5 * (mass * pot).sum()
    e0 = energy(pos, vel)
    x, v, dtm = oracle.hermite_evolve(pos, vel, mass, eps2, 1.0, 0.125, 64)
    assert abs(energy(x, v) - e0) < 1e-6 * abs(e0)  # FP32-rounded inputs to the force bound this, not the scheme
    assert 0.0 < dtm < 1.0


def test_block_time_steps_kepler_orbit_and_equal_level_limit():
    """ph4's individual block time steps, restated in the oracle (oracle.hermite_block_evolve): an e = 0.6 Kepler orbit
    closes and conserves energy better as eta shrinks; with every star forced onto the smallest step the scheme is the
    shared-step Hermite integrator bit for bit."""
    G = 1.0
    m = np.array([1.0, 1e-3])
    e, a = 0.6, 1.0
    mu = G * m.sum()
    rp, vp = a * (1 - e), np.sqrt(mu * (1 + e) / (a * (1 - e)))
    pos = np.array([[0.0, rp], [0.0, 0.0], [0.0, 0.0]])
    vel = np.array([[0.0, 0.0], [0.0, vp], [0.0, 0.0]])
    vel -= (vel * m).sum(axis=1, keepdims=True) / m.sum()
    pos -= (pos * m).sum(axis=1, keepdims=True) / m.sum()
    period = 2 * np.pi * np.sqrt(a ** 3 / mu)

    def energy(x, v):
        return 0.5 * (m * (v * v).sum(axis=0)).sum() - G * m[0] * m[1] / np.linalg.norm(x[:, 1] - x[:, 0])
    e0 = energy(pos, vel)
    errs = []
    for eta in (0.02, 0.005):
        x, v, acc, jerk, steps, star_steps = oracle.hermite_block_evolve(pos, vel, m, 0.0, G, period, 1.0, eta, 16)
        errs.append((abs(energy(x, v) - e0) / abs(e0), np.linalg.norm((x[:, 1] - x[:, 0]) - (pos[:, 1] - pos[:, 0])), steps))
        a_ref, j_ref = oracle.self_gravity_hermite(x, v, m, 0.0, G, 1.0)
        # acc / jerk left behind are the scheme's own end-of-step force (evaluated at the PREDICTED end state)
        assert np.max(np.abs(acc - a_ref)) <= 1e-3 * np.max(np.abs(a_ref)) and np.max(np.abs(jerk - j_ref)) <= 2e-2 * np.max(np.abs(j_ref))
    assert errs[0][0] < 5e-6 and errs[1][0] < errs[0][0] / 20 and errs[1][1] < errs[0][1] / 8 and errs[1][2] > errs[0][2]
    # steps shrink near pericentre: far fewer block steps than the smallest step taken would need as a shared step
    x, v, _, _, steps, _ = oracle.hermite_block_evolve(pos, vel, m, 0.0, G, period, 1.0, 0.005, 16)
    assert steps < 2 ** 12
    # equal-level limit
    xb, vb, _, _, steps, star_steps = oracle.hermite_block_evolve(pos, vel, m, 0.0, G, period / 8, 1.0, 1e-9, 6)
    xs, vs, _ = oracle.hermite_evolve(pos, vel, m, 0.0, G, period / 8, 64, 1.0, 0.14)
    assert steps == 64 and star_steps == 128 and np.array_equal(xb, xs) and np.array_equal(vb, vs)

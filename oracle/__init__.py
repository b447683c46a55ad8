"""CPU ORACLE for the oceanic gravity hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``oc_nbody_b200``) never does: it has no CPU fallback and
fails loudly when its CUDA library is missing.

PARITY UNPINNED BY THE REFERENCE: gusbeane/oc_nbody has no tests, golden vectors or fixtures for this
path and cannot be imported.  What is pinned, and how, is listed in the header
of ``oracle/ocg_oracle.c`` and in DESIGN.md §Oracle.

The arithmetic lives in ``ocg_oracle.c`` (FP64, OpenMP); this module is the numpy-facing loader plus
the pieces that are clearer in numpy (grid layout, time bracketing, the BRIDGE step order).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ocg_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libocg_oracle.so")
_lib = None

KERNEL_PLUMMER = 0
KERNEL_SPLINE = 1


def _cpu_tag():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """gcc -O3 -march=native -fopenmp -ffp-contract=off; rebuilt when the source or the host CPU changes
    (the prebuilt file travels to the GPU box, whose CPU may differ from the build container's)."""
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    stamp = _OUT + ".stamp"
    tag = "%s|%d" % (_cpu_tag(), int(os.path.getmtime(_SRC)))
    if not force and os.path.exists(_OUT) and os.path.exists(stamp) and open(stamp).read() == tag:
        return _OUT
    base = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-math-errno", "-std=c11"]
    for extra in (["-march=native"], []):
        res = subprocess.run(base + extra + ["-o", _OUT, _SRC, "-lm"], capture_output=True, text=True)
        if res.returncode == 0:
            break
    else:
        raise RuntimeError("building the oracle failed:\n" + res.stderr)
    with open(stamp, "w") as f:
        f.write(tag)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_spline_force.restype = ctypes.c_double
        _lib.oracle_spline_force.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib.oracle_spline_pot.restype = ctypes.c_double
        _lib.oracle_spline_pot.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_set_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    """OpenMP threads of the oracle's loops (bench.py's CPU legs: all host cores, also under torchrun). Returns the count."""
    return int(lib().oracle_set_num_threads(ctypes.c_int(int(n))))


# ------------------------------------------------------------------------------------------------
def recentre(pos, mass, center):
    """(float32)(pos - center) | mass -> [n,4] float32; the rounding the GPU path applies."""
    pos = _c(pos, np.float64).reshape(-1, 3)
    mass = _c(mass, np.float64)
    center = _c(center, np.float64)
    out = np.empty((pos.shape[0], 4), np.float32)
    lib().oracle_recentre(_p(pos), _p(mass), ctypes.c_int64(pos.shape[0]), _p(center), _p(out))
    return out


def spline_force(r, h):
    return lib().oracle_spline_force(float(r), float(h))


def spline_pot(r, h):
    return lib().oracle_spline_pot(float(r), float(h))


def field_direct(src_xyzm, src_soft, tgt_xyzw, kernel, G, want_pot=False):
    """FP64 direct sum on FP32-rounded inputs (gizmo_interface.py:561-566 at theta -> 0).
    Returns acc [3,n_tgt] (and pot [n_tgt])."""
    src = _c(src_xyzm, np.float32).reshape(-1, 4)
    soft = _c(src_soft, np.float32)
    tgt = _c(tgt_xyzw, np.float32).reshape(-1, 4)
    acc = np.zeros((3, tgt.shape[0]), np.float64)
    pot = np.zeros(tgt.shape[0], np.float64) if want_pot else None
    lib().oracle_field_direct(_p(src), _p(soft), ctypes.c_int64(src.shape[0]), _p(tgt), ctypes.c_int64(tgt.shape[0]),
                              ctypes.c_int(kernel), ctypes.c_double(G), _p(acc), _p(pot))
    return (acc, pot) if want_pot else acc


def field_direct_abs(src_xyzm, src_soft, tgt_xyzw, kernel, G):
    """[3, n_tgt] sums of the magnitudes of the pair terms of field_direct (the condition number of each sum)."""
    src = _c(src_xyzm, np.float32).reshape(-1, 4)
    soft = _c(src_soft, np.float32)
    tgt = _c(tgt_xyzw, np.float32).reshape(-1, 4)
    out = np.zeros((3, tgt.shape[0]), np.float64)
    lib().oracle_field_direct_abs(_p(src), _p(soft), ctypes.c_int64(src.shape[0]), _p(tgt), ctypes.c_int64(tgt.shape[0]),
                                  ctypes.c_int(kernel), ctypes.c_double(G), _p(out))
    return out


def field_direct_fast(src_pos, src_mass, src_e2, tgt_pos, G):
    """Vectorisable Plummer sum for the timed CPU baseline (bench.py). FP64 SoA inputs."""
    sp = _c(src_pos, np.float64).reshape(-1, 3)
    sx, sy, sz = (np.ascontiguousarray(sp[:, k]) for k in range(3))
    sm = _c(src_mass, np.float64)
    se = _c(src_e2, np.float64)
    tg = _c(tgt_pos, np.float64).reshape(-1, 3)
    acc = np.zeros((3, tg.shape[0]), np.float64)
    lib().oracle_field_direct_fast(_p(sx), _p(sy), _p(sz), _p(sm), _p(se), ctypes.c_int64(sp.shape[0]), _p(tg),
                                   ctypes.c_int64(tg.shape[0]), ctypes.c_double(G), _p(acc))
    return acc


def frame_subtract(acc, row):
    """acc[:, i] -= acc[:, row] (gizmo_interface.py:566,569-571)."""
    acc = np.array(acc, np.float64, order="C", copy=True)
    lib().oracle_frame_subtract(_p(acc), ctypes.c_int64(acc.shape[1]), ctypes.c_int64(row))
    return acc


def self_gravity(pos, mass, eps2, G, seg_offsets=None, t0=0, t1=None, want_pot=False):
    """Plummer self-gravity per segment (ph4 force loop behind oc_code.py:218-229). pos [3,n] fp64."""
    pos = _c(pos, np.float64)
    n = pos.shape[1]
    mass = _c(mass, np.float64)
    seg = np.array([0, n], np.int64) if seg_offsets is None else _c(seg_offsets, np.int64)
    t1 = n if t1 is None else t1
    acc = np.zeros((3, n), np.float64)
    pot = np.zeros(n, np.float64) if want_pot else None
    lib().oracle_self_gravity(_p(pos), _p(mass), ctypes.c_int64(n), _p(seg), ctypes.c_int32(len(seg) - 1),
                              ctypes.c_double(eps2), ctypes.c_double(G), ctypes.c_int64(t0), ctypes.c_int64(t1),
                              _p(acc), _p(pot))
    return (acc, pot) if want_pot else acc


def self_gravity_hermite(pos, vel, mass, eps2, G, vel_to_len=1.0, seg_offsets=None, t0=0, t1=None, want_pot=False):
    """Acceleration and jerk (and potential) of Plummer self-gravity per segment: the ph4 force loop itself
    (4th-order Hermite, oc_code.py:218-229). pos, vel [3,n] fp64. Returns (acc, jerk[, pot])."""
    pos, vel = _c(pos, np.float64), _c(vel, np.float64)
    n = pos.shape[1]
    mass = _c(mass, np.float64)
    seg = np.array([0, n], np.int64) if seg_offsets is None else _c(seg_offsets, np.int64)
    t1 = n if t1 is None else t1
    acc, jerk = np.zeros((3, n), np.float64), np.zeros((3, n), np.float64)
    pot = np.zeros(n, np.float64) if want_pot else None
    lib().oracle_self_gravity_hermite(_p(pos), _p(vel), _p(mass), ctypes.c_int64(n), _p(seg), ctypes.c_int32(len(seg) - 1),
                                      ctypes.c_double(eps2), ctypes.c_double(G), ctypes.c_double(vel_to_len),
                                      ctypes.c_int64(t0), ctypes.c_int64(t1), _p(acc), _p(jerk), _p(pot))
    return (acc, jerk, pot) if want_pot else (acc, jerk)


def self_gravity_abs(pos, mass, eps2, G, vel=None, vel_to_len=1.0, seg_offsets=None, t0=0, t1=None):
    """Sums of the magnitudes of the pair terms of self_gravity (and, with vel, of the jerk terms of
    self_gravity_hermite): the condition numbers the parity tests weigh cancellation with. [3,n] (, [3,n])."""
    pos = _c(pos, np.float64)
    n = pos.shape[1]
    mass = _c(mass, np.float64)
    vel = None if vel is None else _c(vel, np.float64)
    seg = np.array([0, n], np.int64) if seg_offsets is None else _c(seg_offsets, np.int64)
    t1 = n if t1 is None else t1
    a = np.zeros((3, n), np.float64)
    j = np.zeros((3, n), np.float64) if vel is not None else None
    lib().oracle_self_gravity_abs(_p(pos), _p(vel), _p(mass), ctypes.c_int64(n), _p(seg), ctypes.c_int32(len(seg) - 1),
                                  ctypes.c_double(eps2), ctypes.c_double(G), ctypes.c_double(vel_to_len), ctypes.c_int64(t0),
                                  ctypes.c_int64(t1), _p(a), _p(j))
    return a if vel is None else (a, j)


def hermite_predict(pos, vel, acc, jerk, dt, vel_to_len=1.0):
    """Predictor of the 4th-order Hermite scheme (Makino & Aarseth 1992), every product and sum rounded
    separately in the order the CUDA kernel uses (bit-exact contract)."""
    dt = float(dt)
    c2, c3 = dt * dt * 0.5, dt * dt * dt / 6.0
    dx = (vel * dt + acc * c2) + jerk * c3
    return pos + dx * vel_to_len, vel + (acc * dt + jerk * c2)


def hermite_correct(pos_p, vel_p, acc0, jerk0, acc1, jerk1, dt, vel_to_len=1.0, eta=0.14):
    """Corrector + Aarseth step: returns (pos, vel, dt_min).  Same operation order as the CUDA kernel for pos / vel
    (bit-exact); dt_min = min_i sqrt(eta (|a1||a2'| + |j1|^2) / (|j1||a3| + |a2'|^2)), a2' = a2 + dt a3."""
    dt = float(dt)
    dt2, dt3 = dt * dt, dt * dt * dt
    i2, i3 = 1.0 / dt2, 1.0 / dt3
    d3, d4, d5 = dt3 / 6.0, dt2 * dt2 / 24.0, dt2 * dt3 / 120.0
    da = acc0 - acc1
    a2 = (da * -6.0 - (jerk0 * 4.0 + jerk1 * 2.0) * dt) * i2
    a3 = (da * 12.0 + (jerk0 + jerk1) * (6.0 * dt)) * i3
    vel = (vel_p + a2 * d3) + a3 * d4
    pos = pos_p + (a2 * d4 + a3 * d5) * vel_to_len
    a2e = a2 + dt * a3
    s_a, s_j, s_2, s_3 = ((v * v).sum(axis=0) for v in (acc1, jerk1, a2e, a3))
    num, den = np.sqrt(s_a * s_2) + s_j, np.sqrt(s_j * s_3) + s_2
    ok = (den > 0) & (num > 0)
    dti = np.where(ok, np.sqrt(eta * num / np.where(ok, den, 1.0)), np.inf)
    return pos, vel, float(dti.min()) if dti.size else float("inf")


def hermite_evolve(pos, vel, mass, eps2, G, span, substeps, vel_to_len=1.0, eta=0.14):
    """`substeps` shared Hermite steps over `span` (what cluster_code(integrator="hermite").evolve_model does):
    force at the current state, then predict - evaluate - correct per step. Returns (pos, vel, dt_min)."""
    h = span / substeps
    acc, jerk = self_gravity_hermite(pos, vel, mass, eps2, G, vel_to_len)
    dt_min = float("inf")
    for _ in range(substeps):
        xp, vp = hermite_predict(pos, vel, acc, jerk, h, vel_to_len)
        a1, j1 = self_gravity_hermite(xp, vp, mass, eps2, G, vel_to_len)
        pos, vel, dt_min = hermite_correct(xp, vp, acc, jerk, a1, j1, h, vel_to_len, eta)
        acc, jerk = a1, j1
    return pos, vel, dt_min


def hermite_block_evolve(pos, vel, mass, eps2, G, span, vel_to_len=1.0, eta=0.14, max_level=12, force=None):
    """ph4's individual block time steps over `span` (AMUSE ph4 behind oc_code.py:218-229), restated from Makino & Aarseth
    (1992) / Aarseth (1985): per-star power-of-two steps span / 2^k (k <= max_level), t_next = min (t_i + dt_i), every star
    predicted to t_next, force on the active stars from all predicted stars, corrector over the star's own step, next step
    from the Aarseth criterion (halve as needed; double once, onto commensurate times only).  Times are integer ticks
    (span = 2^20), exactly as ocg_hermite_block_evolve keeps them.  `force(pos, vel) -> (acc, jerk)` defaults to the FP64
    oracle sum.  Returns (pos, vel, acc, jerk, block_steps, star_steps)."""
    if force is None:
        def force(x, v):
            return self_gravity_hermite(x, v, mass, eps2, G, vel_to_len)
    pos, vel = np.array(pos, np.float64), np.array(vel, np.float64)
    n = pos.shape[1]
    SPAN = 1 << 20
    tick = span / SPAN
    min_ticks = SPAN >> int(max_level)
    acc, jerk = force(pos, vel)
    a2, j2 = (acc * acc).sum(axis=0), (jerk * jerk).sum(axis=0)
    dt_tick = np.full(n, SPAN, np.int64)
    want = np.where(j2 > 0, eta * np.sqrt(a2 / np.where(j2 > 0, j2, 1.0)), np.inf)
    for i in range(n):
        while dt_tick[i] > min_ticks and dt_tick[i] * tick > want[i]:
            dt_tick[i] >>= 1
    t_tick = np.zeros(n, np.int64)
    steps = star_steps = 0
    while True:
        tn = int((t_tick + dt_tick).min())
        act = np.nonzero(t_tick + dt_tick == tn)[0]
        d = (tn - t_tick) * tick
        c2, c3 = d * d * 0.5, d * d * d / 6.0
        dx = (vel * d + acc * c2) + jerk * c3
        xp, vp = pos + dx * vel_to_len, vel + (acc * d + jerk * c2)
        a1, j1 = force(xp, vp)
        for i in act:
            dt = dt_tick[i] * tick
            dt2, dt3 = dt * dt, dt * dt * dt
            da = acc[:, i] - a1[:, i]
            A2 = (da * -6.0 - (jerk[:, i] * 4.0 + j1[:, i] * 2.0) * dt) * (1.0 / dt2)
            A3 = (da * 12.0 + (jerk[:, i] + j1[:, i]) * (6.0 * dt)) * (1.0 / dt3)
            vel[:, i] = (vp[:, i] + A2 * (dt3 / 6.0)) + A3 * (dt2 * dt2 / 24.0)
            pos[:, i] = xp[:, i] + (A2 * (dt2 * dt2 / 24.0) + A3 * (dt2 * dt3 / 120.0)) * vel_to_len
            acc[:, i], jerk[:, i] = a1[:, i], j1[:, i]
            a2e = A2 + dt * A3
            s_a, s_j, s_2, s_3 = (a1[:, i] ** 2).sum(), (j1[:, i] ** 2).sum(), (a2e ** 2).sum(), (A3 ** 2).sum()
            num, den = np.sqrt(s_a * s_2) + s_j, np.sqrt(s_j * s_3) + s_2
            nxt = int(dt_tick[i])
            if den > 0 and num > 0:
                w = np.sqrt(eta * num / den)
                if w < dt:
                    while nxt > min_ticks and nxt * tick > w:
                        nxt >>= 1
                elif w >= 2.0 * dt and 2 * nxt <= SPAN and tn % (2 * nxt) == 0:
                    nxt <<= 1
            t_tick[i], dt_tick[i] = tn, nxt
        steps += 1
        star_steps += len(act)
        if tn == SPAN:
            break
    return pos, vel, acc, jerk, steps, star_steps


def pack_planes(acc, pot=None):
    acc = _c(acc, np.float64)
    n = acc.shape[1]
    pot = _c(pot, np.float64)
    rec = np.empty((n, 4), np.float32)
    lib().oracle_pack_planes(_p(acc), _p(pot), ctypes.c_int64(n), _p(rec))
    return rec


def time_blend(rec_a, rec_b, wb, want_pot=False):
    ra = _c(rec_a, np.float32).reshape(-1, 4)
    rb = None if rec_b is None else _c(rec_b, np.float32).reshape(-1, 4)
    n = ra.shape[0]
    acc = np.empty((3, n), np.float64)
    pot = np.empty(n, np.float64) if want_pot else None
    lib().oracle_time_blend(_p(ra), _p(rb), ctypes.c_double(wb), ctypes.c_int64(n), _p(acc), _p(pot))
    return (acc, pot) if want_pot else acc


def grid_interp(nodes, origin, rec_a, rec_b, wb, sx, sy, sz, star_cluster=None, want_pot=False, want_cell=False):
    """Trilinear + linear-in-time evaluation at star positions (get_gravity_at_point,
    gizmo_interface.py:677-717, with the north_star's interpolation). nodes = (xg, yg, zg) fp64."""
    xg, yg, zg = (_c(a, np.float64) for a in nodes)
    origin = _c(origin, np.float64).reshape(-1, 3)
    ra = _c(rec_a, np.float32)
    rb = None if rec_b is None else _c(rec_b, np.float32)
    sx, sy, sz = (_c(a, np.float64) for a in (sx, sy, sz))
    scl = _c(star_cluster, np.int32)
    n = sx.shape[0]
    nn = np.array([len(xg), len(yg), len(zg)], np.int32)
    acc = np.empty((3, n), np.float64)
    pot = np.empty(n, np.float64) if want_pot else None
    cell = np.empty((3, n), np.int32) if want_cell else None
    lib().oracle_grid_interp(_p(nn), ctypes.c_int32(origin.shape[0]), _p(xg), _p(yg), _p(zg), _p(origin), _p(ra), _p(rb),
                             ctypes.c_double(wb), _p(sx), _p(sy), _p(sz), _p(scl), ctypes.c_int64(n), _p(acc), _p(pot),
                             _p(cell))
    out = [acc]
    if want_pot:
        out.append(pot)
    if want_cell:
        out.append(cell)
    return out[0] if len(out) == 1 else tuple(out)


def grid_interp_multi(nodes, origin, recs, weights, sx, sy, sz, star_cluster=None, want_pot=False):
    """Same evaluation with 1..4 record planes blended in time (cubic B-spline coefficients: the reference's own
    splrep/splev time interpolation, gizmo_interface.py:587-620, collapsed over the shared knot vector)."""
    xg, yg, zg = (_c(a, np.float64) for a in nodes)
    origin = _c(origin, np.float64).reshape(-1, 3)
    recs = [_c(r, np.float32) for r in recs]
    ptrs = (ctypes.c_void_p * len(recs))(*[r.ctypes.data for r in recs])
    w = _c(weights, np.float64)
    sx, sy, sz = (_c(a, np.float64) for a in (sx, sy, sz))
    scl = _c(star_cluster, np.int32)
    n = sx.shape[0]
    nn = np.array([len(xg), len(yg), len(zg)], np.int32)
    acc = np.empty((3, n), np.float64)
    pot = np.empty(n, np.float64) if want_pot else None
    lib().oracle_grid_interp_multi(_p(nn), ctypes.c_int32(origin.shape[0]), _p(xg), _p(yg), _p(zg), _p(origin), ptrs, _p(w),
                                   ctypes.c_int32(len(recs)), _p(sx), _p(sy), _p(sz), _p(scl), ctypes.c_int64(n), _p(acc),
                                   _p(pot), None)
    return (acc, pot) if want_pot else acc


def grid_interp_nested(nodes, fine_nodes, origin, recs, recs_fine, weights, sx, sy, sz, star_cluster=None,
                       want_pot=False, want_tensor=False, want_level=False, want_cell=False):
    """Two-level evaluation (the reference's nested fine grid, grid_cartesian.py:34-53,71-91): stars inside the
    closed fine box use the fine lattice, the others the coarse one; optional tidal tensor T[3*i+j] = d a_j/d x_i
    (gizmo_interface.py:719-756). fine_nodes None: single level. Returns a dict."""
    xg, yg, zg = (_c(a, np.float64) for a in nodes)
    nn = np.array([len(xg), len(yg), len(zg)], np.int32)
    origin = _c(origin, np.float64).reshape(-1, 3)
    recs = [_c(r, np.float32) for r in recs]
    ptrs = (ctypes.c_void_p * len(recs))(*[r.ctypes.data for r in recs])
    if fine_nodes is not None:
        fx, fy, fz = (_c(a, np.float64) for a in fine_nodes)
        nn2 = np.array([len(fx), len(fy), len(fz)], np.int32)
        recs2 = [_c(r, np.float32) for r in recs_fine]
        ptrs2 = (ctypes.c_void_p * len(recs2))(*[r.ctypes.data for r in recs2])
    else:
        fx = fy = fz = nn2 = ptrs2 = None
    w = _c(weights, np.float64)
    sx, sy, sz = (_c(a, np.float64) for a in (sx, sy, sz))
    scl = _c(star_cluster, np.int32)
    n = sx.shape[0]
    out = dict(acc=np.empty((3, n), np.float64), pot=np.empty(n, np.float64) if want_pot else None,
               tensor=np.empty((9, n), np.float64) if want_tensor else None,
               level=np.empty(n, np.int32) if want_level else None, cell=np.empty((3, n), np.int32) if want_cell else None)
    lib().oracle_grid_interp_nested(_p(nn), _p(xg), _p(yg), _p(zg), _p(nn2), _p(fx), _p(fy), _p(fz), _p(origin), ptrs, ptrs2,
                                    _p(w), ctypes.c_int32(len(recs)), _p(sx), _p(sy), _p(sz), _p(scl), ctypes.c_int64(n),
                                    _p(out["acc"]), _p(out["pot"]), _p(out["tensor"]), _p(out["level"]), _p(out["cell"]))
    return out


def rbf_monomial_powers(order):
    """Exponents of the monomials of total degree <= order in 3-D (56 for the reference's order 5)."""
    return [(a, b, d - a - b) for d in range(order + 1) for a in range(d, -1, -1) for b in range(d - a, -1, -1)]


def rbf_phs(r, phs):
    """rbf.basis.phs<k> up to sign (options.py:178-202): r^k for odd k, r^k log r for even k (0 at r = 0)."""
    r = np.asarray(r, np.float64)
    if phs % 2:
        return r ** phs
    return np.where(r > 0, r ** phs * np.log(np.where(r > 0, r, 1.0)), 0.0)


def rbf_phs_dr_over_r(r, phs):
    """phi'(r) / r."""
    r = np.asarray(r, np.float64)
    if phs % 2:
        return phs * np.where(r > 0, r ** (phs - 2), 0.0) if phs == 1 else phs * r ** (phs - 2)
    return np.where(r > 0, r ** (phs - 2) * (phs * np.log(np.where(r > 0, r, 1.0)) + 1.0), 0.0)


def rbf_interp_points(pts, fields, sx, sy, sz, h, nclose=150, order=5, phs=3, want_tensor=False, want_neighbors=False):
    """kNN(nclose) + RBF-PHS interpolation of `fields` [n_comp, n_pts] given on an arbitrary point list `pts` [n_pts, 3]
    (what gizmo_interface.py:651-717 does on grid.evolved_grid, nested or not).  Brute-force neighbour search ordered by
    (distance^2, point index); the saddle-point system [[K, P], [P', 0]] in coordinates shifted to the star and scaled
    by `h` (the interpolant is invariant under both), solved with LAPACK in FP64."""
    pts = np.asarray(pts, np.float64)
    fields = np.asarray(fields, np.float64)
    pw = np.array(rbf_monomial_powers(order))
    nm = len(pw)
    sx, sy, sz = (np.atleast_1d(np.asarray(a, np.float64)) for a in (sx, sy, sz))
    n = sx.shape[0]
    out = np.empty((fields.shape[0], n))
    tensor = np.empty((3, fields.shape[0], n)) if want_tensor else None
    nbrs = np.empty((nclose, n), np.int64)
    for s in range(n):
        p = np.array([sx[s], sy[s], sz[s]])
        d = pts - p
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        ids = np.lexsort((np.arange(len(d2)), d2))[:nclose]
        nbrs[:, s] = ids
        y = (pts[ids] - p) / h
        r = np.sqrt(((y[:, None, :] - y[None, :, :]) ** 2).sum(axis=-1))
        P = np.prod(y[:, None, :] ** pw[None, :, :], axis=-1)
        A = np.block([[rbf_phs(r, phs), P], [P.T, np.zeros((nm, nm))]])
        rhs = np.concatenate([fields[:, ids].T, np.zeros((nm, fields.shape[0]))])
        sol = np.linalg.solve(A, rhs)            # [nclose + nm, n_comp]: weights, then polynomial coefficients
        r0 = np.sqrt((y * y).sum(axis=1))
        out[:, s] = rbf_phs(r0, phs) @ sol[:nclose] + sol[nclose]    # every monomial but the constant vanishes at the star
        if want_tensor:
            for a in range(3):
                # d/dx_a at the star: -(phi'(|y|)/|y|) y_a for the radial part, the coefficient of x_a for the polynomial
                g = -rbf_phs_dr_over_r(r0, phs) * y[:, a]
                lin = [k for k in range(nm) if pw[k].sum() == 1 and pw[k][a] == 1]
                tensor[a, :, s] = (g @ sol[:nclose] + (sol[nclose + lin[0]] if lin else 0.0)) / h
    res = dict(out=out)
    if want_tensor:
        res["tensor"] = tensor
    if want_neighbors:
        res["neighbors"] = nbrs
    return res


def rbf_interp(nodes, origin, fields, sx, sy, sz, nclose=150, order=5, phs=3, include_origin=True, want_tensor=False,
               want_neighbors=False):
    """get_gravity_at_point as the reference evaluates it (gizmo_interface.py:651-717): per star the `nclose` nearest
    points of the evolved grid (lattice nodes in C order + the appended origin row; cKDTree.query) and
    rbf.interpolate.RBFInterpolant(points, values, basis=phs<phs>, order=order) evaluated at the star.  The `rbf`
    package is absent here; scipy.interpolate.RBFInterpolator (same author, kernel='cubic', degree=order) is the
    same interpolant and pins this function in tests/test_cpu_rbf.py.
    fields [n_comp, n_node]; returns dict(out [n_comp, n], tensor [3, n_comp, n], neighbors)."""
    xg, yg, zg = (np.asarray(a, np.float64) for a in nodes)
    o = np.asarray(origin, np.float64).reshape(3)
    gx, gy, gz = np.meshgrid(xg + o[0], yg + o[1], zg + o[2], indexing="ij")
    pts = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    pts = np.concatenate([pts, o[None]]) if include_origin else pts
    h = max(xg[1] - xg[0], yg[1] - yg[0], zg[1] - zg[0])
    return rbf_interp_points(pts, fields, sx, sy, sz, h, nclose, order, phs, want_tensor, want_neighbors)


def layout_nested(planes, n_coarse, keep_index, hole_index, hole_points, fine_nodes, fine_row0):
    """Point-list planes [P, 4, Npoints] of a nested grid (kept coarse | fine | origin, grid_cartesian.py:71-91) ->
    (coarse records [P, n_coarse+1, 4], fine records [P, n_fine+1, 4]) FP32, the dropped coarse points filled by
    trilinear interpolation of the fine lattice at their positions (same rule as gizmo_field._layout_planes_)."""
    planes = np.asarray(planes, np.float64)
    P = planes.shape[0]
    nf = planes.shape[2] - 1 - fine_row0
    coarse = np.zeros((P, n_coarse + 1, 4), np.float32)
    fine = np.empty((P, nf + 1, 4), np.float32)
    for i in range(P):
        fine[i] = pack_planes(planes[i, :3, fine_row0:], planes[i, 3, fine_row0:])
        coarse[i, keep_index] = pack_planes(planes[i, :3, :fine_row0], planes[i, 3, :fine_row0])
        coarse[i, n_coarse] = pack_planes(planes[i, :3, -1:], planes[i, 3, -1:])[0]
        if len(hole_index):
            acc, pot = grid_interp(fine_nodes, np.zeros(3), fine[i], None, 0.0, hole_points[:, 0], hole_points[:, 1],
                                   hole_points[:, 2], want_pot=True)
            coarse[i, hole_index] = pack_planes(acc, pot)
    return coarse, fine


def bound_com(pos, vel, mass, pot_v2, iterations=1):
    """Bound subset and its centre of mass (oc_nbody.py:60-61, particles.bound_subset().center_of_mass()):
    E_i = |v_i - v_com|^2 / 2 + phi_i < 0 with v_com the velocity of the cluster's centre of mass; everything
    counts when nothing is bound.  iterations > 1: v_com is re-taken from the stars found bound and the test repeated
    until the set stops changing.  pos, vel [3, n]; pot_v2 [n] in velocity^2 units. Returns (com[3], mask[n], v_com[3])."""
    pos, vel, mass, pot_v2 = (np.asarray(a, np.float64) for a in (pos, vel, mass, pot_v2))
    w = mass.copy()
    bound = None
    for _ in range(max(1, iterations)):
        vcom = (vel * w).sum(axis=1) / w.sum()
        dv = vel - vcom[:, None]
        new = 0.5 * (dv * dv).sum(axis=0) + pot_v2 < 0.0
        if not new.any():
            new[:] = True
        if bound is not None and np.array_equal(new, bound):
            break
        bound = new
        w = mass * bound
    return (pos[:, bound] * mass[bound]).sum(axis=1) / mass[bound].sum(), bound, vcom


def eject_keep(pos, len_scale, cut):
    """clean_ejections (oc_code.py:231-246): stars farther than `cut` from the per-axis median position are
    removed. pos [3, n]. Returns (keep[n] bool, median[3])."""
    pos = np.asarray(pos, np.float64)
    med = np.median(pos, axis=1)
    d = (pos - med[:, None]) * len_scale
    return ~(np.sqrt((d * d).sum(axis=0)) > cut), med


def kick(vel, acc, dt):
    vel = np.array(vel, np.float64, order="C", copy=True)
    acc = _c(acc, np.float64)
    lib().oracle_kick(_p(vel), _p(acc), ctypes.c_int64(vel.shape[1]), ctypes.c_double(dt))
    return vel


def drift(pos, vel, dt, vel_to_len=1.0):
    pos = np.array(pos, np.float64, order="C", copy=True)
    vel = _c(vel, np.float64)
    lib().oracle_drift(_p(pos), _p(vel), ctypes.c_int64(pos.shape[1]), ctypes.c_double(dt), ctypes.c_double(vel_to_len))
    return pos


# ---------------------------------------------------------------------------- numpy restatements ----
def grid_layout(x_size, y_size, z_size, resolution):
    """Restatement of grid_cartesian.py:16-32,59-69: n = int(L/res) nodes of np.linspace(-L, L, n) per
    axis, points in C order (x outer, z inner), one origin row appended.
    Returns (x_grid, y_grid, z_grid, init_grid[nx*ny*nz+1, 3])."""
    n = [int(L / resolution) for L in (x_size, y_size, z_size)]
    axes = [np.linspace(-L, L, num=k) for L, k in zip((x_size, y_size, z_size), n)]
    rows = []
    for xi in axes[0]:
        for yj in axes[1]:
            for zk in axes[2]:
                rows.append((xi, yj, zk))
    rows.append((0.0, 0.0, 0.0))
    return axes[0], axes[1], axes[2], np.array(rows, np.float64).reshape(-1, 3)


def time_bracket(times, t):
    """Linear-in-time bracket for snapshot times (north_star's substitute for the per-node cubic splines of
    gizmo_interface.py:587-620): index i with times[i] <= t < times[i+1] (clamped), weight of snapshot i+1."""
    times = np.asarray(times, np.float64)
    if len(times) == 1:
        return 0, 0, 0.0
    i = int(np.searchsorted(times, t, side="right")) - 1
    i = min(max(i, 0), len(times) - 2)
    w = (t - times[i]) / (times[i + 1] - times[i])
    w = min(max(w, 0.0), 1.0)
    return i, i + 1, float(w)


def bridge_step(pos_pc, vel_kms, mass, eps2_pc2, G_pc, dt_myr, tidal_acc, kms_myr_to_pc):
    """One BRIDGE step in the order amuse.couple.bridge runs it for oc_nbody.py:49-56:
    K(dt/2) by the field code, D(dt) of the cluster under self-gravity, K(dt/2).

    The drift is a kick-drift-kick leapfrog (north_star's substitute for ph4's Hermite).
    pos [3,n] pc, vel [3,n] km/s; tidal_acc(pos_pc) -> [3,n] km/s/Myr; self-gravity in (km/s)^2/pc is
    converted to km/s/Myr by ``kms_myr_to_pc`` (pc travelled per Myr at 1 km/s)."""
    v = kick(vel_kms, tidal_acc(pos_pc), 0.5 * dt_myr)
    a = self_gravity(pos_pc, mass, eps2_pc2, G_pc) * kms_myr_to_pc
    v = kick(v, a, 0.5 * dt_myr)
    x = drift(pos_pc, v, dt_myr, kms_myr_to_pc)
    a = self_gravity(x, mass, eps2_pc2, G_pc) * kms_myr_to_pc
    v = kick(v, a, 0.5 * dt_myr)
    v = kick(v, tidal_acc(x), 0.5 * dt_myr)
    return x, v

"""Shared helpers for the parity tests."""
import numpy as np

# BASELINE.json north_star: per-component relative acceleration error <= 1e-5 (fp32-source, fp64-accumulate)
TOL = 1e-5


# relative error of ONE pair term as the FP32 kernels deliver it, in units of the term itself: the pair arithmetic (d: 0.5
# ulp; r^2 by three FMAs: 1.5, times 1.5 for the -3/2 power; r^6 by two multiplies: 1; MUFU.RSQ: 2-3; m*y3: 0.5 => ~8 ulp)
# plus the rounding of the FP32 running sum the term is added to inside one accumulation run (a few ulp of the run's
# terms): 16 ulp of 2^-24
EPS_PAIR = 1e-6


def rel_err(a_gpu, a_ref, floor=1e-3, abs_sum=None, eps_pair=EPS_PAIR):
    """Per-component acceleration error, SURVEY §8(d)'s parity metric. a_* are [3, n].

    err = max_{t,c} |a_gpu[c,t] - a_ref[c,t]| / max(|a_ref[c,t]|, floor * ||a_ref[:,t]||_2)

    floor = 1e-3 (the default, THE GATE): every component is held to 1e-5 of its own size, down to components a
    thousand times smaller than the vector.

    abs_sum ([3, n], optional): sum over the sources of the MAGNITUDES of the pair terms of that component (the oracle's
    *_abs functions).  north_star prescribes FP32 pair arithmetic with FP64 accumulation: each pair term is good to a few
    ulp of ITSELF, so the sum carries an error up to eps_pair * abs_sum however it is accumulated.  Where terms cancel
    (a star between two close neighbours; a target inside a cloud of weakly softened particles) that bound exceeds
    1e-5 of the net component, and no FP32-pair kernel can do better.  With abs_sum given, the denominator becomes
    max(|a_ref|, floor ||a_ref||, (eps_pair / TOL) * abs_sum): a result passes (err <= TOL) when it meets the strict
    bound OR the backward-error bound of its own sum — used for the cluster self-gravity kernels (close pairs are the
    rule there) and for the synthetic K1 cases that put targets inside the source cloud.  The galaxy-field cases
    (configs[0], [1]) are gated without it.

    floor = 1 ("norm" metric, `rel_err_norm`): error relative to the vector norm; only for quantities that are not sums of
    pair terms (trajectories after several steps)."""
    a_gpu, a_ref = np.asarray(a_gpu, np.float64), np.asarray(a_ref, np.float64)
    norm = np.sqrt((a_ref * a_ref).sum(axis=0))
    den = np.maximum(np.abs(a_ref), floor * norm[None, :])
    if abs_sum is not None:
        den = np.maximum(den, (eps_pair / TOL) * np.asarray(abs_sum, np.float64))
    den = np.where(den > 0, den, 1.0)
    return float(np.max(np.abs(a_gpu - a_ref) / den))


def rel_err_norm(a_gpu, a_ref):
    """Error relative to the vector norm (floor = 1); see rel_err for when this is the honest gate."""
    return rel_err(a_gpu, a_ref, floor=1.0)


def rel_err_strict(a_gpu, a_ref):
    """Alias of the gate (floor 1e-3 ||a||)."""
    return rel_err(a_gpu, a_ref, floor=1e-3)


def rel_err_scalar(p_gpu, p_ref):
    p_gpu, p_ref = np.asarray(p_gpu, np.float64), np.asarray(p_ref, np.float64)
    den = np.where(np.abs(p_ref) > 0, np.abs(p_ref), 1.0)
    return float(np.max(np.abs(p_gpu - p_ref) / den))


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def random_sources(rng, n, box=5.0, center=(0.0, 0.0, 0.0), soft_lo=0.004, soft_hi=0.3):
    """fp32 (x,y,z,m) + soft, clustered around `center` with a wide halo."""
    pos = rng.normal(0.0, box, (n, 3)) + np.asarray(center)
    m = np.exp(rng.uniform(np.log(1e3), np.log(1e5), n))
    soft = np.exp(rng.uniform(np.log(soft_lo), np.log(soft_hi), n))
    return np.concatenate([pos, m[:, None]], axis=1).astype(np.float32), soft.astype(np.float32)


def grid_targets(n, half=0.6):
    ax = np.linspace(-half, half, n)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)
    g = np.concatenate([g, np.zeros((1, 3))])
    return np.concatenate([g, np.zeros((g.shape[0], 1))], axis=1).astype(np.float32)

"""Shared helpers for the parity tests."""
import numpy as np

# BASELINE.json north_star: per-component relative acceleration error <= 1e-5 (fp32-source, fp64-accumulate)
TOL = 1e-5


def rel_err(a_gpu, a_ref):
    """SURVEY §8(d) parity metric. a_* are [3, n]: |d| / max(|a_ref_c|, 1e-3 * ||a_ref||_2) per component."""
    a_gpu, a_ref = np.asarray(a_gpu, np.float64), np.asarray(a_ref, np.float64)
    norm = np.sqrt((a_ref * a_ref).sum(axis=0))
    den = np.maximum(np.abs(a_ref), 1e-3 * norm[None, :])
    den = np.where(den > 0, den, 1.0)
    return float(np.max(np.abs(a_gpu - a_ref) / den))


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

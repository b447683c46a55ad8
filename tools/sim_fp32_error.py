"""CPU emulation of K1's FP32 arithmetic (mass-folded, target-paired tile loop) to attribute its error budget.

Runs here (no GPU): every FP32 operation of direct_kernel.cuh::tile_tpair<MF=2> is replayed in numpy with the same
roundings (FMA = one rounding of the exact product-sum; MUFU.RSQ modelled as the correctly rounded value times a
random +-1 ulp factor), for a sample of the 64^3 lattice against a synthetic snapshot, and compared with the FP64
sum on the same FP32-rounded inputs.  Switches isolate the terms of the error budget:

  fold      : sources per FP32 accumulation run before the fold into FP64 (the kernel: 512)
  exact_acc : accumulate the FP32 pair terms in FP64 (isolates per-pair error from accumulation error)
  rsq_noise : relative error model of the approximate reciprocal square root (ulps)

python tools/sim_fp32_error.py [n_src] [n_tgt]  ->  profiles/r02_error_budget_sim.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import CENTER, make_sources, make_targets  # noqa: E402
from util import rel_err  # noqa: E402

F = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def classify(src, e2s, box_lo, box_hi, sc, cap):
    d = np.maximum(np.maximum(box_lo - src[:, :3], src[:, :3] - box_hi), 0).astype(F) * F(sc)
    d2 = (d * d).sum(axis=1)
    D2 = 0.0
    lim = 0.25
    for _ in range(12):
        if (d2 < lim).sum() <= cap:
            D2 = lim
            break
        lim *= 0.25
    return d2 < D2, D2


def run(s32, soft, t32, fold=512, exact_acc=False, rsq_ulp=1.0, near_cap=None, seed=0, mf=True, tchunk=32):
    rng = np.random.default_rng(seed)
    n_s, n_t = s32.shape[0], t32.shape[0]
    lo, hi = t32[:, :3].min(axis=0), t32[:, :3].max(axis=0)
    ext = float((hi - lo).max())
    e = np.frexp(F(ext))[1]
    sc = float(np.ldexp(1.0, -int(e)))
    cap = max(4096, n_s // 512) if near_cap is None else near_cap
    near, D2 = classify(s32, None, lo, hi, sc, cap)
    fs, ns = s32[~near], s32[near]
    fsoft, nsoft = soft[~near], soft[near]
    # ---- tile records
    m = fs[:, 3]
    M0 = float(np.ldexp(1.0, int(np.ceil(np.log2(m.max())))))
    hs = (fsoft * F(sc)).astype(F)
    e2 = (hs * hs).astype(F)
    if mf:
        w = (1.0 / np.sqrt((m / F(M0)).astype(np.float64))).astype(F)
        X = [((fs[:, c] * F(sc)).astype(F) * w).astype(F) for c in range(3)]
        E = ((e2 * w).astype(F) * w).astype(F)
    else:
        w = m.copy()
        X = [(fs[:, c] * F(sc)).astype(F) for c in range(3)]
        E = e2
    nf = fs.shape[0]
    pad = (-nf) % fold
    out = np.zeros((3, n_t))
    for t0 in range(0, n_t, tchunk):
        T = t32[t0:t0 + tchunk]
        nt = T.shape[0]
        ntc = [(-(T[:, c] * F(sc))).astype(F) for c in range(3)]  # negated scaled target coordinate
        d = []
        for c in range(3):
            if mf:
                d.append(fma32(ntc[c][None, :], w[:, None], X[c][:, None]))
            else:
                d.append((ntc[c][None, :] + X[c][:, None]).astype(F))
        r2 = fma32(d[0], d[0], np.broadcast_to(E[:, None], d[0].shape))
        r2 = fma32(d[1], d[1], r2)
        r2 = fma32(d[2], d[2], r2)
        r6 = ((r2 * r2).astype(F) * r2).astype(F)
        y3 = (1.0 / np.sqrt(r6.astype(np.float64)))
        if rsq_ulp > 0:
            y3 = y3 * (1.0 + rng.uniform(-1, 1, y3.shape) * rsq_ulp * 2.0 ** -24)
        y3 = y3.astype(F)
        if not mf:
            y3 = (y3 * w[:, None]).astype(F)
        for c in range(3):
            if exact_acc:
                out[c, t0:t0 + nt] = (d[c].astype(np.float64) * y3.astype(np.float64)).sum(axis=0)
                continue
            dd = np.concatenate([d[c], np.zeros((pad, nt), F)]).reshape(-1, fold, nt)
            yy = np.concatenate([y3, np.zeros((pad, nt), F)]).reshape(-1, fold, nt)
            acc = np.zeros((dd.shape[0], nt), F)
            for k in range(fold):
                acc = fma32(dd[:, k], yy[:, k], acc)
            out[c, t0:t0 + nt] = acc.astype(np.float64).sum(axis=0)
    out *= (M0 if mf else 1.0) * sc * sc
    # ---- near set in FP64 (as near_sum_kernel)
    if ns.shape[0]:
        dn = ns[None, :, :3].astype(np.float64) - t32[:, None, :3].astype(np.float64)
        q2 = (dn * dn).sum(axis=2) + nsoft.astype(np.float64)[None, :] ** 2
        fac = ns[None, :, 3].astype(np.float64) * q2 ** -1.5
        out += (dn * fac[:, :, None]).sum(axis=1).T
    return out, dict(n_near=int(near.sum()), D2_scaled=D2, scale=sc, M0=M0)


def reference(s32, soft, t32, chunk=64):
    out = np.zeros((3, t32.shape[0]))
    S = s32.astype(np.float64)
    e2 = soft.astype(np.float64) ** 2
    for t0 in range(0, t32.shape[0], chunk):
        T = t32[t0:t0 + chunk].astype(np.float64)
        d = S[None, :, :3] - T[:, None, :3]
        q2 = (d * d).sum(axis=2) + e2[None, :]
        fac = S[None, :, 3] * q2 ** -1.5
        out[:, t0:t0 + chunk] = (d * fac[:, :, None]).sum(axis=1).T
    return out


def metrics(a, ref):
    sa, sr = a - a[:, -1:], ref - ref[:, -1:]
    return dict(raw_gate=rel_err(a, ref), raw_strict=rel_err(a, ref, 1e-3),
                resid_gate=rel_err(sa[:, :-1], sr[:, :-1]), resid_strict=rel_err(sa[:, :-1], sr[:, :-1], 1e-3))


def main():
    n_src = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    g = make_targets(64)
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(g.evolved_grid) - 1, n_t, replace=False))
    tgt = np.concatenate([g.evolved_grid[pick], g.evolved_grid[-1:]])
    pos, mass, eps = make_sources(n_src, seed=1776)
    s32 = np.concatenate([pos - CENTER, mass[:, None]], axis=1).astype(F)
    t32 = np.concatenate([tgt - CENTER, np.zeros((tgt.shape[0], 1))], axis=1).astype(F)
    soft = eps.astype(F)
    ref = reference(s32, soft, t32)
    res = {"n_src": n_src, "n_tgt": int(t32.shape[0]), "cases": []}
    cases = [
        dict(fold=512), dict(fold=128), dict(fold=64), dict(fold=32), dict(fold=16),
        dict(exact_acc=True), dict(exact_acc=True, rsq_ulp=0.0),
        dict(fold=512, mf=False), dict(exact_acc=True, mf=False),
        dict(exact_acc=True, near_cap=16 * max(4096, n_src // 512)),
    ]
    for kw in cases:
        a, info = run(s32, soft, t32, **kw)
        row = dict(kw)
        row.update(info)
        row.update(metrics(a, ref))
        res["cases"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_error_budget_sim.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

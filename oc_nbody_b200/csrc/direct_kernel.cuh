// The streaming direct-sum kernel (K1 fast set, K4), templated over its tuning axes.
// Included by direct_sum.cu only.
#pragma once
#include "ocg_internal.cuh"
#include "streamk.cuh"

typedef unsigned long long u64;

// ---------------------------------------------------------------- packed-fp32 + PTX helpers ----
__device__ __forceinline__ u64 f2_pack(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) {
  u64 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------- work items ----
template <int TPT, int NTHR = OCG_CONSUMER_THREADS>
__device__ __forceinline__ void decode_item(const DirectParams& p, int item, long long& tgt_begin, int& tgt_count,
                                            long long& tile_begin, int& tile_count, long long& slot) {
  if (p.items) {
    OcgWorkItem w = p.items[item];
    tgt_begin = w.tgt_begin;
    tgt_count = w.tgt_count;
    tile_begin = w.tile_begin;
    tile_count = w.tile_count;
    slot = w.out_slot;
  } else {
    const int CT = NTHR * TPT;
    int chunk = item / p.n_ttiles;
    int tt = item - chunk * p.n_ttiles;
    tgt_begin = (long long)tt * CT;
    long long rem = p.n_tgt - tgt_begin;
    tgt_count = rem < CT ? (int)rem : CT;
    tile_begin = (long long)chunk * p.tiles_per_chunk;
    long long avail = (long long)(*p.n_fast_tiles) - tile_begin;
    tile_count = avail <= 0 ? 0 : (avail < p.tiles_per_chunk ? (int)avail : p.tiles_per_chunk);
    slot = chunk;
  }
}

// ------------------------------------------------------------------ one tile, packed FP32 ----
// Per pair of sources and target: 12 FMA-pipe + 2 MUFU instructions
//   d   = xs + (-xt)                     3 FADD2
//   r2  = dx*dx + dy*dy + dz*dz + e2     3 FFMA2
//   r6  = (r2*r2)*r2                     2 FMUL2
//   y3  = rsqrt(r6) = r^-3               2 MUFU.RSQ   (ONE approximate op per r^-3: ~3x less error
//   sc  = m * y3                         1 FMUL2       than cubing an approximate r^-1)
//   a  += d * sc                         3 FFMA2
//   phi += sc * r2  (= m/r)              1 FFMA2      (POT only)
// Coordinates are pre-scaled by a power of two so that r6 stays inside the FP32 range; GUARD keeps
// the classic rsqrt(r2)^3 form (no range assumption when eps2 == 0) and skips r2 + e2 == 0 pairs.
// PIPE: the five LDS.128 of source group j+1 are issued before the arithmetic of group j (register double
// buffering across loop iterations); without it every iteration starts by waiting for its own loads.
template <int TPT, bool POT, bool GUARD, int UNR, bool PIPE>
__device__ __forceinline__ void tile_packed(const float* __restrict__ stage, const float (&tx)[TPT],
                                            const float (&ty)[TPT], const float (&tz)[TPT],
                                            double (&dacc)[TPT][POT ? 4 : 3]) {
  const float4* sx = reinterpret_cast<const float4*>(stage);
  const float4* sy = sx + OCG_TS / 4;
  const float4* sz = sy + OCG_TS / 4;
  const float4* sm = sz + OCG_TS / 4;
  const float4* se = sm + OCG_TS / 4;
  u64 ntx[TPT], nty[TPT], ntz[TPT], ax[TPT], ay[TPT], az[TPT], ap[TPT];
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    // negated + duplicated target coordinate; ptxas folds it into a broadcast operand (R.F32)
    ntx[t] = f2_pack(-tx[t], -tx[t]), nty[t] = f2_pack(-ty[t], -ty[t]), ntz[t] = f2_pack(-tz[t], -tz[t]);
    ax[t] = ay[t] = az[t] = ap[t] = 0ull;
  }
  float4 Xn, Yn, Zn, Mn, En;
  if (PIPE) Xn = sx[0], Yn = sy[0], Zn = sz[0], Mn = sm[0], En = se[0];
#pragma unroll UNR
  for (int j = 0; j < OCG_TS / 4; ++j) {
    float4 X, Y, Z, M, E;
    if (PIPE) {
      X = Xn, Y = Yn, Z = Zn, M = Mn, E = En;
      const int jn = j + 1 < OCG_TS / 4 ? j + 1 : j;  // last iteration re-reads its own group (harmless)
      Xn = sx[jn], Yn = sy[jn], Zn = sz[jn], Mn = sm[jn], En = se[jn];
    } else {
      X = sx[j], Y = sy[j], Z = sz[j], M = sm[j], E = se[j];
    }
    const u64 xs[2] = {f2_pack(X.x, X.y), f2_pack(X.z, X.w)};
    const u64 ys[2] = {f2_pack(Y.x, Y.y), f2_pack(Y.z, Y.w)};
    const u64 zs[2] = {f2_pack(Z.x, Z.y), f2_pack(Z.z, Z.w)};
    const u64 ms[2] = {f2_pack(M.x, M.y), f2_pack(M.z, M.w)};
    const u64 es[2] = {f2_pack(E.x, E.y), f2_pack(E.z, E.w)};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
      for (int t = 0; t < TPT; ++t) {
        const u64 dx = f2_add(xs[q], ntx[t]);
        const u64 dy = f2_add(ys[q], nty[t]);
        const u64 dz = f2_add(zs[q], ntz[t]);
        u64 r2 = f2_fma(dx, dx, es[q]);
        r2 = f2_fma(dy, dy, r2);
        r2 = f2_fma(dz, dz, r2);
        u64 sc;
        if (GUARD) {
          float r2a, r2b;
          f2_unpack(r2, r2a, r2b);
          const float ria = r2a > 0.f ? rsqrt_approx(r2a) : 0.f;
          const float rib = r2b > 0.f ? rsqrt_approx(r2b) : 0.f;
          const u64 ri = f2_pack(ria, rib);
          const u64 mri = f2_mul(ms[q], ri);
          sc = f2_mul(mri, f2_mul(ri, ri));
          if (POT) ap[t] = f2_add(ap[t], mri);
        } else {
          const u64 r6 = f2_mul(f2_mul(r2, r2), r2);
          float r6a, r6b;
          f2_unpack(r6, r6a, r6b);
          const u64 y3 = f2_pack(rsqrt_approx(r6a), rsqrt_approx(r6b));
          sc = f2_mul(ms[q], y3);
          if (POT) ap[t] = f2_fma(sc, r2, ap[t]);
        }
        ax[t] = f2_fma(dx, sc, ax[t]);
        ay[t] = f2_fma(dy, sc, ay[t]);
        az[t] = f2_fma(dz, sc, az[t]);
      }
    }
  }
  // fold the tile's FP32 partial sums (even / odd sources) into the FP64 accumulators
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    float lo, hi;
    f2_unpack(ax[t], lo, hi);
    dacc[t][0] += (double)lo + (double)hi;
    f2_unpack(ay[t], lo, hi);
    dacc[t][1] += (double)lo + (double)hi;
    f2_unpack(az[t], lo, hi);
    dacc[t][2] += (double)lo + (double)hi;
    if (POT) {
      f2_unpack(ap[t], lo, hi);
      dacc[t][POT ? 3 : 0] -= (double)lo + (double)hi;
    }
  }
}

// ------------------------------------------------------------------ one tile, scalar FP32 ----
// Same maths with plain FADD/FFMA/FMUL (13 issue slots per interaction): the measured baseline.
template <int TPT, bool POT, bool GUARD, int UNR>
__device__ __forceinline__ void tile_scalar(const float* __restrict__ stage, const float (&tx)[TPT],
                                            const float (&ty)[TPT], const float (&tz)[TPT],
                                            double (&dacc)[TPT][POT ? 4 : 3]) {
  const float4* sx = reinterpret_cast<const float4*>(stage);
  const float4* sy = sx + OCG_TS / 4;
  const float4* sz = sy + OCG_TS / 4;
  const float4* sm = sz + OCG_TS / 4;
  const float4* se = sm + OCG_TS / 4;
  float ax[TPT], ay[TPT], az[TPT], ap[TPT];
#pragma unroll
  for (int t = 0; t < TPT; ++t) ax[t] = ay[t] = az[t] = ap[t] = 0.f;
#pragma unroll UNR
  for (int j = 0; j < OCG_TS / 4; ++j) {
    const float4 X = sx[j], Y = sy[j], Z = sz[j], M = sm[j], E = se[j];
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w};
    const float zs[4] = {Z.x, Z.y, Z.z, Z.w}, ms[4] = {M.x, M.y, M.z, M.w};
    const float es[4] = {E.x, E.y, E.z, E.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int t = 0; t < TPT; ++t) {
        const float dx = xs[q] - tx[t], dy = ys[q] - ty[t], dz = zs[q] - tz[t];
        float r2 = fmaf(dx, dx, es[q]);
        r2 = fmaf(dy, dy, r2);
        r2 = fmaf(dz, dz, r2);
        float sc;
        if (GUARD) {
          const float ri = r2 > 0.f ? rsqrt_approx(r2) : 0.f;
          const float mri = ms[q] * ri;
          sc = mri * (ri * ri);
          if (POT) ap[t] += mri;
        } else {
          const float y3 = rsqrt_approx((r2 * r2) * r2);
          sc = ms[q] * y3;
          if (POT) ap[t] = fmaf(sc, r2, ap[t]);
        }
        ax[t] = fmaf(dx, sc, ax[t]);
        ay[t] = fmaf(dy, sc, ay[t]);
        az[t] = fmaf(dz, sc, az[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    dacc[t][0] += (double)ax[t];
    dacc[t][1] += (double)ay[t];
    dacc[t][2] += (double)az[t];
    if (POT) dacc[t][POT ? 3 : 0] -= (double)ap[t];
  }
}

// ------------------------------------------------------------------------------ the kernel ----
// TPT    : targets per consumer thread (1 or 2)
// POT    : also accumulate the potential (4th component)
// GUARD  : see above
// PACKED : FADD2/FFMA2/FMUL2 (true) or scalar FP32 (false)
// DED    : a dedicated 9th warp issues the TMA copies (false: lane 0 of warp 0 does; 256-thread CTA)
// MINB   : CTAs per SM the register allocation is bounded for
// UNR    : unroll of the 4-source group loop
template <int TPT, bool POT, bool GUARD, bool PACKED, bool DED, int MINB, int UNR, bool PIPE = false>
__global__ void __launch_bounds__(OCG_CONSUMER_THREADS + (DED ? 32 : 0), MINB) direct_sum_kernel(const DirectParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + OCG_NSTAGE * OCG_TILE_BYTES);
  uint64_t* empty_bar = full_bar + OCG_NSTAGE;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  constexpr int NC = POT ? 4 : 3;

  if (tid == 0) {
    for (int s = 0; s < OCG_NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], OCG_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t it = 0;  // running tile counter: stage = it % NSTAGE, phase = (it / NSTAGE) & 1

  auto issue_tile = [&](uint32_t n, const float* src) {
    const uint32_t s = n % OCG_NSTAGE, ph = (n / OCG_NSTAGE) & 1u;
    mbar_wait(&empty_bar[s], ph ^ 1u);
    mbar_expect_tx(&full_bar[s], OCG_TILE_BYTES);
    tma_bulk_g2s(stage_base + s * OCG_TILE_FLOATS, src, OCG_TILE_BYTES, &full_bar[s]);
  };

  if (DED && warp == OCG_CONSUMER_WARPS) {
    // ===== dedicated TMA producer warp: one elected lane streams source tiles into the ring =====
    if (lane == 0) {
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        long long tgt_begin, tile_begin, slot;
        int tgt_count, tile_count;
        decode_item<TPT>(p, item, tgt_begin, tgt_count, tile_begin, tile_count, slot);
        const float* src = p.tiles + tile_begin * (long long)OCG_TILE_FLOATS;
        for (int k = 0; k < tile_count; ++k, ++it) issue_tile(it, src + (long long)k * OCG_TILE_FLOATS);
      }
    }
    return;
  }

  // ===== consumer warps =====
  const float scale = p.scale_ptr ? *p.scale_ptr : p.scale_val;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    long long tgt_begin, tile_begin, slot;
    int tgt_count, tile_count;
    decode_item<TPT>(p, item, tgt_begin, tgt_count, tile_begin, tile_count, slot);
    const float* src = p.tiles + tile_begin * (long long)OCG_TILE_FLOATS;

    if (!DED && tid == 0) {  // prologue of the ring for this item
      const int pre = tile_count < OCG_NSTAGE - 1 ? tile_count : OCG_NSTAGE - 1;
      for (int k = 0; k < pre; ++k) issue_tile(it + k, src + (long long)k * OCG_TILE_FLOATS);
    }

    float tx[TPT], ty[TPT], tz[TPT];
    double dacc[TPT][NC];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int local = t * OCG_CONSUMER_THREADS + tid;
      const long long gi = tgt_begin + (local < tgt_count ? local : tgt_count - 1);
      const float4 T = __ldg(&p.tgt[gi]);
      tx[t] = T.x * scale, ty[t] = T.y * scale, tz[t] = T.z * scale;  // power of two: exact
#pragma unroll
      for (int c = 0; c < NC; ++c) dacc[t][c] = 0.0;
    }

    for (int k = 0; k < tile_count; ++k, ++it) {
      if (!DED && tid == 0 && k + OCG_NSTAGE - 1 < tile_count)
        issue_tile(it + OCG_NSTAGE - 1, src + (long long)(k + OCG_NSTAGE - 1) * OCG_TILE_FLOATS);
      const uint32_t s = it % OCG_NSTAGE, ph = (it / OCG_NSTAGE) & 1u;
      mbar_wait(&full_bar[s], ph);
      const float* stage = stage_base + s * OCG_TILE_FLOATS;
      if (PACKED) tile_packed<TPT, POT, GUARD, UNR, PIPE>(stage, tx, ty, tz, dacc);
      else tile_scalar<TPT, POT, GUARD, UNR>(stage, tx, ty, tz, dacc);
      // this warp is done reading stage s: hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int local = t * OCG_CONSUMER_THREADS + tid;
      if (local < tgt_count) {
        const long long gi = tgt_begin + local;
#pragma unroll
        for (int c = 0; c < NC; ++c) p.partial[(slot * NC + c) * p.out_stride + gi] = dacc[t][c];
      }
    }
  }
}

// =====================================================================================================
// Target-paired form.  Measured on B200 (tools/probe_ops.py): every register the LDS.128 source loads write
// back costs the FMA pipe ~0.75 cycle (48 FFMA2 + 10 LDS.128 run 24% slower than 48 FFMA2 alone), i.e. operand
// feeding, not arithmetic, is what keeps the source-paired loop at ~80% FMA-pipe activity.  Here the two lanes
// of a packed instruction are two TARGETS and the source is the broadcast operand (FADD2/FFMA2/FMUL2 accept
// a scalar .F32 operand), so one 4-source group (5 LDS.128 = 20 registers) feeds 4*NP pair-units instead of
// 2*TPT, and a target needs 3 accumulator registers instead of 6.
//   NP      : target pairs per thread (targets per thread = 2*NP; CTA tile = 512*NP targets)
//   SMEMACC : FP64 accumulators live in shared memory (touched once per tile) instead of registers
// DBG (timing experiments only, results are wrong): 1 = MUFU.RSQ replaced by an ALU-pipe bit trick,
// 2 = source registers loaded once per tile instead of per group, 4 = no MUFU at all, 8 = accumulate FFMA2 with
// two distinct register pairs instead of three, 16 = mass-folded loop with FADD2 differences (bits combine).
// MF ("mass-folded", K1 only, !POT): the tile holds w*x | w*y | w*z | w | w^2*e2 with w = (m/M0)^-1/2, so that
//   d'  = w*xs - w*xt = fma(-xt, w, xs')          3 FFMA2   (instead of 3 FADD2)
//   r2' = w^2 (r^2 + e2),  r6' = r2'^3            3 FFMA2 + 2 FMUL2
//   y3  = rsqrt(r6') = w^-3 (r^2+e2)^-3/2         MUFU.RSQ
//   a  += d' * y3 = (m/M0) d (r^2+e2)^-3/2        3 FFMA2   -- the m*y3 multiply is gone:
// 11 FMA-pipe operations per interaction instead of 12 (ceiling 20/22 = 91% of FP32 peak instead of 83%).
// The price is the rounding of xs' = fl(w*xs): relative error 2^-24 |xs|/|d| in d', bounded because every
// source closer to the target box than the precision radius is in the FP64 NEAR set (direct_sum.cu).
// FOLD: sources per call (one FP32 accumulation run); `stage` points at the first of them inside the tile.
// RT: the number of 4-source groups comes at run time (`nj`, staggered folds) instead of FOLD / 4.
template <int NP, bool POT, int UNR, int DBG, int MF = 0, int FOLD = OCG_TS, bool RT = false>
__device__ __forceinline__ void tile_tpair(const float* __restrict__ stage, const u64 (&ntx)[NP], const u64 (&nty)[NP],
                                           const u64 (&ntz)[NP], u64 (&ax)[NP], u64 (&ay)[NP], u64 (&az)[NP],
                                           u64 (&ap)[NP], int nj = FOLD / 4) {
  const float4* sx = reinterpret_cast<const float4*>(stage);
  const float4* sy = sx + OCG_TS / 4;
  const float4* sz = sy + OCG_TS / 4;
  const float4* sm = sz + OCG_TS / 4;
  const float4* se = sm + OCG_TS / 4;
  const float4* sv = se + OCG_TS / 4;  // mass-folded tiles with potential: 6th array, 1/w = (m/M0)^1/2
  float4 X0 = sx[0], Y0 = sy[0], Z0 = sz[0], M0 = sm[0], E0 = se[0];
#pragma unroll UNR
  for (int j = 0; j < (RT ? nj : FOLD / 4); ++j) {
    float4 X, Y, Z, M, E, V = make_float4(0.f, 0.f, 0.f, 0.f);
    if (DBG & 2) {
      X = X0, Y = Y0, Z = Z0, M = M0, E = E0;
      X0.x += 1e-7f;  // keep the loop body from being hoisted
    } else {
      X = sx[j], Y = sy[j], Z = sz[j], M = sm[j], E = se[j];
      if (MF && POT) V = sv[j];
    }
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
    const float ms[4] = {M.x, M.y, M.z, M.w}, es[4] = {E.x, E.y, E.z, E.w}, vs[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // duplicated source scalars: ptxas folds them into broadcast (.F32) operands
      const u64 xb = f2_pack(xs[q], xs[q]), yb = f2_pack(ys[q], ys[q]), zb = f2_pack(zs[q], zs[q]);
      // MF == 2: the w array of the tile is stored with adjacent sources swapped, so that w comes from a register
      // of the opposite bank parity to w*x, w*y, w*z (LDS.128 puts source q of every component in R(4k+q))
      const float mq = MF == 2 ? ms[q ^ 1] : ms[q];
      const u64 mb = f2_pack(mq, mq), eb = f2_pack(es[q], es[q]);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const bool dfma = MF && !(DBG & 16);
        const u64 dx = dfma ? f2_fma(ntx[p], mb, xb) : f2_add(ntx[p], xb);
        const u64 dy = dfma ? f2_fma(nty[p], mb, yb) : f2_add(nty[p], yb);
        const u64 dz = dfma ? f2_fma(ntz[p], mb, zb) : f2_add(ntz[p], zb);
        u64 r2 = f2_fma(dx, dx, eb);
        r2 = f2_fma(dy, dy, r2);
        r2 = f2_fma(dz, dz, r2);
        const u64 r6 = f2_mul(f2_mul(r2, r2), r2);
        float r6a, r6b;
        f2_unpack(r6, r6a, r6b);
        u64 y3;
        if (DBG & 4) {
          y3 = r6;
        } else if (DBG & 1) {
          y3 = f2_pack(__int_as_float(0x5f3759df - (__float_as_int(r6a) >> 1)),
                       __int_as_float(0x5f3759df - (__float_as_int(r6b) >> 1)));
        } else {
          y3 = f2_pack(rsqrt_approx(r6a), rsqrt_approx(r6b));
        }
        const u64 sc = MF ? y3 : f2_mul(y3, mb);
        if (POT) {
          // plain: m y3 r2 = m/r.  mass-folded: y3' r2' = 1/(w r), times 1/w = (m/M0)/r: one more multiply, the same
          // 13 FMA-pipe operations per interaction as the plain form with potential
          if (MF) ap[p] = f2_fma(f2_mul(y3, r2), f2_pack(vs[q], vs[q]), ap[p]);
          else ap[p] = f2_fma(sc, r2, ap[p]);
        }
        if (DBG & 8) {  // two distinct register pairs per instruction instead of three
          ax[p] = f2_fma(dx, ax[p], ax[p]);
          ay[p] = f2_fma(dy, ay[p], ay[p]);
          az[p] = f2_fma(sc, az[p], az[p]);
        } else {
          ax[p] = f2_fma(dx, sc, ax[p]);
          ay[p] = f2_fma(dy, sc, ay[p]);
          az[p] = f2_fma(dz, sc, az[p]);
        }
      }
    }
  }
}

//   NW      : consumer warps per CTA (CTA = 32*NW threads; CTA tile = 64*NW*NP targets)
//   FOLD    : sources per FP32 accumulation run.  The FP32 partial sums of a run are folded into the FP64 per-target
//             accumulators after FOLD sources; the rounding error of the FP32 running sums grows linearly with FOLD and
//             is THE error term of the kernel (tools/sim_fp32_error.py: with exact accumulation the tidal residual is
//             good to 7e-7 strict, with FOLD = 512 to 7e-5), so FOLD is the accuracy/throughput knob.
//   STAG    : 0 = every warp folds at the same source indices.  1 / 2 = half of the warps (1: warps >= NW/2, 2: odd warps)
//             fold half a run later (their first and last run of a tile are FOLD/2 long), so that the two warps of a
//             scheduler are not both in the latency-bound fold at the same time.
template <int NP, bool POT, bool SMEMACC, int MINB, int UNR, int NW = OCG_CONSUMER_WARPS, int DBG = 0, int MF = 0,
          int FOLD = OCG_TS, int STAG = 0>
__global__ void __launch_bounds__(32 * NW, MINB) direct_sum_tp_kernel(const DirectParams p) {
  static_assert(OCG_TS % FOLD == 0 && FOLD % 4 == 0, "FOLD must divide the tile");
  constexpr int NTHR = 32 * NW;
  constexpr int TILE_FLOATS = OCG_TILE_ARRAYS(MF, POT) * OCG_TS, TILE_BYTES = TILE_FLOATS * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + OCG_NSTAGE * TILE_BYTES);
  uint64_t* empty_bar = full_bar + OCG_NSTAGE;
  // FP64 accumulators: the two targets of a pair sit side by side, sacc2[(c*NP + pp)*NTHR + tid] = {2pp, 2pp+1}
  double2* sacc2 = reinterpret_cast<double2*>(smem_raw + OCG_NSTAGE * TILE_BYTES + 128);
  constexpr int NC = POT ? 4 : 3;
  constexpr int T = 2 * NP;
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < OCG_NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t it = 0;
  auto issue_tile = [&](uint32_t n, const float* src) {
    const uint32_t s = n % OCG_NSTAGE, ph = (n / OCG_NSTAGE) & 1u;
    mbar_wait(&empty_bar[s], ph ^ 1u);
    mbar_expect_tx(&full_bar[s], TILE_BYTES);
    tma_bulk_g2s(stage_base + s * TILE_FLOATS, src, TILE_BYTES, &full_bar[s]);
  };

  const float scale = p.scale_ptr ? *p.scale_ptr : p.scale_val;
  __shared__ int s_last;
  __shared__ long long s_sk[5];  // units of the pass, participating CTAs, end of this CTA's unit range, tiles per row in
                                 // this pass, first tile of the pass: read back in the segment epilogue instead of
                                 // being held in registers across the tile loop
  // ---- stream-K: this CTA's share of the (target tile x source tile) units, pass by pass, see streamk.cuh ----
  if (sk_units(p.sk) == 0) {  // nothing to stream (K1 with an empty fast set): the field of no sources is zero
    if (!p.accumulate)
      for (long long i = blockIdx.x * (long long)NTHR + tid; i < p.out_n; i += (long long)gridDim.x * NTHR) {
        p.out_acc[i] = 0.0, p.out_acc[p.out_n + i] = 0.0, p.out_acc[2 * p.out_n + i] = 0.0;
        if (POT) p.out_pot[i] = 0.0;
      }
    return;
  }
  int n_pass = 1;
  if (!p.sk.rows) {
    const long long nst_total = sk_nst(p.sk), cap = sk_pass_cap(p.sk, nst_total);
    n_pass = (int)((nst_total + cap - 1) / cap);
  }
  for (int pass = 0; pass < n_pass; ++pass) {
  long long u;
  {
    long long nst_pass = 0, tile0 = 0;
    if (!p.sk.rows) {
      const long long nst_total = sk_nst(p.sk), cap = sk_pass_cap(p.sk, nst_total);
      tile0 = pass * cap;
      nst_pass = nst_total - tile0 < cap ? nst_total - tile0 : cap;
    }
    const long long U = sk_units(p.sk, nst_pass);
    const long long G_ = sk_ctas(U, gridDim.x);  // CTAs that take part: every one of them gets at least one unit
    u = blockIdx.x < G_ ? sk_first_unit(blockIdx.x, U, G_) : 0;
    __syncthreads();  // the previous pass's epilogue has read s_sk
    if (tid == 0)
      s_sk[0] = U, s_sk[1] = G_, s_sk[2] = blockIdx.x < G_ ? sk_first_unit(blockIdx.x + 1, U, G_) : 0, s_sk[3] = nst_pass, s_sk[4] = tile0;
    __syncthreads();
  }
  for (int row = u < s_sk[2] ? sk_find_row(p.sk, u, s_sk[3]) : 0; u < s_sk[2]; ++row) {
    int tgt_count, tile_count;
    long long tgt_begin;
    const float* src;
    {
      const long long rs = sk_row_start(p.sk, row, s_sk[3]), re = sk_row_start(p.sk, row + 1, s_sk[3]), u_end = s_sk[2];
      long long tile_begin;
      tile_count = (int)((re < u_end ? re : u_end) - u);
      sk_row(p.sk, row, tgt_begin, tgt_count, tile_begin);
      src = p.tiles + (tile_begin + s_sk[4] + (u - rs)) * (long long)TILE_FLOATS;
      u += tile_count;
    }
    if (tid == 0) {
      const int pre = tile_count < OCG_NSTAGE - 1 ? tile_count : OCG_NSTAGE - 1;
      for (int k = 0; k < pre; ++k) issue_tile(it + k, src + (long long)k * TILE_FLOATS);
    }

    u64 ntx[NP], nty[NP], ntz[NP];
    double dacc[SMEMACC ? 1 : T][NC];
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) {
      float c[2][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int local = (2 * pp + h) * NTHR + tid;
        const long long gi = tgt_begin + (local < tgt_count ? local : tgt_count - 1);
        const float4 Tg = __ldg(&p.tgt[gi]);
        c[h][0] = -Tg.x * scale, c[h][1] = -Tg.y * scale, c[h][2] = -Tg.z * scale;
      }
      ntx[pp] = f2_pack(c[0][0], c[1][0]), nty[pp] = f2_pack(c[0][1], c[1][1]), ntz[pp] = f2_pack(c[0][2], c[1][2]);
    }
    if (SMEMACC) {
#pragma unroll
      for (int i = 0; i < NC * NP; ++i) sacc2[i * NTHR + tid] = make_double2(0.0, 0.0);
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int c = 0; c < NC; ++c) dacc[t][c] = 0.0;
    }

    for (int k = 0; k < tile_count; ++k, ++it) {
      if (tid == 0 && k + OCG_NSTAGE - 1 < tile_count)
        issue_tile(it + OCG_NSTAGE - 1, src + (long long)(k + OCG_NSTAGE - 1) * TILE_FLOATS);
      const uint32_t s = it % OCG_NSTAGE, ph = (it / OCG_NSTAGE) & 1u;
      mbar_wait(&full_bar[s], ph);
      // fold one run's FP32 sums into the FP64 accumulators (lane lo -> target 2p, hi -> target 2p+1)
      auto fold_run = [&](const u64 (&ax)[NP], const u64 (&ay)[NP], const u64 (&az)[NP], const u64 (&ap)[NP]) {
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
          float v[4][2];
          f2_unpack(ax[pp], v[0][0], v[0][1]);
          f2_unpack(ay[pp], v[1][0], v[1][1]);
          f2_unpack(az[pp], v[2][0], v[2][1]);
          f2_unpack(ap[pp], v[3][0], v[3][1]);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const double a0 = c < 3 ? (double)v[c][0] : -(double)v[3][0];
            const double a1 = c < 3 ? (double)v[c][1] : -(double)v[3][1];
            if (SMEMACC) {
              double2 q = sacc2[(c * NP + pp) * NTHR + tid];
              q.x += a0, q.y += a1;
              sacc2[(c * NP + pp) * NTHR + tid] = q;
            } else {
              dacc[SMEMACC ? 0 : 2 * pp][c] += a0;
              dacc[SMEMACC ? 0 : 2 * pp + 1][c] += a1;
            }
          }
        }
      };
      if (STAG == 0) {
#pragma unroll 1
        for (int b = 0; b < OCG_TS / FOLD; ++b) {
          u64 ax[NP], ay[NP], az[NP], ap[NP];
#pragma unroll
          for (int pp = 0; pp < NP; ++pp) ax[pp] = ay[pp] = az[pp] = ap[pp] = 0ull;
          tile_tpair<NP, POT, UNR, DBG, MF, FOLD>(stage_base + s * TILE_FLOATS + b * FOLD, ntx, nty, ntz, ax, ay, az, ap);
          if (b == OCG_TS / FOLD - 1) {  // the stage has been read: hand it back before folding
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
          }
          fold_run(ax, ay, az, ap);
        }
      } else {
        const int warp = tid >> 5;
        const bool late = STAG == 1 ? warp >= NW / 2 : (warp & 1);
        int off = 0, len = late ? FOLD / 2 : FOLD;
#pragma unroll 1
        while (off < OCG_TS) {
          u64 ax[NP], ay[NP], az[NP], ap[NP];
#pragma unroll
          for (int pp = 0; pp < NP; ++pp) ax[pp] = ay[pp] = az[pp] = ap[pp] = 0ull;
          tile_tpair<NP, POT, UNR, DBG, MF, FOLD, true>(stage_base + s * TILE_FLOATS + off, ntx, nty, ntz, ax, ay, az, ap, len / 4);
          off += len;
          len = OCG_TS - off < FOLD ? OCG_TS - off : FOLD;
          if (off >= OCG_TS) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
          }
          fold_run(ax, ay, az, ap);
        }
      }
    }

    // ---- segment epilogue: who shares this row, and the final scaling (recomputed here, not carried over the loop) ----
    const long long U = s_sk[0], G_ = s_sk[1];
    const long long rs = sk_row_start(p.sk, row, s_sk[3]), re = sk_row_start(p.sk, row + 1, s_sk[3]);
    const int first_cta = sk_cta_of(rs, U, G_);
    int n_sharers = sk_cta_of(re - 1, U, G_) - first_cta + 1;
    long long slot = (long long)blockIdx.x - first_cta;
    if (n_pass > 1) {
      // the row has sharers in every pass: slots pass-major, one ticket counter.  All passes but the last have `cap` tiles.
      const long long nst_total = sk_nst(p.sk), cap = sk_pass_cap(p.sk, nst_total);
      const long long Uf = p.sk.n_rows * cap, Gf = sk_ctas(Uf, gridDim.x);
      const int sf = sk_sharers(row * cap, (row + 1) * cap, Uf, Gf);
      const long long cl = nst_total - (n_pass - 1) * cap, Ul = p.sk.n_rows * cl, Gl = sk_ctas(Ul, gridDim.x);
      const int sl = sk_sharers(row * cl, (row + 1) * cl, Ul, Gl);
      slot += (long long)pass * sf;
      n_sharers = (n_pass - 1) * sf + sl;
    }
    {
      long long unused_tile;
      sk_row(p.sk, row, tgt_begin, tgt_count, unused_tile);  // re-read: not live across the tile loop
    }
    // final scaling of the sums: G s^2 (potential: G s), times M0 when the tiles are mass-folded
    const double gm = p.G * (p.m0_ptr ? (double)*p.m0_ptr : 1.0), fa = gm * (double)scale * (double)scale, fp = gm * (double)scale;
    auto write_out = [&](long long gi, int c, double sum) {
      if (c < 3) {
        double* dst = p.out_acc + (long long)c * p.out_n + gi;
        const double v = sum * fa;
        *dst = p.accumulate ? *dst + v : v;
      } else {
        if (p.self_e2s > 0.f)  // K4: targets are sources; take the self term -m/eps out with the kernel's own FP32 expression
          sum += (double)((__ldg(&p.tgt[gi]).w * rsqrt_approx((p.self_e2s * p.self_e2s) * p.self_e2s)) * p.self_e2s);
        const double v = sum * fp;
        p.out_pot[gi] = p.accumulate ? p.out_pot[gi] + v : v;
      }
    };
    auto own = [&](int t, int c) -> double {
      if (SMEMACC) {
        const double2 q = sacc2[(c * NP + (t >> 1)) * NTHR + tid];
        return (t & 1) ? q.y : q.x;
      }
      return dacc[SMEMACC ? 0 : t][c];
    };
    if (n_sharers == 1) {
      // the whole row was streamed here: scale and write the field
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int local = t * NTHR + tid;
        if (local < tgt_count)
#pragma unroll
          for (int c = 0; c < NC; ++c) write_out(tgt_begin + local, c, own(t, c));
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int local = t * NTHR + tid;
        if (local < tgt_count)
#pragma unroll
          for (int c = 0; c < NC; ++c) p.partial[(slot * NC + c) * p.out_stride + tgt_begin + local] = own(t, c);
      }
      if (sk_last_of_row(p.sk, row, n_sharers, &s_last)) {
        // last of the row's CTAs: add the slots in slot order (deterministic whatever the arrival order); two targets x NC
        // components per step so that a slot costs one round trip to L2, not 2 NC dependent ones
#pragma unroll
        for (int t = 0; t < T; t += 2) {
          const int l0 = t * NTHR + tid, l1 = (t + 1) * NTHR + tid;
          const long long g0 = tgt_begin + (l0 < tgt_count ? l0 : 0), g1 = tgt_begin + (l1 < tgt_count ? l1 : 0);
          double sum[2][NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) sum[0][c] = sum[1][c] = 0.0;
#pragma unroll 2
          for (int k = 0; k < n_sharers; ++k) {
            const double* base = p.partial + (long long)k * NC * p.out_stride;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              sum[0][c] += __ldcg(base + c * p.out_stride + g0);
              sum[1][c] += __ldcg(base + c * p.out_stride + g1);
            }
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (l0 < tgt_count) write_out(g0, c, sum[0][c]);
            if (l1 < tgt_count) write_out(g1, c, sum[1][c]);
          }
        }
      }
    }
  }
  }  // pass
}

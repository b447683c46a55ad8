// K6: the Hermite force loop — acceleration AND jerk of the cluster's self-gravity — plus the predictor and
// corrector of the 4th-order Hermite scheme (Makino & Aarseth 1992) with the Aarseth time-step criterion.
//
// This is the arithmetic of the AMUSE ph4 worker the reference instantiates at oc_code.py:218-229 (4th-order
// Hermite, Plummer softening epsilon_squared, oc_code.py:225) — SURVEY §8f rank 5.  ph4's block (individual)
// time steps are not reproduced: every star takes the same step (the BRIDGE step divided into `substeps`), the
// Aarseth criterion is evaluated on the device and its minimum handed back so the host can choose `substeps`.
//
//   a_i = G sum_j m_j d / r^3                       d = x_j - x_i,  r^2 = |d|^2 + eps^2
//   j_i = G sum_j m_j [ w / r^3 - 3 (d.w) d / r^5 ] w = v_j - v_i
//
// Same structure as K4 (direct_kernel.cuh): TMA-staged source tiles in a shared-memory ring, target-paired
// packed FP32 (the two lanes of an FFMA2 are two targets, the source is the broadcast operand), FP32 partial
// sums over one 512-source tile, FP64 per-target accumulators in shared memory; the work is cut stream-K fashion
// (streamk.cuh) and the last CTA of a target row adds the row's partial slots in slot order and writes the result
// (run-to-run deterministic, no finish kernel).  26 FMA-pipe operations + 1 MUFU per interaction.
#include "direct_kernel.cuh"

#include <math.h>

#define HM_NSTAGE 3

struct HermiteParams {
  const float* tiles;     // [tile][7][HM_TS]: scaled positions, mass, unscaled velocities
  const float4* tgt_pos;  // [n] recentred (x, y, z, m), unscaled
  const float4* tgt_vel;  // [n] recentred (vx, vy, vz, 0)
  double* partial;        // [slot][NC][out_stride]
  long long out_stride;
  StreamKParams sk;       // rows = target tiles of the shard x the source tiles of their segment (streamk.cuh)
  double* out_acc;        // final outputs [3][out_stride], [3][out_stride], [out_stride], written by a row's last CTA
  double* out_jerk;
  double* out_pot;
  double G, vel_to_len;
  float e2s;    // eps^2 in scaled units
  float scale;  // power-of-two length scale
};

// FP64 particles -> source tiles and float4 targets; each segment recentred on its first particle (position AND
// velocity) in FP64 before rounding to FP32.
__global__ void pack_hermite_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                    const double* __restrict__ mass, long long n,
                                    const long long* __restrict__ seg_off, const long long* __restrict__ seg_tile,
                                    int n_seg, float scale, float* __restrict__ tiles, float4* __restrict__ tgt_pos,
                                    float4* __restrict__ tgt_vel) {
  const long long total_tiles = seg_tile[n_seg];
  const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (slot >= total_tiles * HM_TS) return;
  const long long tile = slot / HM_TS;
  const int j = (int)(slot - tile * HM_TS);
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_tile[mid] <= tile) lo = mid;
    else hi = mid;
  }
  const long long first = seg_off[lo];
  const long long idx = first + (tile - seg_tile[lo]) * HM_TS + j;
  float* T = tiles + tile * (long long)HM_TILE_FLOATS;
  float v[HM_NARR] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (idx < seg_off[lo + 1]) {
    const float x = (float)(pos[idx] - pos[first]), y = (float)(pos[n + idx] - pos[n + first]),
                z = (float)(pos[2 * n + idx] - pos[2 * n + first]);
    const float vx = (float)(vel[idx] - vel[first]), vy = (float)(vel[n + idx] - vel[n + first]),
                vz = (float)(vel[2 * n + idx] - vel[2 * n + first]);
    const float m = (float)mass[idx];
    v[0] = x * scale, v[1] = y * scale, v[2] = z * scale, v[3] = m, v[4] = vx, v[5] = vy, v[6] = vz;
    tgt_pos[idx] = make_float4(x, y, z, m);
    tgt_vel[idx] = make_float4(vx, vy, vz, 0.f);
  }
#pragma unroll
  for (int a = 0; a < HM_NARR; ++a) T[a * HM_TS + j] = v[a];
}

// One source tile against NP target pairs.  Per interaction (one lane of each packed instruction):
//   d  = xs - xt, w = vs - vt                6 FADD
//   r2 = d.d + e2                            3 FFMA
//   rv = d.w                                 1 FMUL + 2 FFMA
//   ri = rsqrt(r2)                           MUFU
//   q  = ri*ri, mri = m*ri, mr3 = mri*q      3 FMUL      (phi += mri: +1 FADD)
//   al = (rv*q)*(-3)                         2 FMUL
//   t  = al*d + w                            3 FFMA
//   j += mr3*t, a += mr3*d                   6 FFMA
// = 26 FMA-pipe operations (41 flop counting an FMA as two).
template <int NP, bool POT, bool GUARD, int UNR>
__device__ __forceinline__ void hermite_tile(const float* __restrict__ stage, const u64 (&ntx)[NP], const u64 (&nty)[NP],
                                             const u64 (&ntz)[NP], const u64 (&ntu)[NP], const u64 (&ntv)[NP],
                                             const u64 (&ntw)[NP], const u64 eb, u64 (&ax)[NP], u64 (&ay)[NP],
                                             u64 (&az)[NP], u64 (&jx)[NP], u64 (&jy)[NP], u64 (&jz)[NP], u64 (&ap)[NP]) {
  const float4* sx = reinterpret_cast<const float4*>(stage);
  const float4* sy = sx + HM_TS / 4;
  const float4* sz = sy + HM_TS / 4;
  const float4* sm = sz + HM_TS / 4;
  const float4* su = sm + HM_TS / 4;
  const float4* sv = su + HM_TS / 4;
  const float4* sw = sv + HM_TS / 4;
  const u64 c3 = f2_pack(-3.0f, -3.0f);
#pragma unroll UNR
  for (int j = 0; j < HM_TS / 4; ++j) {
    const float4 X = sx[j], Y = sy[j], Z = sz[j], M = sm[j], U = su[j], V = sv[j], W = sw[j];
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
    const float ms[4] = {M.x, M.y, M.z, M.w};
    const float us[4] = {U.x, U.y, U.z, U.w}, vs[4] = {V.x, V.y, V.z, V.w}, ws[4] = {W.x, W.y, W.z, W.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // duplicated source scalars: ptxas folds them into broadcast (.F32) operands
      const u64 xb = f2_pack(xs[q], xs[q]), yb = f2_pack(ys[q], ys[q]), zb = f2_pack(zs[q], zs[q]);
      const u64 ub = f2_pack(us[q], us[q]), vb = f2_pack(vs[q], vs[q]), wb = f2_pack(ws[q], ws[q]);
      const u64 mb = f2_pack(ms[q], ms[q]);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const u64 dx = f2_add(ntx[p], xb), dy = f2_add(nty[p], yb), dz = f2_add(ntz[p], zb);
        const u64 du = f2_add(ntu[p], ub), dv = f2_add(ntv[p], vb), dw = f2_add(ntw[p], wb);
        u64 r2 = f2_fma(dx, dx, eb);
        r2 = f2_fma(dy, dy, r2);
        r2 = f2_fma(dz, dz, r2);
        u64 rv = f2_mul(dx, du);
        rv = f2_fma(dy, dv, rv);
        rv = f2_fma(dz, dw, rv);
        float r2a, r2b;
        f2_unpack(r2, r2a, r2b);
        float ria = rsqrt_approx(r2a), rib = rsqrt_approx(r2b);
        if (GUARD) {  // eps2 == 0: coincident pairs (the self pair among them) contribute nothing
          ria = r2a > 0.f ? ria : 0.f;
          rib = r2b > 0.f ? rib : 0.f;
        }
        const u64 ri = f2_pack(ria, rib);
        const u64 qq = f2_mul(ri, ri);
        const u64 mri = f2_mul(mb, ri);
        const u64 mr3 = f2_mul(mri, qq);
        if (POT) ap[p] = f2_add(ap[p], mri);
        const u64 al = f2_mul(f2_mul(rv, qq), c3);
        const u64 tx = f2_fma(al, dx, du), ty = f2_fma(al, dy, dv), tz = f2_fma(al, dz, dw);
        jx[p] = f2_fma(mr3, tx, jx[p]);
        jy[p] = f2_fma(mr3, ty, jy[p]);
        jz[p] = f2_fma(mr3, tz, jz[p]);
        ax[p] = f2_fma(mr3, dx, ax[p]);
        ay[p] = f2_fma(mr3, dy, ay[p]);
        az[p] = f2_fma(mr3, dz, az[p]);
      }
    }
  }
}

//   NP : target pairs per thread (targets per thread = 2*NP); NW : warps per CTA; CTA tile = 64*NW*NP targets
template <int NP, bool POT, bool GUARD, int NW, int UNR, int MINB = 1>
__global__ void __launch_bounds__(32 * NW, MINB) hermite_tp_kernel(const HermiteParams p) {
  constexpr int NTHR = 32 * NW;
  constexpr int NC = POT ? 7 : 6;
  constexpr int T = 2 * NP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + HM_NSTAGE * HM_TILE_BYTES);
  uint64_t* empty_bar = full_bar + HM_NSTAGE;
  double* sacc = reinterpret_cast<double*>(smem_raw + HM_NSTAGE * HM_TILE_BYTES + 128);  // [NC][T][NTHR]
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < HM_NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t it = 0;  // running tile counter: stage = it % NSTAGE, phase = (it / NSTAGE) & 1
  auto issue_tile = [&](uint32_t n, const float* src) {
    const uint32_t s = n % HM_NSTAGE, ph = (n / HM_NSTAGE) & 1u;
    mbar_wait(&empty_bar[s], ph ^ 1u);
    mbar_expect_tx(&full_bar[s], HM_TILE_BYTES);
    tma_bulk_g2s(stage_base + s * HM_TILE_FLOATS, src, HM_TILE_BYTES, &full_bar[s]);
  };

  const float scale = p.scale;
  const u64 eb = f2_pack(p.e2s, p.e2s);
  __shared__ int s_last;
  __shared__ long long s_sk[3];  // units of the launch, participating CTAs, end of this CTA's unit range: read back in
                                 // the segment epilogue instead of being held in registers across the tile loop
  // ---- stream-K: this CTA's share of the (target tile x source tile) units, see streamk.cuh ----
  long long u;
  {
    const long long U = sk_units(p.sk);
    const long long G_ = sk_ctas(U, gridDim.x);  // CTAs that take part: every one of them gets at least one unit
    u = blockIdx.x < G_ ? sk_first_unit(blockIdx.x, U, G_) : 0;
    if (tid == 0) s_sk[0] = U, s_sk[1] = G_, s_sk[2] = blockIdx.x < G_ ? sk_first_unit(blockIdx.x + 1, U, G_) : 0;
    __syncthreads();
  }
  for (int row = u < s_sk[2] ? sk_find_row(p.sk, u) : 0; u < s_sk[2]; ++row) {
    int tgt_count, tile_count;
    long long tgt_begin;
    const float* src;
    {
      const long long rs = sk_row_start(p.sk, row), re = sk_row_start(p.sk, row + 1), u_end = s_sk[2];
      long long tile_begin;
      tile_count = (int)((re < u_end ? re : u_end) - u);
      sk_row(p.sk, row, tgt_begin, tgt_count, tile_begin);
      src = p.tiles + (tile_begin + (u - rs)) * (long long)HM_TILE_FLOATS;
      u += tile_count;
    }
    if (tid == 0) {
      const int pre = tile_count < HM_NSTAGE - 1 ? tile_count : HM_NSTAGE - 1;
      for (int k = 0; k < pre; ++k) issue_tile(it + k, src + (long long)k * HM_TILE_FLOATS);
    }

    u64 ntx[NP], nty[NP], ntz[NP], ntu[NP], ntv[NP], ntw[NP];
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) {
      float c[2][6];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int local = (2 * pp + h) * NTHR + tid;
        const long long gi = tgt_begin + (local < tgt_count ? local : tgt_count - 1);
        const float4 P = __ldg(&p.tgt_pos[gi]);
        const float4 V = __ldg(&p.tgt_vel[gi]);
        c[h][0] = -P.x * scale, c[h][1] = -P.y * scale, c[h][2] = -P.z * scale;  // power of two: exact
        c[h][3] = -V.x, c[h][4] = -V.y, c[h][5] = -V.z;
      }
      ntx[pp] = f2_pack(c[0][0], c[1][0]), nty[pp] = f2_pack(c[0][1], c[1][1]), ntz[pp] = f2_pack(c[0][2], c[1][2]);
      ntu[pp] = f2_pack(c[0][3], c[1][3]), ntv[pp] = f2_pack(c[0][4], c[1][4]), ntw[pp] = f2_pack(c[0][5], c[1][5]);
    }
#pragma unroll
    for (int i = 0; i < NC * T; ++i) sacc[i * NTHR + tid] = 0.0;

    for (int k = 0; k < tile_count; ++k, ++it) {
      if (tid == 0 && k + HM_NSTAGE - 1 < tile_count)
        issue_tile(it + HM_NSTAGE - 1, src + (long long)(k + HM_NSTAGE - 1) * HM_TILE_FLOATS);
      const uint32_t s = it % HM_NSTAGE, ph = (it / HM_NSTAGE) & 1u;
      mbar_wait(&full_bar[s], ph);
      u64 ax[NP], ay[NP], az[NP], jx[NP], jy[NP], jz[NP], ap[NP];
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) ax[pp] = ay[pp] = az[pp] = jx[pp] = jy[pp] = jz[pp] = ap[pp] = 0ull;
      hermite_tile<NP, POT, GUARD, UNR>(stage_base + s * HM_TILE_FLOATS, ntx, nty, ntz, ntu, ntv, ntw, eb, ax, ay, az,
                                        jx, jy, jz, ap);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      // fold this tile's FP32 sums into the FP64 accumulators (lane lo -> target 2p, hi -> target 2p+1)
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) {
        float v[7][2];
        f2_unpack(ax[pp], v[0][0], v[0][1]);
        f2_unpack(ay[pp], v[1][0], v[1][1]);
        f2_unpack(az[pp], v[2][0], v[2][1]);
        f2_unpack(jx[pp], v[3][0], v[3][1]);
        f2_unpack(jy[pp], v[4][0], v[4][1]);
        f2_unpack(jz[pp], v[5][0], v[5][1]);
        f2_unpack(ap[pp], v[6][0], v[6][1]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < NC; ++c) sacc[(c * T + 2 * pp + h) * NTHR + tid] += (double)v[c][h];
      }
    }

    // ---- segment epilogue: who shares this row, and the final scaling (recomputed here, not carried over the loop) ----
    const long long U = s_sk[0], G_ = s_sk[1];
    const long long rs = sk_row_start(p.sk, row), re = sk_row_start(p.sk, row + 1);
    const int first_cta = sk_cta_of(rs, U, G_);
    const int n_sharers = sk_cta_of(re - 1, U, G_) - first_cta + 1;
    const long long slot = (long long)blockIdx.x - first_cta;
    {
      long long unused_tile;
      sk_row(p.sk, row, tgt_begin, tgt_count, unused_tile);  // re-read: not live across the tile loop
    }
    // acc = G s^2 sum, jerk = G s^3 vel_to_len sum (positions scaled by s, velocities not), pot = -G s sum with the self
    // term (m/eps, included because targets == sources) removed with the kernel's own FP32 expression
    const double sc = (double)scale, fa = p.G * sc * sc, fj = fa * sc * p.vel_to_len, fp = p.G * sc;
    auto write_out = [&](long long gi, int c, double sum) {
      if (c < 3) p.out_acc[(long long)c * p.out_stride + gi] = sum * fa;
      else if (c < 6) p.out_jerk[(long long)(c - 3) * p.out_stride + gi] = sum * fj;
      else {
        if (p.e2s > 0.f) sum -= (double)(__ldg(&p.tgt_pos[gi]).w * rsqrt_approx(p.e2s));
        p.out_pot[gi] = -sum * fp;
      }
    };
    if (n_sharers == 1) {
      // the whole row was streamed here: scale and write acceleration, jerk and potential
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int local = t * NTHR + tid;
        if (local < tgt_count)
#pragma unroll
          for (int c = 0; c < NC; ++c) write_out(tgt_begin + local, c, sacc[(c * T + t) * NTHR + tid]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int local = t * NTHR + tid;
        if (local < tgt_count)
#pragma unroll
          for (int c = 0; c < NC; ++c) p.partial[(slot * NC + c) * p.out_stride + tgt_begin + local] = sacc[(c * T + t) * NTHR + tid];
      }
      if (sk_last_of_row(p.sk, row, n_sharers, &s_last)) {
        // last of the row's CTAs: add the slots in slot order (deterministic whatever the arrival order)
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int local = t * NTHR + tid;
          if (local < tgt_count) {
            const long long gi = tgt_begin + local;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              double sum = 0.0;
              for (int k = 0; k < n_sharers; ++k) sum += __ldcg(&p.partial[((long long)k * NC + c) * p.out_stride + gi]);
              write_out(gi, c, sum);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------ small clusters ----
// One launch for one cluster of <= 4096 stars (the reference's own run is N = 1 024, test_options:57): every CTA
// stages all sources in shared memory, a warp owns one target, its 32 lanes split the sources, FP32 pair arithmetic
// with two independent partial sums per lane, FP64 from the warp-shuffle reduction on.
#define HM_SMALL_WARPS 8
#define HM_SMALL_MAX_N 4096
__global__ void __launch_bounds__(32 * HM_SMALL_WARPS) hermite_small_kernel(
    const double* __restrict__ pos, const double* __restrict__ vel, const double* __restrict__ mass, long long n,
    float e2s, float scale, double G, double vel_to_len, long long tgt_begin, long long tgt_end,
    double* __restrict__ acc, double* __restrict__ jerk, double* __restrict__ pot) {
  extern __shared__ float4 s_all[];
  float4* s_pos = s_all;
  float4* s_vel = s_all + n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    s_pos[i] = make_float4((float)(pos[i] - pos[0]) * scale, (float)(pos[n + i] - pos[n]) * scale,
                           (float)(pos[2 * n + i] - pos[2 * n]) * scale, (float)mass[i]);
    s_vel[i] = make_float4((float)(vel[i] - vel[0]), (float)(vel[n + i] - vel[n]), (float)(vel[2 * n + i] - vel[2 * n]), 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t = tgt_begin + (long long)blockIdx.x * HM_SMALL_WARPS + warp;
  if (t >= tgt_end) return;
  const float4 P = s_pos[t], V = s_vel[t];
  float a[2][7];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int c = 0; c < 7; ++c) a[u][c] = 0.f;
  for (long long j0 = lane; j0 < n; j0 += 64) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long j = j0 + 32 * u;
      if (j < n) {
        const float4 S = s_pos[j], W = s_vel[j];
        const float dx = S.x - P.x, dy = S.y - P.y, dz = S.z - P.z;
        const float du = W.x - V.x, dv = W.y - V.y, dw = W.z - V.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, e2s)));
        const float rv = fmaf(dz, dw, fmaf(dy, dv, dx * du));
        float ri = rsqrt_approx(r2);
        if (!(r2 > 0.f)) ri = 0.f;
        const float qq = ri * ri, mri = S.w * ri, mr3 = mri * qq;
        const float al = (rv * qq) * -3.0f;
        a[u][0] = fmaf(mr3, dx, a[u][0]), a[u][1] = fmaf(mr3, dy, a[u][1]), a[u][2] = fmaf(mr3, dz, a[u][2]);
        a[u][3] = fmaf(mr3, fmaf(al, dx, du), a[u][3]);
        a[u][4] = fmaf(mr3, fmaf(al, dy, dv), a[u][4]);
        a[u][5] = fmaf(mr3, fmaf(al, dz, dw), a[u][5]);
        if (j != t) a[u][6] += mri;
      }
    }
  }
  double v[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    v[c] = (double)a[0][c] + (double)a[1][c];
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
  }
  if (lane == 0) {
    const double sc = (double)scale;
    const double ka = G * sc * sc, kj = G * sc * sc * sc * vel_to_len;
    acc[t] = ka * v[0], acc[n + t] = ka * v[1], acc[2 * n + t] = ka * v[2];
    jerk[t] = kj * v[3], jerk[n + t] = kj * v[4], jerk[2 * n + t] = kj * v[5];
    if (pot) pot[t] = -(G * sc) * v[6];
  }
}

// ------------------------------------------------------------------------------ launcher ----
typedef void (*hermite_fn)(const HermiteParams);
struct HermiteVariant {
  const char* name;
  int np, nw, minb;
  hermite_fn fn[2][2];  // [pot][guard]
};
#define HM_V(NP, NW, UNR, MINB)                                                                                          \
  {                                                                                                                      \
    {hermite_tp_kernel<NP, false, false, NW, UNR, MINB>, hermite_tp_kernel<NP, false, true, NW, UNR, MINB>}, {           \
      hermite_tp_kernel<NP, true, false, NW, UNR, MINB>, hermite_tp_kernel<NP, true, true, NW, UNR, MINB>                \
    }                                                                                                                    \
  }
// Only the production shape is compiled into the shipped library; the others exist in the -DOCG_TUNING build.
#ifdef OCG_TUNING
#define HM_TUNE(...) __VA_ARGS__
#else
#define HM_TUNE(...) {{nullptr, nullptr}, {nullptr, nullptr}}
#endif
#define HM_PRODUCTION 6
static const HermiteVariant g_hm_variants[] = {
    {"np4 8w (2048-target tiles)", 4, 8, 1, HM_TUNE(HM_V(4, 8, 1, 1))},
    {"np3 8w (1536-target tiles)", 3, 8, 1, HM_TUNE(HM_V(3, 8, 1, 1))},
    {"np2 8w (1024-target tiles)", 2, 8, 1, HM_TUNE(HM_V(2, 8, 1, 1))},
    {"np2 12w (1536-target tiles)", 2, 12, 1, HM_TUNE(HM_V(2, 12, 1, 1))},
    {"np3 12w (2304-target tiles)", 3, 12, 1, HM_TUNE(HM_V(3, 12, 1, 1))},
    {"np4 8w unroll 2", 4, 8, 1, HM_TUNE(HM_V(4, 8, 2, 1))},
    {"np1 8w x 2 CTA/SM (512-target tiles)", 1, 8, 2, HM_V(1, 8, 1, 2)},
    {"np1 8w x 2 CTA/SM unroll 2", 1, 8, 2, HM_TUNE(HM_V(1, 8, 2, 2))},
    {"np2 8w unroll 2", 2, 8, 1, HM_TUNE(HM_V(2, 8, 2, 1))},
    {"np1 4w x 3 CTA/SM (256-target tiles)", 1, 4, 3, HM_TUNE(HM_V(1, 4, 1, 3))},
};
#define HM_N_VARIANTS ((int)(sizeof(g_hm_variants) / sizeof(g_hm_variants[0])))
int ocg_hermite_n_variants() { return HM_N_VARIANTS; }
const char* ocg_hermite_variant_name(int v) { return v >= 0 && v < HM_N_VARIANTS ? g_hm_variants[v].name : ""; }

static float hermite_scale(float e2f) {
  // power-of-two length scale that puts eps at ~2^-8 (as K4 does): 1/r^3 <= 2^24 for softened pairs, and
  // separations up to ~1e8 eps stay far from FP32 range limits; eps2 == 0 keeps scale 1 (guarded form)
  if (!(e2f > 0.f)) return 1.0f;
  int e;
  frexpf(sqrtf(e2f), &e);
  float scale = ldexpf(1.0f, -8 - e);
  if (!(scale > 0.f) || !isfinite(scale) || !(e2f * scale * scale > 0.f)) scale = 1.0f;
  return scale;
}

// The streaming path: work plan, scratch, tile pack (resident arrays, or `pack` for a cluster whose stars live on several
// ranks), stream-K kernel.
static int hermite_force_rows(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* mass_dev, int64_t n,
                              const int64_t* seg_offsets_host, int32_t n_seg, double eps2, double G, double vel_to_len,
                              int64_t tgt_begin, int64_t tgt_end, double* acc_dev, double* jerk_dev, double* pot_dev,
                              cudaStream_t st, ocg_hermite_pack_fn pack, void* pack_user) {
  const bool want_pot = pot_dev != nullptr;
  const int NC = want_pot ? 7 : 6;
  const float e2f = (float)eps2;
  const bool guard = !(e2f > 0.f);
  const float scale = hermite_scale(e2f);
  const float e2s = guard ? 0.f : e2f * scale * scale;
  int variant = ctx->knobs.hermite_variant;
  if (variant >= HM_N_VARIANTS || (variant >= 0 && !g_hm_variants[variant].fn[0][0])) variant = -1;
  // measured on B200 (tools/bench_hermite.py, profiles/r01_bench_hermite.json): one target pair per thread, two
  // 8-warp CTAs per SM (122 registers, 16 warps per SM), 4-source-group loop unrolled twice is the fastest shape at
  // N = 65 536 (60.6 % of FP32 peak, kernel 62.2 %) and on batches of 4096-star clusters (59.9 %)
  if (variant < 0) variant = HM_PRODUCTION;
  const HermiteVariant& v = g_hm_variants[variant];
  const int NTHR = 32 * v.nw, CT = 2 * v.np * NTHR;

  const long long grid_ctas = (long long)ctx->sm_count * v.minb;
  OcgClusterRows plan;
  int rc;
  if ((rc = ocg_plan_cluster_rows(ctx, n, seg_offsets_host, n_seg, tgt_begin, tgt_end, CT, HM_TS, grid_ctas, st, &plan, /*which=*/1)))
    return rc;
  float* tiles;
  float4* tgt;
  double* partial;
  unsigned int* tickets;
  // own scratch slots: a CUDA graph that captured this call must survive K4 calls (bound_center_of_mass) that would
  // otherwise grow - i.e. reallocate - a shared buffer; the partials are always sized for the 7-component form
  rc = ocg_scratch(ctx, OCG_SCR_TILES_HM, (size_t)plan.total_tiles * HM_TILE_BYTES, (void**)&tiles);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TGT_HM, 2 * sizeof(float4) * (size_t)n, (void**)&tgt);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_PARTIAL_HM, sizeof(double) * (size_t)plan.n_slots * 7 * (size_t)n, (void**)&partial);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TICKETS_HM, sizeof(unsigned int) * (size_t)(plan.n_rows > 0 ? plan.n_rows : 1), (void**)&tickets, true);
  if (rc) return rc;
  if (pack) {
    if ((rc = pack(ctx, pack_user, scale, plan.total_tiles, tiles, tgt, tgt + n, st))) return rc;
  } else {
    const long long nslots = plan.total_tiles * HM_TS;
    pack_hermite_kernel<<<(int)((nslots + 255) / 256), 256, 0, st>>>(pos_dev, vel_dev, mass_dev, n, plan.d_seg_off,
                                                                    plan.d_seg_tile, n_seg, scale, tiles, tgt, tgt + n);
    OCG_CHECK_LAUNCH(ctx, "pack_hermite_kernel");
  }
  if (plan.n_rows == 0) return OCG_OK;
  HermiteParams p;
  p.tiles = tiles, p.tgt_pos = tgt, p.tgt_vel = tgt + n, p.partial = partial, p.out_stride = n;
  p.sk.rows = plan.d_rows, p.sk.row_prefix = plan.d_prefix, p.sk.n_rows = plan.n_rows, p.sk.n_slots = plan.n_slots;
  p.sk.n_tgt = n, p.sk.ct = CT, p.sk.nst_uniform = nullptr, p.sk.tickets = tickets;
  p.out_acc = acc_dev, p.out_jerk = jerk_dev, p.out_pot = pot_dev, p.G = G, p.vel_to_len = vel_to_len;
  p.e2s = e2s, p.scale = scale;
  hermite_fn fn = v.fn[want_pot][guard];
  const size_t smem = HM_NSTAGE * HM_TILE_BYTES + 128 + (size_t)NC * 2 * v.np * NTHR * sizeof(double);
  OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ctx->timing) OCG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  fn<<<(int)grid_ctas, NTHR, smem, st>>>(p);
  OCG_CHECK_LAUNCH(ctx, "hermite_tp_kernel");
  if (ctx->timing) {
    OCG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = 1;
  }
  return OCG_OK;
}

int ocg_hermite_force_packed(ocg_ctx* ctx, int64_t n, int64_t tgt_begin, int64_t tgt_end, double eps2, double G,
                             double vel_to_len, double* acc_dev, double* jerk_dev, double* pot_dev, cudaStream_t st,
                             ocg_hermite_pack_fn pack, void* user) {
  int64_t seg[2] = {0, n};
  return hermite_force_rows(ctx, nullptr, nullptr, nullptr, n, seg, 1, eps2, G, vel_to_len, tgt_begin, tgt_end, acc_dev, jerk_dev,
                            pot_dev, st, pack, user);
}

extern "C" int ocg_self_gravity_hermite(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* mass_dev,
                                        int64_t n, const int64_t* seg_offsets_host, int32_t n_seg, double eps2, double G,
                                        double vel_to_len, int64_t tgt_begin, int64_t tgt_end, double* acc_dev,
                                        double* jerk_dev, double* pot_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || n_seg < 1) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: n < 0 or n_seg < 1");
  if (n == 0) return OCG_OK;
  if (!pos_dev || !vel_dev || !mass_dev || !acc_dev || !jerk_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: NULL argument");
  if (!(eps2 >= 0.0)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: eps2 = %g must be >= 0", eps2);
  if (tgt_begin < 0 || tgt_end > n || tgt_begin > tgt_end)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: target range [%lld,%lld) outside [0,%lld)",
                    (long long)tgt_begin, (long long)tgt_end, (long long)n);
  int64_t one_seg[2] = {0, n};
  if (!seg_offsets_host) {
    if (n_seg != 1)
      return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: seg_offsets is NULL but n_seg = %d", n_seg);
    seg_offsets_host = one_seg;
  }
  if (seg_offsets_host[0] != 0 || seg_offsets_host[n_seg] != n)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: seg_offsets must run from 0 to n");
  for (int s = 0; s < n_seg; ++s)
    if (seg_offsets_host[s + 1] < seg_offsets_host[s])
      return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite: seg_offsets not monotone at %d", s);
  if (tgt_begin == tgt_end) return OCG_OK;

  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const float e2f = (float)eps2;
  const bool guard = !(e2f > 0.f);
  const float scale = hermite_scale(e2f);
  const float e2s = guard ? 0.f : e2f * scale * scale;

  if (ctx->knobs.hermite_small_path && n_seg == 1 && n <= HM_SMALL_MAX_N) {
    const size_t smem = 2 * sizeof(float4) * (size_t)n;
    OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)hermite_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(2 * sizeof(float4) * HM_SMALL_MAX_N)));
    const long long nb = (tgt_end - tgt_begin + HM_SMALL_WARPS - 1) / HM_SMALL_WARPS;
    hermite_small_kernel<<<(int)nb, 32 * HM_SMALL_WARPS, smem, st>>>(pos_dev, vel_dev, mass_dev, n, e2s, scale, G, vel_to_len,
                                                                    tgt_begin, tgt_end, acc_dev, jerk_dev, pot_dev);
    OCG_CHECK_LAUNCH(ctx, "hermite_small_kernel");
    ctx->ev_valid = 0;
    return OCG_OK;
  }

  return hermite_force_rows(ctx, pos_dev, vel_dev, mass_dev, n, seg_offsets_host, n_seg, eps2, G, vel_to_len, tgt_begin, tgt_end,
                            acc_dev, jerk_dev, pot_dev, st, nullptr, nullptr);
}

// ------------------------------------------------------------------ predictor / corrector ----
// FP64, every multiply and add rounded separately in a fixed order, so that the numpy restatement in the oracle
// reproduces the bits (as K5 does).
//   xp = x + ((v*dt + a*(dt^2/2)) + j*(dt^3/6)) * vel_to_len        vp = v + (a*dt + j*(dt^2/2))
__global__ void hermite_predict_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                       const double* __restrict__ acc, const double* __restrict__ jerk, long long n3,
                                       double c1, double c2, double c3, double vel_to_len, double* __restrict__ pos_p,
                                       double* __restrict__ vel_p) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const double v = vel[i], a = acc[i], j = jerk[i];
  const double dx = __dadd_rn(__dadd_rn(__dmul_rn(v, c1), __dmul_rn(a, c2)), __dmul_rn(j, c3));
  pos_p[i] = __dadd_rn(pos[i], __dmul_rn(dx, vel_to_len));
  vel_p[i] = __dadd_rn(v, __dadd_rn(__dmul_rn(a, c1), __dmul_rn(j, c2)));
}

struct HermiteCoef {
  double k6, k4, k2, i2;     // a2 = ((a0-a1)*(-6) - (j0*4 + j1*2)*dt) * (1/dt^2)
  double k12, k6dt, i3;      // a3 = ((a0-a1)*12 + (j0+j1)*(6 dt)) * (1/dt^3)
  double dt, d3, d4, d5;     // dt, dt^3/6, dt^4/24, dt^5/120
  double vel_to_len, eta;
};

// Corrector + Aarseth step.  One thread per star (the step criterion needs vector norms):
//   v1 = vp + a2*(dt^3/6) + a3*(dt^4/24)          x1 = xp + (a2*(dt^4/24) + a3*(dt^5/120)) * vel_to_len
//   dt_i = sqrt(eta (|a1||a2'| + |j1|^2) / (|j1||a3| + |a2'|^2)),  a2' = a2 + dt a3
// and acc0 <- acc1, jerk0 <- jerk1 (the end-of-step force is the start-of-step force of the next step).
__global__ void hermite_correct_kernel(double* __restrict__ pos, double* __restrict__ vel, double* __restrict__ acc0,
                                       double* __restrict__ jerk0, const double* __restrict__ pos_p,
                                       const double* __restrict__ vel_p, const double* __restrict__ acc1,
                                       const double* __restrict__ jerk1, long long n, HermiteCoef k,
                                       unsigned long long* __restrict__ dt_min_bits) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double dti = __longlong_as_double(0x7ff0000000000000ll);
  if (i < n) {
    double s_a = 0.0, s_j = 0.0, s_2 = 0.0, s_3 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long q = c * n + i;
      const double a0 = acc0[q], j0 = jerk0[q], a1 = acc1[q], j1 = jerk1[q];
      const double da = __dadd_rn(a0, -a1);
      const double a2 = __dmul_rn(__dadd_rn(__dmul_rn(da, k.k6), -__dmul_rn(__dadd_rn(__dmul_rn(j0, k.k4), __dmul_rn(j1, k.k2)), k.dt)), k.i2);
      const double a3 = __dmul_rn(__dadd_rn(__dmul_rn(da, k.k12), __dmul_rn(__dadd_rn(j0, j1), k.k6dt)), k.i3);
      vel[q] = __dadd_rn(__dadd_rn(vel_p[q], __dmul_rn(a2, k.d3)), __dmul_rn(a3, k.d4));
      pos[q] = __dadd_rn(pos_p[q], __dmul_rn(__dadd_rn(__dmul_rn(a2, k.d4), __dmul_rn(a3, k.d5)), k.vel_to_len));
      acc0[q] = a1, jerk0[q] = j1;
      const double a2e = a2 + k.dt * a3;
      s_a += a1 * a1, s_j += j1 * j1, s_2 += a2e * a2e, s_3 += a3 * a3;
    }
    const double num = sqrt(s_a * s_2) + s_j, den = sqrt(s_j * s_3) + s_2;
    if (den > 0.0 && num > 0.0) dti = sqrt(k.eta * num / den);
  }
  if (dt_min_bits) {
    // positive doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(dti);
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
      b = other < b ? other : b;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(dt_min_bits, b);
  }
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

extern "C" int ocg_hermite_predict(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* acc_dev,
                                   const double* jerk_dev, int64_t n, double dt, double vel_to_len, double* pos_pred_dev,
                                   double* vel_pred_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || (n > 0 && (!pos_dev || !vel_dev || !acc_dev || !jerk_dev || !pos_pred_dev || !vel_pred_dev)))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_predict: bad arguments");
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  const double c2 = dt * dt * 0.5, c3 = dt * dt * dt / 6.0;
  hermite_predict_kernel<<<(int)((3 * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos_dev, vel_dev, acc_dev, jerk_dev, 3 * n,
                                                                                      dt, c2, c3, vel_to_len, pos_pred_dev,
                                                                                      vel_pred_dev);
  OCG_CHECK_LAUNCH(ctx, "hermite_predict_kernel");
  return OCG_OK;
}

extern "C" int ocg_hermite_correct(ocg_ctx* ctx, double* pos_dev, double* vel_dev, double* acc0_dev, double* jerk0_dev,
                                   const double* pos_pred_dev, const double* vel_pred_dev, const double* acc1_dev,
                                   const double* jerk1_dev, int64_t n, double dt, double vel_to_len, double eta,
                                   double* dt_min_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || (n > 0 && (!pos_dev || !vel_dev || !acc0_dev || !jerk0_dev || !pos_pred_dev || !vel_pred_dev || !acc1_dev ||
                          !jerk1_dev)))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_correct: bad arguments");
  if (!(dt != 0.0)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_correct: dt must be non-zero");
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  HermiteCoef k;
  const double dt2 = dt * dt, dt3 = dt * dt * dt;
  k.k6 = -6.0, k.k4 = 4.0, k.k2 = 2.0, k.i2 = 1.0 / dt2;
  k.k12 = 12.0, k.k6dt = 6.0 * dt, k.i3 = 1.0 / dt3;
  k.dt = dt, k.d3 = dt3 / 6.0, k.d4 = dt2 * dt2 / 24.0, k.d5 = dt2 * dt3 / 120.0;
  k.vel_to_len = vel_to_len, k.eta = eta;
  if (dt_min_dev) {
    set_u64_kernel<<<1, 1, 0, st>>>((unsigned long long*)dt_min_dev, 0x7ff0000000000000ull);
    OCG_CHECK_LAUNCH(ctx, "set_u64_kernel");
  }
  hermite_correct_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(pos_dev, vel_dev, acc0_dev, jerk0_dev, pos_pred_dev,
                                                                vel_pred_dev, acc1_dev, jerk1_dev, n, k,
                                                                (unsigned long long*)dt_min_dev);
  OCG_CHECK_LAUNCH(ctx, "hermite_correct_kernel");
  return OCG_OK;
}

// =====================================================================================================
// ph4's individual block time steps (oc_code.py:218-229 instantiates AMUSE ph4; options.py:248-253).
//
// Every star i carries its own time t_i and step dt_i = span / 2^k_i; all steps are powers of two of one span and every
// star's time is a multiple of its step, so the stars synchronise at the end of the span.  One block step:
//   t_next  = min_i (t_i + dt_i);  active = { i : t_i + dt_i == t_next }        (times are integer ticks: exact)
//   predict EVERY star to t_next with its own elapsed time (Makino & Aarseth 1992 predictor)
//   force (acc + jerk) on the ACTIVE stars from ALL predicted stars              (K6, active targets x all sources)
//   correct the active stars over their own dt_i, set t_i = t_next, choose the next dt_i from the Aarseth criterion:
//   halve while it exceeds the criterion; double (once) when the criterion allows 2 dt_i and t_next is a multiple of 2 dt_i.
// The host loop reads back 16 bytes per block step (t_next, number of active stars) to size the launches.
#define HB_LEVELS 20                      /* the span is 2^20 ticks: smallest step = span / 1 048 576 */
#define HB_SPAN_TICKS (1ll << HB_LEVELS)

struct HermiteBlockSel {
  long long t_next;
  int n_active;
  int pad;
};

// first force done: per-star starting step dt = eta * |a| / |j| (Aarseth 1985), rounded down to a power-of-two fraction of
// the span, at most the span and at most 2^max_level_up... (min_ticks = smallest allowed step)
__global__ void hermite_block_init_kernel(const double* __restrict__ acc, const double* __restrict__ jerk, long long n, double eta,
                                          double tick_len, long long min_ticks, long long* __restrict__ t_tick,
                                          long long* __restrict__ dt_tick) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a2 = 0.0, j2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) a2 += acc[c * n + i] * acc[c * n + i], j2 += jerk[c * n + i] * jerk[c * n + i];
  long long ticks = HB_SPAN_TICKS;
  if (j2 > 0.0) {
    const double dt = eta * sqrt(a2 / j2);
    while (ticks > min_ticks && (double)ticks * tick_len > dt) ticks >>= 1;
  }
  t_tick[i] = 0;
  dt_tick[i] = ticks;
}

// one CTA: t_next = min (t + dt), then the stable list of the stars that reach it
__global__ void __launch_bounds__(1024) hermite_block_select_kernel(const long long* __restrict__ t_tick, const long long* __restrict__ dt_tick,
                                                                    long long n, int* __restrict__ active_idx, HermiteBlockSel* __restrict__ sel) {
  __shared__ long long s_min[32];
  __shared__ int wtot[32];
  __shared__ long long s_tnext;
  __shared__ int s_base;
  long long m = 0x7fffffffffffffffll;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const long long v = t_tick[i] + dt_tick[i];
    m = v < m ? v : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other < m ? other : m;
  }
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long v = s_min[0];
    for (int w = 1; w < 32; ++w) v = s_min[w] < v ? s_min[w] : v;
    s_tnext = v;
    s_base = 0;
  }
  __syncthreads();
  const long long tn = s_tnext;
  for (long long start = 0; start < n; start += 1024) {
    const long long i = start + threadIdx.x;
    const bool k = i < n && t_tick[i] + dt_tick[i] == tn;
    const unsigned b = __ballot_sync(0xffffffffu, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wtot[warp] = __popc(b);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < 32; ++w) {
      const int t = wtot[w];
      if (w < warp) before += t;
      all += t;
    }
    if (k) active_idx[s_base + before + __popc(b & ((1u << lane) - 1u))] = (int)i;
    __syncthreads();
    if (threadIdx.x == 0) s_base += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) sel->t_next = tn, sel->n_active = s_base, sel->pad = 0;
}

// predictor with per-star elapsed time: the operation order of hermite_predict_kernel, delta = (t_next - t_i) * tick_len
__global__ void hermite_block_predict_kernel(const double* __restrict__ pos, const double* __restrict__ vel, const double* __restrict__ acc,
                                             const double* __restrict__ jerk, const long long* __restrict__ t_tick, long long n,
                                             const HermiteBlockSel* __restrict__ sel, double tick_len, double vel_to_len,
                                             double* __restrict__ pos_p, double* __restrict__ vel_p) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= 3 * n) return;
  const long long i = q % n;
  const double dt = (double)(sel->t_next - t_tick[i]) * tick_len;
  const double c2 = dt * dt * 0.5, c3 = dt * dt * dt / 6.0;
  const double v = vel[q], a = acc[q], j = jerk[q];
  const double dx = __dadd_rn(__dadd_rn(__dmul_rn(v, dt), __dmul_rn(a, c2)), __dmul_rn(j, c3));
  pos_p[q] = __dadd_rn(pos[q], __dmul_rn(dx, vel_to_len));
  vel_p[q] = __dadd_rn(v, __dadd_rn(__dmul_rn(a, dt), __dmul_rn(j, c2)));
}

// compact float4 targets of the active stars out of the packed full arrays
__global__ void hermite_block_gather_kernel(const float4* __restrict__ tgt_pos, const float4* __restrict__ tgt_vel,
                                            const int* __restrict__ active_idx, int n_active, float4* __restrict__ out_pos,
                                            float4* __restrict__ out_vel) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_active) return;
  const int i = active_idx[k];
  out_pos[k] = tgt_pos[i], out_vel[k] = tgt_vel[i];
}

// corrector of the active stars (operation order of hermite_correct_kernel) + their next step
__global__ void hermite_block_correct_kernel(double* __restrict__ pos, double* __restrict__ vel, double* __restrict__ acc0,
                                             double* __restrict__ jerk0, const double* __restrict__ pos_p, const double* __restrict__ vel_p,
                                             const double* __restrict__ acc1c, const double* __restrict__ jerk1c,
                                             const int* __restrict__ active_idx, int n_active, long long n,
                                             const HermiteBlockSel* __restrict__ sel, double tick_len, double vel_to_len, double eta,
                                             long long min_ticks, long long* __restrict__ t_tick, long long* __restrict__ dt_tick) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_active) return;
  const long long i = active_idx[k];
  const long long ticks = dt_tick[i], tn = sel->t_next;
  const double dt = (double)ticks * tick_len;
  const double dt2 = dt * dt, dt3 = dt2 * dt, i2 = 1.0 / dt2, i3 = 1.0 / dt3;
  const double d3 = dt3 / 6.0, d4 = dt2 * dt2 / 24.0, d5 = dt2 * dt3 / 120.0, k6dt = 6.0 * dt;
  double s_a = 0.0, s_j = 0.0, s_2 = 0.0, s_3 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long long q = c * n + i;
    const double a0 = acc0[q], j0 = jerk0[q], a1 = acc1c[(long long)c * n_active + k], j1 = jerk1c[(long long)c * n_active + k];
    const double da = __dadd_rn(a0, -a1);
    const double a2 = __dmul_rn(__dadd_rn(__dmul_rn(da, -6.0), -__dmul_rn(__dadd_rn(__dmul_rn(j0, 4.0), __dmul_rn(j1, 2.0)), dt)), i2);
    const double a3 = __dmul_rn(__dadd_rn(__dmul_rn(da, 12.0), __dmul_rn(__dadd_rn(j0, j1), k6dt)), i3);
    vel[q] = __dadd_rn(__dadd_rn(vel_p[q], __dmul_rn(a2, d3)), __dmul_rn(a3, d4));
    pos[q] = __dadd_rn(pos_p[q], __dmul_rn(__dadd_rn(__dmul_rn(a2, d4), __dmul_rn(a3, d5)), vel_to_len));
    acc0[q] = a1, jerk0[q] = j1;
    const double a2e = a2 + dt * a3;
    s_a += a1 * a1, s_j += j1 * j1, s_2 += a2e * a2e, s_3 += a3 * a3;
  }
  const double num = sqrt(s_a * s_2) + s_j, den = sqrt(s_j * s_3) + s_2;
  long long next = ticks;
  if (den > 0.0 && num > 0.0) {
    const double want = sqrt(eta * num / den);
    if (want < dt) {
      while (next > min_ticks && (double)next * tick_len > want) next >>= 1;
    } else if (want >= 2.0 * dt && 2 * next <= HB_SPAN_TICKS && tn % (2 * next) == 0) {
      next <<= 1;
    }
  }
  t_tick[i] = tn;
  dt_tick[i] = next;
}

extern "C" int ocg_hermite_block_evolve(ocg_ctx* ctx, double* pos_dev, double* vel_dev, const double* mass_dev, double* acc_dev,
                                        double* jerk_dev, int64_t n, double eps2, double G, double vel_to_len, double span, double eta,
                                        int32_t max_level, int64_t* n_block_steps_host, int64_t* n_star_steps_host, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 1 || !pos_dev || !vel_dev || !mass_dev || !acc_dev || !jerk_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_block_evolve: bad arguments");
  if (!(span > 0.0) || !(eta > 0.0) || !(eps2 >= 0.0) || max_level < 0 || max_level > HB_LEVELS)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_block_evolve: span %g, eta %g, eps2 %g, max_level %d (0..%d)", span, eta, eps2,
                    max_level, HB_LEVELS);
  if (n > 0x7fffffffll) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_hermite_block_evolve: too many stars");
  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const double tick_len = span / (double)HB_SPAN_TICKS;
  const long long min_ticks = HB_SPAN_TICKS >> max_level;
  const float e2f = (float)eps2;
  const bool guard = !(e2f > 0.f);
  const float scale = hermite_scale(e2f), e2s = guard ? 0.f : e2f * scale * scale;
  // ---- workspace: predicted state, per-star clocks, the active list, compact targets and forces
  const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_posp = carve(3 * n8 * 8), o_velp = carve(3 * n8 * 8), o_t = carve(n8 * 8), o_dt = carve(n8 * 8), o_idx = carve(n8 * 4),
               o_acc1 = carve(3 * n8 * 8), o_jerk1 = carve(3 * n8 * 8), o_tp = carve(n8 * 16), o_tv = carve(n8 * 16), o_sel = carve(256),
               o_seg = carve(256);
  char* ws;
  int rc = ocg_scratch(ctx, OCG_SCR_BLOCK, off, (void**)&ws);
  if (rc) return rc;
  double *pos_p = (double*)(ws + o_posp), *vel_p = (double*)(ws + o_velp), *acc1 = (double*)(ws + o_acc1), *jerk1 = (double*)(ws + o_jerk1);
  long long *t_tick = (long long*)(ws + o_t), *dt_tick = (long long*)(ws + o_dt), *d_seg = (long long*)(ws + o_seg);
  int* active = (int*)(ws + o_idx);
  float4 *ctgt_pos = (float4*)(ws + o_tp), *ctgt_vel = (float4*)(ws + o_tv);
  HermiteBlockSel* d_sel = (HermiteBlockSel*)(ws + o_sel);
  const long long total_tiles = (n + HM_TS - 1) / HM_TS;
  float* tiles;
  float4* tgt;
  double* partial;
  unsigned int* tickets;
  // shape: the production shape of the force kernel; rows in uniform mode (every row streams all tiles), no plan upload
  int variant = ctx->knobs.hermite_variant;
  if (variant >= HM_N_VARIANTS || variant < 0 || !g_hm_variants[variant].fn[0][0]) variant = HM_PRODUCTION;
  const HermiteVariant& v = g_hm_variants[variant];
  const int NTHR = 32 * v.nw, CT = 2 * v.np * NTHR;
  const long long grid_ctas = (long long)ctx->sm_count * v.minb;
  const long long max_rows = (n + CT - 1) / CT;
  const long long n_slots = grid_ctas;  // a single row (few active stars) may be shared by every CTA; the slots are indexed
                                        // with stride n_t, and sharers * n_t <= (CTAs + rows) * CT whatever n_t is
  rc = ocg_scratch(ctx, OCG_SCR_TILES_HM, (size_t)total_tiles * HM_TILE_BYTES, (void**)&tiles);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TGT_HM, 2 * sizeof(float4) * (size_t)n, (void**)&tgt);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_PARTIAL_HM, sizeof(double) * 7 * (size_t)(grid_ctas + max_rows + 1) * CT, (void**)&partial);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TICKETS_HM, sizeof(unsigned int) * (size_t)(max_rows > 0 ? max_rows : 1), (void**)&tickets, true);
  if (rc) return rc;
  const long long h_seg[4] = {0, total_tiles, 0, n};  // seg_tile[2] | seg_off[2]
  OCG_CUDA(ctx, cudaMemcpyAsync(d_seg, h_seg, sizeof(h_seg), cudaMemcpyHostToDevice, st));
  OCG_CUDA(ctx, cudaStreamSynchronize(st));

  hermite_fn fn = v.fn[0][guard];
  const size_t smem = HM_NSTAGE * HM_TILE_BYTES + 128 + (size_t)6 * 2 * v.np * NTHR * sizeof(double);
  OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // force on targets [0, n_t) of (tp, tv) from the sources (src_pos, src_vel): compact outputs [3][n_t]
  auto force = [&](const double* src_pos, const double* src_vel, const int* idx, int n_t, double* out_acc, double* out_jerk) -> int {
    const long long nslots = total_tiles * HM_TS;
    pack_hermite_kernel<<<(int)((nslots + 255) / 256), 256, 0, st>>>(src_pos, src_vel, mass_dev, n, d_seg + 2, d_seg, 1, scale, tiles, tgt, tgt + n);
    OCG_CHECK_LAUNCH(ctx, "pack_hermite_kernel");
    const float4 *tp = tgt, *tv = tgt + n;
    if (idx) {
      hermite_block_gather_kernel<<<(n_t + 255) / 256, 256, 0, st>>>(tgt, tgt + n, idx, n_t, ctgt_pos, ctgt_vel);
      OCG_CHECK_LAUNCH(ctx, "hermite_block_gather_kernel");
      tp = ctgt_pos, tv = ctgt_vel;
    }
    HermiteParams p{};
    p.tiles = tiles, p.tgt_pos = tp, p.tgt_vel = tv, p.partial = partial, p.out_stride = n_t;
    p.sk.rows = nullptr, p.sk.row_prefix = nullptr, p.sk.n_rows = (n_t + CT - 1) / CT, p.sk.n_slots = (int)n_slots;
    p.sk.n_tgt = n_t, p.sk.ct = CT, p.sk.nst_uniform = nullptr, p.sk.nst_value = (int)total_tiles, p.sk.tickets = tickets;
    p.out_acc = out_acc, p.out_jerk = out_jerk, p.out_pot = nullptr, p.G = G, p.vel_to_len = vel_to_len;
    p.e2s = e2s, p.scale = scale;
    fn<<<(int)grid_ctas, NTHR, smem, st>>>(p);
    OCG_CHECK_LAUNCH(ctx, "hermite_tp_kernel");
    return OCG_OK;
  };

  // ---- start: force at the current state for every star (the BRIDGE kick has just changed the velocities), first steps
  if ((rc = force(pos_dev, vel_dev, nullptr, (int)n, acc_dev, jerk_dev))) return rc;
  hermite_block_init_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(acc_dev, jerk_dev, n, eta, tick_len, min_ticks, t_tick, dt_tick);
  OCG_CHECK_LAUNCH(ctx, "hermite_block_init_kernel");
  long long steps = 0, star_steps = 0;
  for (;;) {
    hermite_block_select_kernel<<<1, 1024, 0, st>>>(t_tick, dt_tick, n, active, d_sel);
    OCG_CHECK_LAUNCH(ctx, "hermite_block_select_kernel");
    HermiteBlockSel h;
    OCG_CUDA(ctx, cudaMemcpyAsync(&h, d_sel, sizeof(h), cudaMemcpyDeviceToHost, st));
    OCG_CUDA(ctx, cudaStreamSynchronize(st));
    if (h.n_active <= 0 || h.t_next > HB_SPAN_TICKS)
      return ocg_fail(ctx, OCG_ERR_CUDA, "ocg_hermite_block_evolve: block schedule broke (t_next %lld, %d active)", h.t_next, h.n_active);
    hermite_block_predict_kernel<<<(int)((3 * n + 255) / 256), 256, 0, st>>>(pos_dev, vel_dev, acc_dev, jerk_dev, t_tick, n, d_sel, tick_len,
                                                                            vel_to_len, pos_p, vel_p);
    OCG_CHECK_LAUNCH(ctx, "hermite_block_predict_kernel");
    if ((rc = force(pos_p, vel_p, active, h.n_active, acc1, jerk1))) return rc;
    hermite_block_correct_kernel<<<(h.n_active + 255) / 256, 256, 0, st>>>(pos_dev, vel_dev, acc_dev, jerk_dev, pos_p, vel_p, acc1, jerk1,
                                                                          active, h.n_active, n, d_sel, tick_len, vel_to_len, eta,
                                                                          min_ticks, t_tick, dt_tick);
    OCG_CHECK_LAUNCH(ctx, "hermite_block_correct_kernel");
    ++steps, star_steps += h.n_active;
    if (h.t_next == HB_SPAN_TICKS) break;  // every step divides the span: all stars arrive here together
  }
  if (n_block_steps_host) *n_block_steps_host = steps;
  if (n_star_steps_host) *n_star_steps_host = star_steps;
  return OCG_OK;
}

// K1 / K4: softened pairwise gravity as a direct sum, hand-written for sm_100a.
//
// Replaces the field-build hot loop of the reference, ConstructKDTree + GetAccelParallel
// (gizmo_interface.py:561,564,566 — a theta=0.5 monopole tree in un-vendored pykdgrav), with the
// exact theta->0 sum, and the ph4 force loop behind oc_code.py:218-229.
//
// Design:
//  * sources are re-laid once per call into component-major tiles of OCG_TS sources — plain
//    (x|y|z|m|e2) or mass-folded (w*x|w*y|w*z|w|w^2*e2, w = (m/M0)^-1/2, the w array pair-swizzled) — and
//    split into a FAST set (every target sees them as Plummer(e2) or pure Newtonian) and a NEAR set
//    (spline sources whose support can reach the target box; singular e2 == 0 sources inside it; sources
//    inside the precision radius), the latter evaluated pair by pair in FP64;
//  * the fast kernel streams tiles through a 4-stage shared-memory ring filled by one TMA bulk copy per
//    tile (cp.async.bulk + mbarrier), issued in line by thread 0 of the CTA (no producer warp);
//  * the inner loop is packed FP32 with the two lanes of an instruction holding two TARGETS and the source
//    a broadcast operand: 11 (mass-folded) or 12 FMA-pipe instructions + 2 MUFU.RSQ per 2 interactions;
//    FP32 partial sums live for one run of FOLD sources and are then folded into FP64 per-target
//    accumulators in shared memory;
//  * work = (target tile x source chunk) items, statically strided over a persistent grid of one CTA per
//    SM; chunk partials are summed in fixed order by a finish kernel => run-to-run deterministic.
#include "ocg_internal.cuh"

#include <math.h>

#include "direct_kernel.cuh"

// ---------------------------------------------------------------- source classification ----
// Every source goes to exactly one of two sets:
//   FAST : handled by the streaming FP32 kernel as Plummer(e2) or pure Newtonian (e2 = 0);
//   NEAR : handled pair by pair in FP64 by near_sum_kernel.
// A source is NEAR when (a) correctness demands it — a spline source whose support h can reach the
// target box, or a source whose r^6 could leave the FP32 range for some target — or (b) accuracy
// profits from it — it lies within the "precision radius" D of the target box, where single pair terms
// are large compared with the net field and FP32 rounding of them would dominate the error budget.
// D is the largest of 4, 2.8, 2, 1.4, ... x (box extent) for which the NEAR set stays below a small cap
// (so the FP64 work is <~1% of the total); it is chosen on the device from a distance histogram.
//
// All distances here are in SCALED units: lengths x scale, scale = 2^k with the largest extent of the
// target box mapped into (0.5, 1].  misc scratch layout (32 ints):
//   [0..5]  target bbox as ordered ints: min xyz, max xyz (unscaled)
//   [8] n_fast  [9] n_near  [10] n_fast_tiles  [11] scale (float bits)  [12] D^2 scaled (float bits)
//   [13] M0 (float bits): power of two >= the largest source mass (mass-folded tiles, see tile_tpair MF)
//   [16..31] histogram: hist[b] = #sources with scaled box distance < 4 * 2^(-b/2)
#define MISC_BBOX 0
#define MISC_NFAST 8
#define MISC_NNEAR 9
#define MISC_NFAST_TILES 10
#define MISC_SCALE 11
#define MISC_D2 12
#define MISC_M0 13
#define MISC_HIST 16
#define MISC_NBINS 16
#define MISC_INTS 32

#define R2_MIN_SCALED 1e-12f  /* r^6 >= 1e-36 stays a normal FP32 number */
#define R2_MAX_FOLDED 1e12f   /* mass-folded r'^6 <= 1e36 stays finite */

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__global__ void misc_init_kernel(int* misc) {
  int i = threadIdx.x;
  if (i < 3) misc[MISC_BBOX + i] = 0x7fffffff;            // +max ordered
  else if (i < 6) misc[MISC_BBOX + i] = (int)0x80000000;  // -max ordered
  else if (i < MISC_INTS) misc[i] = 0;
}

__global__ void bbox_kernel(const float4* __restrict__ tgt, long long n, int* misc) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float4 T = tgt[i];
    mn[0] = fminf(mn[0], T.x), mn[1] = fminf(mn[1], T.y), mn[2] = fminf(mn[2], T.z);
    mx[0] = fmaxf(mx[0], T.x), mx[1] = fmaxf(mx[1], T.y), mx[2] = fmaxf(mx[2], T.z);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicMin(&misc[MISC_BBOX + c], float_to_ordered(mn[c]));
      atomicMax(&misc[MISC_BBOX + 3 + c], float_to_ordered(mx[c]));
    }
  }
}

// scale = 2^k such that the largest box extent lands in (0.5, 1]; a degenerate box (one target, or
// all targets coincident) falls back to the sources' typical distance via scale_hint (host, from eps).
__global__ void scale_kernel(int* misc, float fallback_len) {
  float ext = 0.f;
  for (int c = 0; c < 3; ++c)
    ext = fmaxf(ext, ordered_to_float(misc[MISC_BBOX + 3 + c]) - ordered_to_float(misc[MISC_BBOX + c]));
  if (!(ext > 0.f) || !isfinite(ext)) ext = fallback_len;
  int e;
  frexpf(ext, &e);  // ext = f * 2^e, f in [0.5, 1)
  float sc = ldexpf(1.0f, -e);
  if (!(sc > 0.f) || !isfinite(sc)) sc = 1.0f;
  misc[MISC_SCALE] = __float_as_int(sc);
}

// scaled squared distance from a source to the target box
__device__ __forceinline__ float box_dist2_scaled(float x, float y, float z, const int* misc, float sc) {
  float bx0 = ordered_to_float(misc[MISC_BBOX + 0]), by0 = ordered_to_float(misc[MISC_BBOX + 1]);
  float bz0 = ordered_to_float(misc[MISC_BBOX + 2]), bx1 = ordered_to_float(misc[MISC_BBOX + 3]);
  float by1 = ordered_to_float(misc[MISC_BBOX + 4]), bz1 = ordered_to_float(misc[MISC_BBOX + 5]);
  float dx = fmaxf(fmaxf(bx0 - x, x - bx1), 0.f) * sc;
  float dy = fmaxf(fmaxf(by0 - y, y - by1), 0.f) * sc;
  float dz = fmaxf(fmaxf(bz0 - z, z - bz1), 0.f) * sc;
  return dx * dx + dy * dy + dz * dz;
}

// scaled squared distance from a source to the farthest corner of the target box
__device__ __forceinline__ float box_far2_scaled(float x, float y, float z, const int* misc, float sc) {
  float bx0 = ordered_to_float(misc[MISC_BBOX + 0]), by0 = ordered_to_float(misc[MISC_BBOX + 1]);
  float bz0 = ordered_to_float(misc[MISC_BBOX + 2]), bx1 = ordered_to_float(misc[MISC_BBOX + 3]);
  float by1 = ordered_to_float(misc[MISC_BBOX + 4]), bz1 = ordered_to_float(misc[MISC_BBOX + 5]);
  float dx = fmaxf(fabsf(x - bx0), fabsf(x - bx1)) * sc;
  float dy = fmaxf(fabsf(y - by0), fabsf(y - by1)) * sc;
  float dz = fmaxf(fabsf(z - bz0), fabsf(z - bz1)) * sc;
  return dx * dx + dy * dy + dz * dz;
}

// (a)-type criterion only: must this source be handled pair by pair?
__device__ __forceinline__ bool source_must_be_near(float d2, float soft_scaled, int kernel) {
  const float e2 = soft_scaled * soft_scaled;
  if (kernel == OCG_KERNEL_SPLINE) return !(d2 > R2_MIN_SCALED && d2 > e2 * 1.00001f);
  return !(e2 > R2_MIN_SCALED || d2 > R2_MIN_SCALED);
}

// mass-folded tiles only: w^2 = M0/m >= 1 multiplies every r^2 the kernel forms; the source must have a
// positive mass and its folded r'^6 must stay finite for the farthest target (the near side is covered by
// source_must_be_near because w^2 >= 1).
__device__ __forceinline__ bool source_unfoldable(float x, float y, float z, float m, float soft_scaled, int kernel,
                                                  const int* misc, float sc) {
  if (m == 0.f) return false;  // massless: stored as a null record (w = 0), contributes exactly 0
  if (!(m > 0.f)) return true; // negative or NaN mass: FP64 pair path
  const float w2 = __int_as_float(misc[MISC_M0]) / m;
  const float e2 = kernel == OCG_KERNEL_PLUMMER ? soft_scaled * soft_scaled : 0.f;
  return !(w2 * (box_far2_scaled(x, y, z, misc, sc) + e2) < R2_MAX_FOLDED);
}

#define CLS_BLOCK 1024

// distance histogram of the sources that are not already NEAR by criterion (a)
// Also reduces the largest finite positive source mass into misc[MISC_M0] (positive floats order as ints).
__global__ void __launch_bounds__(CLS_BLOCK) classify_hist_kernel(
    const float4* __restrict__ src, const float* __restrict__ soft, long long n, int kernel, int* misc) {
  __shared__ int h[MISC_NBINS];
  __shared__ int mmax;
  if (threadIdx.x < MISC_NBINS) h[threadIdx.x] = 0;
  if (threadIdx.x == 0) mmax = 0;
  __syncthreads();
  const float sc = __int_as_float(misc[MISC_SCALE]);
  long long i = blockIdx.x * (long long)CLS_BLOCK + threadIdx.x;
  int mbits = 0;
  if (i < n) {
    float4 S = src[i];
    if (S.w > 0.f && isfinite(S.w)) mbits = __float_as_int(S.w);
    float d2 = box_dist2_scaled(S.x, S.y, S.z, misc, sc);
    if (!source_must_be_near(d2, (soft ? soft[i] : 0.f) * sc, kernel)) {
      // every b with d < 4 * 2^(-b/2)  <=>  d2 < 16 * 2^-b
      float lim = 16.0f;
      for (int b = 0; b < MISC_NBINS && d2 < lim; ++b, lim *= 0.5f) atomicAdd(&h[b], 1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mbits = max(mbits, __shfl_xor_sync(0xffffffffu, mbits, o));
  if ((threadIdx.x & 31) == 0 && mbits) atomicMax(&mmax, mbits);
  __syncthreads();
  if (threadIdx.x < MISC_NBINS && h[threadIdx.x]) atomicAdd(&misc[MISC_HIST + threadIdx.x], h[threadIdx.x]);
  if (threadIdx.x == 0 && mmax) atomicMax(&misc[MISC_M0], mmax);
}

// precise == 0: no precision radius (criterion (a) only).  Rounds M0 up to a power of two.
__global__ void choose_radius_kernel(int* misc, int cap, int precise) {
  float d2 = 0.f, lim = 16.0f;
  for (int b = 0; precise && b < MISC_NBINS; ++b, lim *= 0.5f)
    if (misc[MISC_HIST + b] <= cap) {
      d2 = lim;
      break;
    }
  misc[MISC_D2] = __float_as_int(d2);
  float m0 = __int_as_float(misc[MISC_M0]);
  if (!(m0 > 0.f)) m0 = 1.f;
  int e;
  const float f = frexpf(m0, &e);  // m0 = f * 2^e, f in [0.5, 1)
  m0 = ldexpf(1.0f, f > 0.5f ? e : e - 1);
  if (!(m0 > 0.f) || !isfinite(m0)) m0 = 1.f;
  misc[MISC_M0] = __float_as_int(m0);
}

__device__ __forceinline__ bool source_is_fast(float x, float y, float z, float m, float soft, int kernel,
                                               const int* misc, int mf) {
  const float sc = __int_as_float(misc[MISC_SCALE]);
  const float d2 = box_dist2_scaled(x, y, z, misc, sc);
  if (source_must_be_near(d2, soft * sc, kernel)) return false;
  if (mf && source_unfoldable(x, y, z, m, soft * sc, kernel, misc, sc)) return false;
  return !(d2 < __int_as_float(misc[MISC_D2]));
}

__global__ void __launch_bounds__(CLS_BLOCK) classify_count_kernel(
    const float4* __restrict__ src, const float* __restrict__ soft, long long n, int kernel,
    const int* __restrict__ misc, int* __restrict__ counts, int mf) {
  __shared__ int wsum[CLS_BLOCK / 32];
  long long i = blockIdx.x * (long long)CLS_BLOCK + threadIdx.x;
  bool fast = false;
  if (i < n) {
    float4 S = src[i];
    fast = source_is_fast(S.x, S.y, S.z, S.w, soft ? soft[i] : 0.f, kernel, misc, mf);
  }
  unsigned b = __ballot_sync(0xffffffffu, fast);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) counts[blockIdx.x] = v;
  }
}

// Single-block exclusive scan of per-block fast counts (in place); totals to misc.
__global__ void __launch_bounds__(1024) classify_scan_kernel(int* counts, int nblocks, long long n, int* misc) {
  __shared__ long long carry;
  __shared__ int wtot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? counts[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wtot[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += u;
      }
      wtot[threadIdx.x] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    long long c = carry;
    int excl = incl - v + wtot[threadIdx.x >> 5];
    // ints suffice: n_src < 2^31 per call (checked on the host)
    if (i < nblocks) counts[i] = (int)(c + excl);
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    long long nf = carry;
    misc[MISC_NFAST] = (int)nf;
    misc[MISC_NNEAR] = (int)(n - nf);
    misc[MISC_NFAST_TILES] = (int)((nf + OCG_TS - 1) / OCG_TS);
  }
}

// Stable scatter: fast sources -> scaled tiles, near sources -> near list (float4 xyzm + float soft).
__global__ void __launch_bounds__(CLS_BLOCK) classify_scatter_kernel(
    const float4* __restrict__ src, const float* __restrict__ soft, long long n, int kernel,
    const int* __restrict__ misc, const int* __restrict__ fast_off, float* __restrict__ tiles,
    float4* __restrict__ near_xyzm, float* __restrict__ near_soft, int mf, int narr) {
  __shared__ int wbase[CLS_BLOCK / 32];
  long long i = blockIdx.x * (long long)CLS_BLOCK + threadIdx.x;
  bool valid = i < n, fast = false;
  float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
  float h = 0.f;
  if (valid) {
    S = src[i];
    h = soft ? soft[i] : 0.f;
    fast = source_is_fast(S.x, S.y, S.z, S.w, h, kernel, misc, mf);
  }
  unsigned b = __ballot_sync(0xffffffffu, fast);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wbase[w] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wbase[threadIdx.x], incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (threadIdx.x >= o) incl += u;
    }
    wbase[threadIdx.x] = incl - v;
  }
  __syncthreads();
  if (!valid) return;
  int rank_fast = wbase[w] + __popc(b & ((1u << lane) - 1u));
  long long block_start = blockIdx.x * (long long)CLS_BLOCK;
  long long foff = fast_off[blockIdx.x];
  if (fast) {
    const float sc = __int_as_float(misc[MISC_SCALE]);
    long long pos = foff + rank_fast;
    long long tile = pos / OCG_TS;
    int j = (int)(pos - tile * OCG_TS);
    float* T = tiles + tile * (long long)narr * OCG_TS;
    // spline sources that reach the fast set are Newtonian for every target: e2 = 0;
    // Plummer sources whose scaled e2 is negligible are far enough (d2 > R2_MIN) to drop it too
    float hs = h * sc;
    float e2 = kernel == OCG_KERNEL_PLUMMER ? hs * hs : 0.f;
    if (mf) {
      // mass-folded record: w = (m/M0)^-1/2 (M0 a power of two: m/M0 is exact), coordinates and e2 carry it
      const bool massless = S.w == 0.f;  // null record, same as the tile padding: d' = 0, y3 = 1, d'*y3 = 0
      const float w = massless ? 0.f : (float)rsqrt((double)(S.w / __int_as_float(misc[MISC_M0])));
      T[j] = massless ? 0.f : (S.x * sc) * w;
      T[OCG_TS + j] = massless ? 0.f : (S.y * sc) * w;
      T[2 * OCG_TS + j] = massless ? 0.f : (S.z * sc) * w;
      T[3 * OCG_TS + (mf == 2 ? (j ^ 1) : j)] = w;
      T[4 * OCG_TS + j] = massless ? 1.f : (e2 * w) * w;
      if (narr == 6) T[5 * OCG_TS + j] = massless ? 0.f : (float)sqrt((double)(S.w / __int_as_float(misc[MISC_M0])));  // 1/w
    } else {
      T[j] = S.x * sc;
      T[OCG_TS + j] = S.y * sc;
      T[2 * OCG_TS + j] = S.z * sc;
      T[3 * OCG_TS + j] = S.w;
      T[4 * OCG_TS + j] = e2;
    }
  } else {
    long long pos = (block_start - foff) + (threadIdx.x - rank_fast);
    near_xyzm[pos] = S;
    near_soft[pos] = h;
  }
}

// Pad the tail of the last fast tile with zero-mass sources (contribute exactly 0; in the mass-folded
// layout the same record reads w = 0, w*x = 0, w^2*e2 = 1, 1/w := 0: d' = 0, y3 = 1, d'*y3 = 0, (y3*r2)*0 = 0).
__global__ void pad_tiles_kernel(float* tiles, const int* misc, int mf, int narr) {
  int nf = misc[MISC_NFAST];
  int nt = misc[MISC_NFAST_TILES];
  long long end = (long long)nt * OCG_TS;
  for (long long pos = nf + threadIdx.x; pos < end; pos += blockDim.x) {
    long long tile = pos / OCG_TS;
    int j = (int)(pos - tile * OCG_TS);
    float* T = tiles + tile * (long long)narr * OCG_TS;
    T[j] = 0.f, T[OCG_TS + j] = 0.f, T[2 * OCG_TS + j] = 0.f;
    T[3 * OCG_TS + (mf == 2 ? (j ^ 1) : j)] = 0.f;  // pair-swizzled w array: the pad owns slot j^1, not j
    T[4 * OCG_TS + j] = 1.f;
    if (narr == 6) T[5 * OCG_TS + j] = 0.f;
  }
}

// ------------------------------------------------------------------- finish: sum partials ----
// out[c][t] (+)= G * s^2 * sum_{slot} partial[slot][c][t]  (potential: G * s), slots in index order.
// (partials are in scaled units: acc' = acc / s^2, phi' = phi / s.)
__global__ void finish_kernel(const double* __restrict__ partial, long long stride, int n_slots,
                              int nc_partial, double G, const int* __restrict__ misc, long long n_tgt,
                              double* __restrict__ acc, double* __restrict__ pot, int accumulate, int mf) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_tgt) return;
  const double sc = (double)__int_as_float(misc[MISC_SCALE]);
  if (mf) G *= (double)__int_as_float(misc[MISC_M0]);  // mass-folded partials are in units of M0
  for (int c = 0; c < nc_partial; ++c) {
    double s = 0.0;
    for (int k = 0; k < n_slots; ++k) s += partial[((long long)k * nc_partial + c) * stride + t];
    s *= c < 3 ? G * sc * sc : G * sc;
    double* dst = c < 3 ? acc + (long long)c * n_tgt + t : pot + t;
    *dst = accumulate ? *dst + s : s;
  }
}

// ---------------------------------------------------------------------- near (FP64) kernel ----
// Pair-by-pair FP64 evaluation of the NEAR set; adds G * sum straight into acc/pot.
// Spline forms: pykdgrav ForceKernel / PotentialKernel (cubic spline of support h; Springel 2001).
__device__ __forceinline__ void near_pair(double dx, double dy, double dz, double m, double h, int kernel,
                                          double& fac, double& pfac) {
  const double r2 = dx * dx + dy * dy + dz * dz;
  fac = 0.0, pfac = 0.0;
  if (kernel == OCG_KERNEL_PLUMMER) {
    const double q2 = r2 + h * h;
    if (q2 > 0.0) {
      const double ri = rsqrt(q2);
      pfac = -m * ri;
      fac = m * ri * ri * ri;
    }
    return;
  }
  if (!(r2 > 0.0)) return;
  const double r = sqrt(r2);
  if (r >= h) {
    const double ri = 1.0 / r;
    pfac = -m * ri;
    fac = m * ri * ri * ri;
    return;
  }
  const double hinv = 1.0 / h, q = r * hinv, h3 = hinv * hinv * hinv;
  if (q <= 0.5) {
    fac = m * h3 * (32.0 / 3.0 + q * q * (32.0 * q - 38.4));
    pfac = m * hinv * (-2.8 + q * q * (16.0 / 3.0 + q * q * (6.4 * q - 9.6)));
  } else {
    const double q3 = q * q * q;
    fac = m * h3 * (64.0 / 3.0 - 48.0 * q + 38.4 * q * q - (32.0 / 3.0) * q3 - (1.0 / 15.0) / q3);
    pfac = m * hinv * (-3.2 + (1.0 / 15.0) / q + q * q * (32.0 / 3.0 + q * (-16.0 + q * (9.6 - (32.0 / 15.0) * q))));
  }
}

#define NEAR_BLOCK 128
// grid = (target blocks, source slices).  One slice: the sums are added straight into acc / pot.  Several slices (few targets:
// a 16^3 grid is 33 target blocks on 148 SMs): every slice writes its sums into near_partial[slice][NC][n_tgt] and
// near_finish_kernel adds the slices in slice order — deterministic, unlike atomics.
__global__ void __launch_bounds__(NEAR_BLOCK) near_sum_kernel(
    const float4* __restrict__ near_xyzm, const float* __restrict__ near_soft,
    const int* __restrict__ misc, const float4* __restrict__ tgt, long long n_tgt, int kernel,
    double G, double* __restrict__ acc, double* __restrict__ pot, double* __restrict__ near_partial) {
  const int n_near = misc[MISC_NNEAR];
  if (n_near == 0) return;
  __shared__ float4 sS[NEAR_BLOCK];
  __shared__ float sH[NEAR_BLOCK];
  long long t = blockIdx.x * (long long)NEAR_BLOCK + threadIdx.x;
  const float4 T = tgt[t < n_tgt ? t : n_tgt - 1];
  const double tx = T.x, ty = T.y, tz = T.z;
  double a0 = 0, a1 = 0, a2 = 0, ph = 0;
  // this slice's share of the near list, in whole staging blocks
  const int nblk = (n_near + NEAR_BLOCK - 1) / NEAR_BLOCK;
  const int b0 = (int)((long long)blockIdx.y * nblk / gridDim.y), b1 = (int)((long long)(blockIdx.y + 1) * nblk / gridDim.y);
  for (int base = b0 * NEAR_BLOCK; base < b1 * NEAR_BLOCK; base += NEAR_BLOCK) {
    int i = base + threadIdx.x;
    __syncthreads();
    if (i < n_near) {
      sS[threadIdx.x] = near_xyzm[i];
      sH[threadIdx.x] = near_soft[i];
    }
    __syncthreads();
    const int cnt = min(NEAR_BLOCK, n_near - base);
    for (int j = 0; j < cnt; ++j) {
      const float4 S = sS[j];
      const double dx = (double)S.x - tx, dy = (double)S.y - ty, dz = (double)S.z - tz;
      double fac, pfac;
      near_pair(dx, dy, dz, (double)S.w, (double)sH[j], kernel, fac, pfac);
      a0 += dx * fac, a1 += dy * fac, a2 += dz * fac;
      ph += pfac;
    }
  }
  if (t < n_tgt) {
    if (gridDim.y == 1) {
      acc[t] += G * a0;
      acc[n_tgt + t] += G * a1;
      acc[2 * n_tgt + t] += G * a2;
      if (pot) pot[t] += G * ph;
    } else {
      double* P = near_partial + (long long)blockIdx.y * 4 * n_tgt;
      P[t] = a0, P[n_tgt + t] = a1, P[2 * n_tgt + t] = a2, P[3 * n_tgt + t] = ph;
    }
  }
}
__global__ void near_finish_kernel(const double* __restrict__ near_partial, int n_slices, const int* __restrict__ misc, long long n_tgt,
                                   double G, double* __restrict__ acc, double* __restrict__ pot) {
  if (misc[MISC_NNEAR] == 0) return;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_tgt) return;
  for (int c = 0; c < (pot ? 4 : 3); ++c) {
    double s = 0.0;
    for (int k = 0; k < n_slices; ++k) s += near_partial[((long long)k * 4 + c) * n_tgt + t];
    if (c < 3) acc[(long long)c * n_tgt + t] += G * s;
    else pot[t] += G * s;
  }
}

// ------------------------------------------------------------------------------ launchers ----
typedef void (*direct_fn)(const DirectParams);
static size_t direct_smem_bytes() { return OCG_NSTAGE * OCG_TILE_BYTES + 2 * OCG_NSTAGE * 8; }

// A shape of the streaming kernel. fn[pot][guard]; entries without a form are null.  The table index is the variant id
// quoted in profiles/ (stable across rounds).  Only the PRODUCTION shapes are compiled into the shipped library; the
// sweep shapes and the timing experiments (DBG != 0: wrong results by construction) exist only in the -DOCG_TUNING
// build (oc_nbody_b200/build.py tuning=True -> liboc_nbody_b200_tuning.so), where tools/probe.py selects them through
// ocg_debug_set (include/ocg_debug.h).
struct DirectVariant {
  const char* name;
  int tpt, minb;
  bool ded;
  direct_fn fn[2][2];
  int smem_acc_comps;  // > 0: FP64 accumulators in shared memory, this many doubles per thread and component
  int nwarps;          // consumer warps per CTA (0 = OCG_CONSUMER_WARPS)
  int mf;              // consumes mass-folded tiles (K1 without potential only); 2 = w array pair-swizzled
};
static inline int variant_threads(const DirectVariant& v) { return 32 * (v.nwarps ? v.nwarps : OCG_CONSUMER_WARPS); }
// target-paired shapes (direct_sum_tp_kernel: stream-K rows, fused finish) vs source-paired ones (direct_sum_kernel: items)
static inline bool variant_is_tp(const DirectVariant& v) { return !strncmp(v.name, "tpair", 5) || !strncmp(v.name, "DBG", 3); }
#define OCG_NOFN {{nullptr, nullptr}, {nullptr, nullptr}}
#ifdef OCG_TUNING
#define TUNE(...) __VA_ARGS__
#else
#define TUNE(...) OCG_NOFN
#endif
#define OCG_FULL(TPT, PACKED, DED, MINB, UNR)                                                            \
  {                                                                                                      \
    {direct_sum_kernel<TPT, false, false, PACKED, DED, MINB, UNR>,                                       \
     direct_sum_kernel<TPT, false, true, PACKED, DED, MINB, UNR>},                                       \
    {direct_sum_kernel<TPT, true, false, PACKED, DED, MINB, UNR>,                                        \
     direct_sum_kernel<TPT, true, true, PACKED, DED, MINB, UNR>}                                         \
  }
#define OCG_ONE(TPT, PACKED, DED, MINB, UNR)                                                             \
  {                                                                                                      \
    {direct_sum_kernel<TPT, false, false, PACKED, DED, MINB, UNR>, nullptr}, { nullptr, nullptr }        \
  }
#define OCG_ONEP(TPT, DED, MINB, UNR)                                                                    \
  {                                                                                                      \
    {direct_sum_kernel<TPT, false, false, true, DED, MINB, UNR, true>, nullptr}, { nullptr, nullptr }    \
  }
#define OCG_TP(NP, SMEMACC, MINB, UNR)                                                                   \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR>, nullptr},                                      \
    { direct_sum_tp_kernel<NP, true, SMEMACC, MINB, UNR>, nullptr }                                      \
  }
#define OCG_TPW(NP, SMEMACC, MINB, UNR, NW)                                                              \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW>, nullptr},                                  \
    { direct_sum_tp_kernel<NP, true, SMEMACC, MINB, UNR, NW>, nullptr }                                  \
  }
/* plain tiles, with and without potential, FP32 runs of FOLD sources */
#define OCG_TPWF(NP, SMEMACC, MINB, UNR, NW, FOLD)                                                       \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW, 0, 0, FOLD>, nullptr},                      \
    { direct_sum_tp_kernel<NP, true, SMEMACC, MINB, UNR, NW, 0, 0, FOLD>, nullptr }                      \
  }
#define OCG_TPMF(NP, SMEMACC, MINB, UNR, NW)                                                             \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW, 0, 1>, nullptr}, { nullptr, nullptr }       \
  }
#define OCG_TPMFX(NP, SMEMACC, MINB, UNR, NW, DBG, MF)                                                   \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW, DBG, MF>, nullptr}, { nullptr, nullptr }    \
  }
/* mass-folded, pair-swizzled tiles, FP32 runs of FOLD sources */
#define OCG_TPMFF(NP, MINB, UNR, NW, FOLD)                                                               \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, true, MINB, UNR, NW, 0, 2, FOLD>, nullptr}, { nullptr, nullptr }    \
  }
/* mass-folded, pair-swizzled tiles, with and without potential (6- and 5-array tiles), SMEMACC selectable */
#define OCG_TPMFP(NP, SMEMACC, MINB, UNR, NW, FOLD)                                                      \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW, 0, 2, FOLD>, nullptr},                      \
    { direct_sum_tp_kernel<NP, true, SMEMACC, MINB, UNR, NW, 0, 2, FOLD>, nullptr }                      \
  }
/* the same with staggered folds (STAG: which half of the warps folds half a run later) */
#define OCG_TPMFS(NP, MINB, UNR, NW, FOLD, STAG)                                                         \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, true, MINB, UNR, NW, 0, 2, FOLD, STAG>, nullptr},                   \
    { direct_sum_tp_kernel<NP, true, true, MINB, UNR, NW, 0, 2, FOLD, STAG>, nullptr }                   \
  }
#define OCG_TPD(NP, SMEMACC, MINB, UNR, NW, DBG)                                                         \
  {                                                                                                      \
    {direct_sum_tp_kernel<NP, false, SMEMACC, MINB, UNR, NW, DBG>, nullptr}, { nullptr, nullptr }        \
  }
static const DirectVariant g_variants[] = {
    /* 0 */ {"tpt2 packed ded minb2 unr2", 2, 2, true, TUNE(OCG_FULL(2, true, true, 2, 2))},
    /* 1 */ {"tpt1 packed ded minb2 unr2", 1, 2, true, OCG_FULL(1, true, true, 2, 2)},  // production: SMALL
    /* 2 */ {"tpt2 scalar ded minb2 unr2", 2, 2, true, TUNE(OCG_FULL(2, false, true, 2, 2))},
    /* 3 */ {"tpt1 scalar ded minb2 unr2", 1, 2, true, TUNE(OCG_FULL(1, false, true, 2, 2))},
    /* 4 */ {"tpt2 packed inl minb2 unr2", 2, 2, false, OCG_FULL(2, true, false, 2, 2)},  // production: MID_GUARD
    /* 5 */ {"tpt1 packed inl minb3 unr2", 1, 3, false, TUNE(OCG_ONE(1, true, false, 3, 2))},
    /* 6 */ {"tpt1 packed inl minb4 unr2", 1, 4, false, TUNE(OCG_ONE(1, true, false, 4, 2))},
    /* 7 */ {"tpt2 packed inl minb3 unr2", 2, 3, false, TUNE(OCG_ONE(2, true, false, 3, 2))},
    /* 8 */ {"tpt2 packed inl minb2 unr4", 2, 2, false, TUNE(OCG_ONE(2, true, false, 2, 4))},
    /* 9 */ {"tpt1 packed inl minb3 unr4", 1, 3, false, TUNE(OCG_ONE(1, true, false, 3, 4))},
    /* 10 */ {"tpt2 packed ded minb2 unr1", 2, 2, true, TUNE(OCG_ONE(2, true, true, 2, 1))},
    /* 11 */ {"tpt2 packed inl minb2 unr1", 2, 2, false, TUNE(OCG_ONE(2, true, false, 2, 1))},
    /* 12 */ {"tpt1 packed inl minb2 unr2", 1, 2, false, TUNE(OCG_ONE(1, true, false, 2, 2))},
    /* 13 */ {"tpt1 packed inl minb3 unr1", 1, 3, false, TUNE(OCG_ONE(1, true, false, 3, 1))},
    /* 14 */ {"tpt2 packed ded minb2 unr4", 2, 2, true, TUNE(OCG_ONE(2, true, true, 2, 4))},
    /* 15 */ {"tpt2 packed inl minb2 unr2 pipe", 2, 2, false, TUNE(OCG_ONEP(2, false, 2, 2))},
    /* 16 */ {"tpt2 packed inl minb2 unr1 pipe", 2, 2, false, TUNE(OCG_ONEP(2, false, 2, 1))},
    /* 17 */ {"tpt2 packed ded minb2 unr2 pipe", 2, 2, true, TUNE(OCG_ONEP(2, true, 2, 2))},
    /* 18 */ {"tpt1 packed inl minb3 unr2 pipe", 1, 3, false, TUNE(OCG_ONEP(1, false, 3, 2))},
    /* 19 */ {"tpt1 packed inl minb2 unr4 pipe", 1, 2, false, TUNE(OCG_ONEP(1, false, 2, 4))},
    /* 20 */ {"tpt2 packed inl minb2 unr4 pipe", 2, 2, false, TUNE(OCG_ONEP(2, false, 2, 4))},
    /* 21 */ {"tpair np2 (4 tgt/thr) regacc minb2 unr1", 4, 2, false, TUNE(OCG_TP(2, false, 2, 1)), 0},
    /* 22 */ {"tpair np2 (4 tgt/thr) regacc minb2 unr2", 4, 2, false, TUNE(OCG_TP(2, false, 2, 2)), 0},
    /* 23 */ {"tpair np2 (4 tgt/thr) smemacc minb2 unr1", 4, 2, false, TUNE(OCG_TP(2, true, 2, 1)), 4},
    /* 24 */ {"tpair np4 (8 tgt/thr) smemacc minb2 unr1", 8, 2, false, TUNE(OCG_TP(4, true, 2, 1)), 8},
    /* 25 */ {"tpair np3 (6 tgt/thr) smemacc minb2 unr1", 6, 2, false, TUNE(OCG_TP(3, true, 2, 1)), 6},
    /* 26 */ {"tpair np4 (8 tgt/thr) smemacc minb1 unr1", 8, 1, false, OCG_TP(4, true, 1, 1), 8},  // production: WIDE (K4)
    /* 27 */ {"tpair np1 (2 tgt/thr) regacc minb2 unr2", 2, 2, false, OCG_TP(1, false, 2, 2), 0},  // production: MID (K4)
    /* 28 */ {"tpair np2 (4 tgt/thr) regacc minb3 unr1", 4, 3, false, TUNE(OCG_TP(2, false, 3, 1)), 0},
    /* 29 */ {"tpair np4 smemacc 4w x minb3 (12 w/SM)", 8, 3, false, TUNE(OCG_TPW(4, true, 3, 1, 4)), 8, 4},
    /* 30 */ {"tpair np3 smemacc 4w x minb3 (12 w/SM)", 6, 3, false, TUNE(OCG_TPW(3, true, 3, 1, 4)), 6, 4},
    /* 31 */ {"tpair np4 smemacc 12w x minb1 (12 w/SM)", 8, 1, false, OCG_TPW(4, true, 1, 1, 12), 8, 12},  // production: BIG
    /* 32 */ {"tpair np4 smemacc 6w x minb2 (12 w/SM)", 8, 2, false, TUNE(OCG_TPW(4, true, 2, 1, 6)), 8, 6},
    /* 33 */ {"tpair np2 regacc 4w x minb3 (12 w/SM)", 4, 3, false, TUNE(OCG_TPW(2, false, 3, 1, 4)), 0, 4},
    /* 34 */ {"tpair np2 regacc 4w x minb3 unr2 (12 w/SM)", 4, 3, false, TUNE(OCG_TPW(2, false, 3, 2, 4)), 0, 4},
    /* 35 */ {"tpair np4 smemacc 4w x minb2 (8 w/SM)", 8, 2, false, TUNE(OCG_TPW(4, true, 2, 1, 4)), 8, 4},
    /* 36 */ {"tpair np4 smemacc 10w x minb1 (10 w/SM)", 8, 1, false, TUNE(OCG_TPW(4, true, 1, 1, 10)), 8, 10},
    /* 37 */ {"DBG np4 12w: no MUFU (timing only)", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 1)), 8, 12},
    /* 38 */ {"DBG np4 12w: no LDS (timing only)", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 2)), 8, 12},
    /* 39 */ {"DBG np4 12w: no MUFU, no LDS (timing only)", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 3)), 8, 12},
    /* 40 */ {"tpair-mf np4 smemacc 12w x minb1", 8, 1, false, TUNE(OCG_TPMF(4, true, 1, 1, 12)), 8, 12, 1},
    /* 41 */ {"tpair-mf np4 smemacc 4w x minb3", 8, 3, false, TUNE(OCG_TPMF(4, true, 3, 1, 4)), 8, 4, 1},
    /* 42 */ {"tpair-mf np2 smemacc 8w x minb2", 4, 2, false, TUNE(OCG_TPMF(2, true, 2, 1, 8)), 4, 8, 1},
    /* 43 */ {"tpair-mf np1 regacc 8w x minb2 unr2", 2, 2, false,
              TUNE({{direct_sum_tp_kernel<1, false, false, 2, 2, OCG_CONSUMER_WARPS, 0, 1>, nullptr}, {nullptr, nullptr}}), 0, 0, 1},
    /* 44 */ {"tpair-mf np3 smemacc 12w x minb1", 6, 1, false, TUNE(OCG_TPMF(3, true, 1, 1, 12)), 6, 12, 1},
    /* 45 */ {"tpair-mf np4 smemacc 12w x minb1 unr2", 8, 1, false, TUNE(OCG_TPMF(4, true, 1, 2, 12)), 8, 12, 1},
    /* 46 */ {"tpair-mf np4 12w, w pair-swizzled", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 0, 2)), 8, 12, 2},
    /* 47 */ {"tpair-mf np2 8w x minb2, w pair-swizzled", 4, 2, false, TUNE(OCG_TPMFX(2, true, 2, 1, 8, 0, 2)), 4, 8, 2},
    /* 48 */ {"DBG mf np4 12w: no MUFU", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 4, 1)), 8, 12, 1},
    /* 49 */ {"DBG mf np4 12w: 2-pair accumulate", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 8, 1)), 8, 12, 1},
    /* 50 */ {"DBG mf np4 12w: FADD2 differences", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 16, 1)), 8, 12, 1},
    /* 51 */ {"DBG mf np4 12w: no MUFU + 2-pair accumulate", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 12, 1)), 8, 12, 1},
    /* 52 */ {"DBG mf np4 12w: no MUFU + 2-pair + FADD2", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 28, 1)), 8, 12, 1},
    /* 53 */ {"DBG mf np4 12w swizzled: no MUFU", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 4, 2)), 8, 12, 2},
    /* 54 */ {"DBG mf np4 12w swizzled: no MUFU + 2-pair", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 12, 12, 2)), 8, 12, 2},
    /* 55 */ {"DBG plain np4 12w: no MUFU", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 4)), 8, 12},
    /* 56 */ {"DBG plain np4 12w: 2-pair accumulate", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 8)), 8, 12},
    /* 57 */ {"DBG plain np4 12w: no MUFU + 2-pair", 8, 1, false, TUNE(OCG_TPD(4, true, 1, 1, 12, 12)), 8, 12},
    /* 58 */ {"tpair-mf np6 8w x minb1, swizzled", 12, 1, false, TUNE(OCG_TPMFX(6, true, 1, 1, 8, 0, 2)), 12, 8, 2},
    /* 59 */ {"tpair-mf np8 8w x minb1, swizzled", 16, 1, false, TUNE(OCG_TPMFX(8, true, 1, 1, 8, 0, 2)), 16, 8, 2},
    /* 60 */ {"tpair-mf np5 12w x minb1, swizzled", 10, 1, false, TUNE(OCG_TPMFX(5, true, 1, 1, 12, 0, 2)), 10, 12, 2},
    /* 61 */ {"tpair-mf np6 12w x minb1, swizzled", 12, 1, false, TUNE(OCG_TPMFX(6, true, 1, 1, 12, 0, 2)), 12, 12, 2},
    /* 62 */ {"tpair-mf np4 12w x minb1 unr2, swizzled", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 2, 12, 0, 2)), 8, 12, 2},
    /* 63 */ {"tpair-mf np4 4w x minb3, swizzled", 8, 3, false, TUNE(OCG_TPMFX(4, true, 3, 1, 4, 0, 2)), 8, 4, 2},
    /* 64 */ {"tpair-mf np3 12w x minb1, swizzled", 6, 1, false, TUNE(OCG_TPMFX(3, true, 1, 1, 12, 0, 2)), 6, 12, 2},
    /* 65 */ {"tpair-mf np4 16w x minb1, swizzled", 8, 1, false, TUNE(OCG_TPMFX(4, true, 1, 1, 16, 0, 2)), 8, 16, 2},
    // ---- round 2: FP32 accumulation runs shorter than the tile (FOLD) ----
    /* 66 */ {"tpair-mf np6 8w swizzled fold128", 12, 1, false, TUNE(OCG_TPMFF(6, 1, 1, 8, 128)), 12, 8, 2},
    /* 67 */ {"tpair-mf np6 8w swizzled fold64", 12, 1, false, TUNE(OCG_TPMFF(6, 1, 1, 8, 64)), 12, 8, 2},
    /* 68 */ {"tpair-mf np6 8w swizzled fold32", 12, 1, false, TUNE(OCG_TPMFF(6, 1, 1, 8, 32)), 12, 8, 2},
    /* 69 */ {"tpair np4 12w fold128", 8, 1, false, TUNE(OCG_TPWF(4, true, 1, 1, 12, 128)), 8, 12},
    /* 70 */ {"tpair np4 12w fold64", 8, 1, false, TUNE(OCG_TPWF(4, true, 1, 1, 12, 64)), 8, 12},
    /* 71 */ {"tpair np4 12w fold32", 8, 1, false, TUNE(OCG_TPWF(4, true, 1, 1, 12, 32)), 8, 12},
    /* 72 */ {"tpair np1 regacc minb2 unr2 fold64", 2, 2, false, TUNE(OCG_TPWF(1, false, 2, 2, OCG_CONSUMER_WARPS, 64)), 0},
    /* 73 */ {"tpair np1 regacc minb2 unr2 fold128", 2, 2, false, TUNE(OCG_TPWF(1, false, 2, 2, OCG_CONSUMER_WARPS, 128)), 0},
    // ---- mass-folded tiles with a potential form (6-array tiles), and for the mid-size target counts ----
    /* 74 */ {"tpair-mf(+pot) np4 12w swizzled fold64", 8, 1, false, TUNE(OCG_TPMFP(4, true, 1, 1, 12, 64)), 8, 12, 2},
    /* 75 */ {"tpair-mf(+pot) np5 8w swizzled fold64", 10, 1, false, TUNE(OCG_TPMFP(5, true, 1, 1, 8, 64)), 10, 8, 2},
    /* 76 */ {"tpair-mf(+pot) np4 12w swizzled fold512", 8, 1, false, TUNE(OCG_TPMFP(4, true, 1, 1, 12, 512)), 8, 12, 2},
    /* 77 */ {"tpair-mf(+pot) np1 regacc minb2 unr2 swizzled fold32", 2, 2, false, TUNE(OCG_TPMFP(1, false, 2, 2, OCG_CONSUMER_WARPS, 32)), 0, 0, 2},
    /* 78 */ {"tpair-mf(+pot) np1 regacc minb2 unr2 swizzled fold512", 2, 2, false, TUNE(OCG_TPMFP(1, false, 2, 2, OCG_CONSUMER_WARPS, 512)), 0, 0, 2},
    /* 79 */ {"tpair-mf(+pot) np1 regacc minb2 unr2 swizzled fold64", 2, 2, false, TUNE(OCG_TPMFP(1, false, 2, 2, OCG_CONSUMER_WARPS, 64)), 0, 0, 2},
    /* 80 */ {"tpair-mf(+pot) np1 regacc minb2 unr2 swizzled fold16", 2, 2, false, OCG_TPMFP(1, false, 2, 2, OCG_CONSUMER_WARPS, 16), 0, 0, 2},  // production: MID_MF
    // ---- staggered folds: the two warps of a scheduler fold half a run apart ----
    /* 81 */ {"tpair-mf(+pot) np6 8w fold64, warps >= 4 fold half a run later", 12, 1, false, OCG_TPMFS(6, 1, 1, 8, 64, 1), 12, 8, 2},  // production: BIG_MF
    /* 82 */ {"tpair-mf np6 8w fold64, odd warps fold half a run later", 12, 1, false, TUNE(OCG_TPMFS(6, 1, 1, 8, 64, 2)), 12, 8, 2},
    /* 83 */ {"tpair-mf(+pot) np4 12w fold64, odd warps fold half a run later", 8, 1, false, TUNE(OCG_TPMFS(4, 1, 1, 12, 64, 2)), 8, 12, 2},
    /* 84 */ {"tpair-mf np6 8w fold32, warps >= 4 fold half a run later", 12, 1, false, TUNE(OCG_TPMFS(6, 1, 1, 8, 32, 1)), 12, 8, 2},
};
static const int g_n_variants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));
// Production choices (tools/probe.py sweeps on B200, profiles/r01_variant_sweep*.json, profiles/r02_fold_sweep.json):
// FOLD (profiles/r02_fold_sweep.json, 2e6 particles x 3001 lattice targets, strict metric on the tidal residual):
//   mass-folded  FOLD 512: 1.6e-4 at 76.2 % of FP32 peak | 128: 1.9e-5 at 74.3 % | 64: 1.1e-5 at 73.1 % | 32: 1.2e-5 at 71.1 %
//   plain tiles  FOLD 512: 2.0e-4 at 64.9 %              | 128: 1.4e-4 at 60.6 % | 64: 1.4e-4 at 59.9 % (no gain: their error is
//   the rounding of d = x_s - x_t, coherent over all sources of one binade; the mass-folded d' = fma(-x_t, w, w x_s) dithers it)
#define OCG_VARIANT_BIG 31        /* >= 64k targets, plain tiles: 8 targets/thread, 12 warps (mass folding switched off) */
#define OCG_VARIANT_BIG_MF 81     /* K1 >= 64k targets: mass-folded, w-swizzled tiles, 12 targets/thread, 8 warps, FOLD 64, the second
                                     warp of every scheduler folding half a run after the first (tools/probe.py, 262 145 targets x
                                     1e6 sources: 77.4 % against 76.6 % with all warps folding together; with the potential 61.5 %
                                     against 60.7 % for the 8-target 12-warp shape it replaces) */
#define OCG_VARIANT_BIG_MF_POT 81 /* the same kernel's potential form (6-array tiles) */
#define OCG_VARIANT_MID 27        /* >= 16k targets, plain tiles (K4): target-paired, 2 targets/thread              */
#define OCG_VARIANT_WIDE 26       /* K4 with >= 6 tiles per CTA: plain tiles, 8 targets/thread, 8 warps, 1 CTA/SM (tools/probe_k4.py:
                                     70.2 vs 66.8 % at N = 65 536, 68.1 vs 62.4 % at 256 x 4 096, 70.4 vs 65.9 % at 16 x 16 384; but
                                     41.2 vs 46.2 % at N = 16 384 where a CTA gets fewer than two 2048-target tiles) */
#define OCG_VARIANT_MID_MF 80     /* K1 mid-size target counts: mass-folded, 2 targets/thread, register FP64 accumulators (a fold costs
                                     12 instructions): FOLD 16, with or without potential.  configs[0] (16^3 grid, 1e6 particles,
                                     profiles/r02_accuracy_c0.json), tidal residual strict: FOLD 64 1.2e-5, 32 6.7e-6, 16 3.9e-6 */
#define OCG_VARIANT_MID_GUARD 4   /* source-paired 2 targets/thread (carries the eps2 == 0 guarded form)          */
#define OCG_VARIANT_SMALL 1       /* few targets: 1 target/thread spreads them over more CTAs (has guard form)    */

int ocg_direct_n_variants() { return g_n_variants; }
const char* ocg_direct_variant_name(int id) { return id >= 0 && id < g_n_variants ? g_variants[id].name : ""; }
bool ocg_direct_variant_built(int id) { return id >= 0 && id < g_n_variants && g_variants[id].fn[0][0] != nullptr; }

// n_tgt: targets in the call (or shard); seg_len: typical length of one independent target run (= n_tgt for the
// field build, the cluster size for batched self-gravity) — a tile never spans two runs.
int ocg_pick_variant(ocg_ctx* ctx, int64_t n_tgt, int64_t seg_len, bool guard, bool allow_mf, int64_t src_tiles,
                     bool fine_tiles) {
  const int forced = ctx->knobs.direct_variant;
  if (forced >= 0 && forced < g_n_variants && g_variants[forced].fn[0][0]) {
    // a forced variant without the guarded form falls back to the guarded production kernels; so does a
    // mass-folded one when the caller lays out plain tiles (K4)
    if ((!guard || g_variants[forced].fn[0][1]) && (allow_mf || !g_variants[forced].mf)) return forced;
  }
  auto waste_ok = [&](int v) {
    const long long ct = (long long)variant_threads(g_variants[v]) * g_variants[v].tpt;
    const long long padded = (seg_len + ct - 1) / ct * ct;
    return padded * 8 <= seg_len * 9;  // <= 12.5% of the tile slots idle
  };
  // few targets are fine for the wide-tile kernels when there are enough source tiles to cut into work items
  // (>= 4 tile-granular items per resident CTA); this also keeps the result of a target shard bit-identical to
  // the unsharded call (all target-paired kernels accumulate a target's sources in the same order)
  auto enough = [&](int v) {
    const long long ct = (long long)variant_threads(g_variants[v]) * g_variants[v].tpt;
    return (n_tgt + ct - 1) / ct * src_tiles >= 4ll * ctx->sm_count * g_variants[v].minb;
  };
  if (guard) return (n_tgt >= 16384 && waste_ok(OCG_VARIANT_MID_GUARD)) ? OCG_VARIANT_MID_GUARD : OCG_VARIANT_SMALL;
  // K4 (fine_tiles): clusters of >= 2048 stars take a target-paired stream-K shape.  All of them accumulate a target's
  // sources in the same order (FP32 runs of one tile, FP64 across tiles), so a target shard (ocg_self_gravity_sharded, or
  // tgt_begin/tgt_end) may run another shape than the unsharded call and still reproduce it
  if (fine_tiles) {
    if (!(seg_len >= 2048 || waste_ok(OCG_VARIANT_MID))) return OCG_VARIANT_SMALL;
    // wide (2048-target) or mid (512-target) rows: inner-loop rate x row fill x stream-K balance (units per CTA against the
    // next integer), the wide shape only with >= 6 tiles per CTA to amortise its longer prologue and epilogue
    auto model = [&](int v, double rate, double* per_cta) {
      const long long ct = (long long)variant_threads(g_variants[v]) * g_variants[v].tpt;
      const long long seg_tgt = seg_len < n_tgt ? seg_len : n_tgt;  // targets of one cluster in this call (a shard cuts one cluster)
      const long long n_segs = (n_tgt + seg_tgt - 1) / seg_tgt, rows = n_segs * ((seg_tgt + ct - 1) / ct);
      const double units = (double)rows * (double)src_tiles, g = (double)ctx->sm_count * g_variants[v].minb;
      const double share = units / g, ceil_share = share > (long long)share ? (long long)share + 1.0 : share;
      *per_cta = share;
      return rate * ((double)n_tgt / (double)(rows * ct)) * (share / (ceil_share > 0 ? ceil_share : 1.0));
    };
    double wide_share, mid_share;
    const double e_wide = model(OCG_VARIANT_WIDE, 0.70, &wide_share), e_mid = model(OCG_VARIANT_MID, 0.668, &mid_share);
    return (g_variants[OCG_VARIANT_WIDE].fn[0][0] && wide_share >= 6.0 && e_wide > e_mid) ? OCG_VARIANT_WIDE : OCG_VARIANT_MID;
  }
  // fine_tiles (K4): a cluster has few target tiles, so the 512-target kernel balances better over the SMs than the
  // 3072-target one although its inner loop is ~1 point slower (tools/probe_k4.py: 65.0 vs 63.0 % at N = 65 536,
  // 64.8 vs 57.4 % at 16 x 16 384)
  if (!fine_tiles && (n_tgt >= 65536 || enough(OCG_VARIANT_BIG)) && waste_ok(OCG_VARIANT_BIG)) return OCG_VARIANT_BIG;
  if ((n_tgt >= 16384 || enough(OCG_VARIANT_MID)) && waste_ok(OCG_VARIANT_MID)) return OCG_VARIANT_MID;
  return OCG_VARIANT_SMALL;
}
bool ocg_variant_is_tp(int v) { return variant_is_tp(g_variants[v]); }
int ocg_variant_cluster_tp() { return OCG_VARIANT_MID; }
int ocg_variant_tpt(int v) { return g_variants[v].tpt; }
int ocg_variant_threads(int v) { return variant_threads(g_variants[v]); }
int ocg_variant_slots(ocg_ctx* ctx, int v) { return ctx->sm_count * g_variants[v].minb; }

// Launch the fast kernel over a prepared parameter block.
int ocg_launch_direct(ocg_ctx* ctx, DirectParams& p, int variant, bool pot, bool guard, cudaStream_t st) {
  const DirectVariant* v = &g_variants[variant];
  if (!v->fn[pot][guard]) {
    // sweep-only variant asked for a potential form it does not carry: use a production one of equal tile shape
    const DirectVariant* w = &g_variants[v->tpt == 1 ? OCG_VARIANT_SMALL : OCG_VARIANT_MID_GUARD];
    if (w->tpt != v->tpt || variant_threads(*w) != variant_threads(*v) || !w->fn[pot][guard])
      return ocg_fail(ctx, OCG_ERR_INVALID, "kernel variant %d (%s) has no %s%s form", variant, v->name,
                      pot ? "potential " : "", guard ? "guarded" : "");
    v = w;
  }
  direct_fn fn = v->fn[pot][guard];
  size_t smem = direct_smem_bytes();
  const size_t tile_bytes = (size_t)OCG_TILE_ARRAYS(v->mf, pot) * OCG_TS * sizeof(float);
  if (v->mf) smem = OCG_NSTAGE * tile_bytes + 128;  // target-paired kernels: ring | barriers | FP64 accumulators
  if (v->smem_acc_comps) smem = OCG_NSTAGE * tile_bytes + 128 + (size_t)(pot ? 4 : 3) * v->smem_acc_comps * variant_threads(*v) * sizeof(double);
  OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ctx->sm_count * v->minb;
  if (!variant_is_tp(*v) && grid > p.n_items) grid = p.n_items;  // stream-K kernels always fill the machine
  if (grid < 1) grid = 1;
  if (ctx->timing) OCG_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  fn<<<grid, variant_threads(*v) + (v->ded ? 32 : 0), smem, st>>>(p);
  OCG_CHECK_LAUNCH(ctx, "direct_sum_kernel");
  if (ctx->timing) {
    OCG_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = 1;
  }
  return OCG_OK;
}

int ocg_direct_sum_impl(ocg_ctx* ctx, const float* src_xyzm, const float* src_soft, int64_t n_src,
                        const float* tgt_xyzw, int64_t n_tgt, int kernel, double G, double* acc,
                        double* pot, int accumulate, cudaStream_t st) {
  if (n_tgt <= 0) return OCG_OK;
  if (n_src >= (1ll << 31) - 2 * CLS_BLOCK)
    return ocg_fail(ctx, OCG_ERR_INVALID, "n_src %lld too large for one call; stream in chunks with accumulate=1",
                    (long long)n_src);
  const bool want_pot = pot != nullptr;
  const int NC = want_pot ? 4 : 3;
  if (n_src <= 0) {
    if (!accumulate) {
      OCG_CUDA(ctx, cudaMemsetAsync(acc, 0, sizeof(double) * 3 * n_tgt, st));
      if (pot) OCG_CUDA(ctx, cudaMemsetAsync(pot, 0, sizeof(double) * n_tgt, st));
    }
    return OCG_OK;
  }

  int variant = ocg_pick_variant(ctx, n_tgt, n_tgt, /*guard=*/false, /*allow_mf=*/true, (n_src + OCG_TS - 1) / OCG_TS);
  if (ctx->knobs.mass_fold && ctx->knobs.direct_variant < 0) {
    // the field build takes mass-folded tiles wherever a target-paired kernel runs: one FMA-pipe operation fewer per
    // interaction, and the w factor dithers the rounding of the coordinate difference (see the FOLD table above)
    if (variant == OCG_VARIANT_BIG) variant = want_pot ? OCG_VARIANT_BIG_MF_POT : OCG_VARIANT_BIG_MF;
    else if (variant == OCG_VARIANT_MID) variant = OCG_VARIANT_MID_MF;
  }
  if (g_variants[variant].mf && want_pot && !g_variants[variant].fn[1][0]) variant = OCG_VARIANT_BIG;  // no potential form
  const int mf = g_variants[variant].mf;
  const int narr = OCG_TILE_ARRAYS(mf, want_pot);
  const int tpt = ocg_variant_tpt(variant);
  const int CT = ocg_variant_threads(variant) * tpt;
  const long long n_ttiles = (n_tgt + CT - 1) / CT;
  const long long n_tiles_max = (n_src + OCG_TS - 1) / OCG_TS;
  const long long slots = ocg_variant_slots(ctx, variant);
  const bool tp = variant_is_tp(g_variants[variant]);
  long long n_chunks, tiles_per_chunk = 0;
  // source tiles per pass of the stream-K kernel: what stays in L2 while every CTA streams it (streamk.cuh, PASSES)
  long long pass_tiles = ctx->knobs.pass_bytes > 0 ? (ctx->knobs.pass_bytes + narr * OCG_TS * 4 - 1) / (narr * OCG_TS * 4) : 0;
  // at most 64 passes: every pass adds a set of partial slots (scratch = passes x slots x field size), and beyond
  // 64 x 32 MB of tiles (1e8 sources) what a longer pass re-reads from HBM is a negligible share of the call anyway
  if (pass_tiles > 0 && (n_tiles_max + pass_tiles - 1) / pass_tiles > 64) pass_tiles = (n_tiles_max + 63) / 64;
  const long long n_pass_max = pass_tiles > 0 && pass_tiles < n_tiles_max ? (n_tiles_max + pass_tiles - 1) / pass_tiles : 1;
  if (tp) {
    // stream-K (streamk.cuh): a row of targets is shared by at most ceil(CTAs / rows) + 1 consecutive CTAs per pass
    n_chunks = (slots + n_ttiles - 1) / n_ttiles + 1;
    if (n_chunks > slots) n_chunks = slots;
    n_chunks *= n_pass_max;
    if (n_ttiles > 0x7fffffffll) return ocg_fail(ctx, OCG_ERR_INVALID, "too many target tiles");
  } else {
    // (target tile x source chunk) items dealt round robin: aim for >= 64 equal-cost items per resident CTA slot
    n_chunks = (64 * slots + n_ttiles - 1) / n_ttiles;
    if (n_chunks > n_tiles_max) n_chunks = n_tiles_max;
    if (n_chunks > 1024) n_chunks = 1024;
    if (n_chunks < 1) n_chunks = 1;
    tiles_per_chunk = (n_tiles_max + n_chunks - 1) / n_chunks;
    n_chunks = (n_tiles_max + tiles_per_chunk - 1) / tiles_per_chunk;
    if (n_ttiles * n_chunks > 0x7fffffffll) return ocg_fail(ctx, OCG_ERR_INVALID, "too many work items");
  }

  int* misc;
  float* tiles;
  double* partial;
  float4* near_xyzm;
  float* near_soft;
  int* counts;
  const long long n_cls_blocks = (n_src + CLS_BLOCK - 1) / CLS_BLOCK;
  int rc;
  if ((rc = ocg_scratch(ctx, OCG_SCR_MISC, MISC_INTS * sizeof(int), (void**)&misc))) return rc;
  if ((rc = ocg_scratch(ctx, OCG_SCR_TILES, (size_t)n_tiles_max * narr * OCG_TS * sizeof(float), (void**)&tiles))) return rc;
  if ((rc = ocg_scratch(ctx, OCG_SCR_PARTIAL, (size_t)n_chunks * NC * n_tgt * sizeof(double), (void**)&partial))) return rc;
  if ((rc = ocg_scratch(ctx, OCG_SCR_NEAR, (size_t)n_src * 20, (void**)&near_xyzm))) return rc;
  near_soft = reinterpret_cast<float*>(near_xyzm + n_src);
  if ((rc = ocg_scratch(ctx, OCG_SCR_COUNTS, (size_t)n_cls_blocks * sizeof(int), (void**)&counts))) return rc;
  unsigned int* tickets = nullptr;
  if (tp && (rc = ocg_scratch(ctx, OCG_SCR_TICKETS, (size_t)n_ttiles * sizeof(unsigned int), (void**)&tickets, /*zero_on_alloc=*/true)))
    return rc;

  const float4* src4 = reinterpret_cast<const float4*>(src_xyzm);
  const float4* tgt4 = reinterpret_cast<const float4*>(tgt_xyzw);

  misc_init_kernel<<<1, 32, 0, st>>>(misc);
  OCG_CHECK_LAUNCH(ctx, "misc_init_kernel");
  {
    long long nb = (n_tgt + 255) / 256;
    if (nb > ctx->sm_count * 8) nb = ctx->sm_count * 8;
    bbox_kernel<<<(int)nb, 256, 0, st>>>(tgt4, n_tgt, misc);
    OCG_CHECK_LAUNCH(ctx, "bbox_kernel");
  }
  scale_kernel<<<1, 1, 0, st>>>(misc, 1.0f);
  OCG_CHECK_LAUNCH(ctx, "scale_kernel");
  if (ctx->knobs.precise_near || mf) {
    classify_hist_kernel<<<(int)n_cls_blocks, CLS_BLOCK, 0, st>>>(src4, src_soft, n_src, kernel, misc);
    OCG_CHECK_LAUNCH(ctx, "classify_hist_kernel");
    // cap the FP64 set at ~0.2% of the sources: its pair cost is ~4x the FP32 one, so the near pass stays ~1% of the call.
    // Small snapshots (few, heavy particles: the reference's test_options scale) get a floor of 2^36 / n_src: their single
    // pair terms are a larger share of the field (configs[0], 1e6 particles: tidal residual 1.1e-5 strict with 16384 near
    // sources, 3.9e-6 with 65536, profiles/r02_accuracy_c0.json) and the whole call is milliseconds anyway.
    // A source shard of a build over P ranks (ocg_set_source_shards) takes 1/P of the limit of the WHOLE build, so that
    // the union of the ranks' sets is the unsharded call's (strided shards have the same distance histogram / P) and the
    // FP64 pass does not grow with P (before: 2^36 / n_shard made it 17 % of a rank's step at P = 8 on configs[1]).
    const long long P = ctx->source_shards > 1 ? ctx->source_shards : 1, n_all = n_src * P;
    long long cap = ctx->knobs.near_cap > 0 ? ctx->knobs.near_cap : n_all / 512;
    const long long floor_cap = (1ll << 36) / (n_all > 0 ? n_all : 1);  // 68 719 at 1e6 particles, 6 871 at 1e7, everything below 2.6e5
    if (ctx->knobs.near_cap <= 0 && cap < floor_cap) cap = floor_cap;
    if (ctx->knobs.near_cap <= 0) cap = (cap + P - 1) / P;
    choose_radius_kernel<<<1, 1, 0, st>>>(misc, (int)cap, ctx->knobs.precise_near);
    OCG_CHECK_LAUNCH(ctx, "choose_radius_kernel");
  }
  classify_count_kernel<<<(int)n_cls_blocks, CLS_BLOCK, 0, st>>>(src4, src_soft, n_src, kernel, misc, counts, mf);
  OCG_CHECK_LAUNCH(ctx, "classify_count_kernel");
  classify_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)n_cls_blocks, n_src, misc);
  OCG_CHECK_LAUNCH(ctx, "classify_scan_kernel");
  classify_scatter_kernel<<<(int)n_cls_blocks, CLS_BLOCK, 0, st>>>(src4, src_soft, n_src, kernel, misc, counts,
                                                                 tiles, near_xyzm, near_soft, mf, narr);
  OCG_CHECK_LAUNCH(ctx, "classify_scatter_kernel");
  pad_tiles_kernel<<<1, OCG_TS, 0, st>>>(tiles, misc, mf, narr);
  OCG_CHECK_LAUNCH(ctx, "pad_tiles_kernel");

  DirectParams p;
  p.tiles = tiles;
  p.tgt = tgt4;
  p.partial = partial;
  p.out_stride = n_tgt;
  p.items = nullptr;
  p.n_items = (int)(n_ttiles * n_chunks);
  p.n_tgt = n_tgt;
  p.n_ttiles = (int)n_ttiles;
  p.tiles_per_chunk = (int)tiles_per_chunk;
  p.n_fast_tiles = misc + MISC_NFAST_TILES;
  p.scale_ptr = reinterpret_cast<const float*>(misc + MISC_SCALE);
  p.scale_val = 1.0f;
  p.sk.rows = nullptr, p.sk.row_prefix = nullptr;
  p.sk.n_rows = (int)n_ttiles, p.sk.n_slots = (int)n_chunks;
  p.sk.n_tgt = n_tgt, p.sk.ct = CT;
  p.sk.nst_uniform = misc + MISC_NFAST_TILES;
  p.sk.tickets = tickets;
  p.sk.tile_cap = n_pass_max > 1 ? (int)pass_tiles : 0;
  p.out_acc = acc, p.out_pot = pot, p.out_n = n_tgt;
  p.G = G, p.accumulate = accumulate, p.self_e2s = -1.f;
  p.m0_ptr = mf ? reinterpret_cast<const float*>(misc + MISC_M0) : nullptr;
  if ((rc = ocg_launch_direct(ctx, p, variant, want_pot, /*guard=*/false, st))) return rc;
  // what the launch has to move through HBM: tiles and targets once, the FP64 field once, and for the rows shared by
  // several CTAs one partial slot per extra sharer (at most one extra sharer per CTA boundary)
  ctx->last_traffic_bytes = n_tiles_max * (long long)narr * OCG_TS * 4 + n_tgt * 16 + (long long)NC * n_tgt * 8 +
                            (tp ? 2 * n_pass_max * ((n_pass_max > 1 ? n_ttiles : 0) + (slots < n_ttiles ? slots : n_ttiles)) : n_chunks * n_tgt / CT) *
                                (long long)NC * CT * 8;

  {
    long long nb = (n_tgt + 255) / 256;
    if (!tp) {
      finish_kernel<<<(int)nb, 256, 0, st>>>(partial, n_tgt, (int)n_chunks, NC, G, misc, n_tgt, acc, pot, accumulate, mf);
      OCG_CHECK_LAUNCH(ctx, "finish_kernel");
    }
    const long long nbn = (n_tgt + NEAR_BLOCK - 1) / NEAR_BLOCK;
    // few targets: cut the near list into slices so that the FP64 pass fills the machine (>= 4 CTAs per SM in flight)
    long long slices = (4ll * ctx->sm_count + nbn - 1) / nbn;
    if (slices > 64) slices = 64;
    if (slices < 1) slices = 1;
    double* near_partial = nullptr;
    if (slices > 1 && (rc = ocg_scratch(ctx, OCG_SCR_NEARPART, (size_t)slices * 4 * n_tgt * sizeof(double), (void**)&near_partial))) return rc;
    near_sum_kernel<<<dim3((unsigned)nbn, (unsigned)slices), NEAR_BLOCK, 0, st>>>(near_xyzm, near_soft, misc, tgt4, n_tgt, kernel, G, acc, pot,
                                                                                 near_partial);
    OCG_CHECK_LAUNCH(ctx, "near_sum_kernel");
    if (slices > 1) {
      near_finish_kernel<<<(int)nb, 256, 0, st>>>(near_partial, (int)slices, misc, n_tgt, G, acc, pot);
      OCG_CHECK_LAUNCH(ctx, "near_finish_kernel");
    }
  }
  return OCG_OK;
}

// Per-step cluster bookkeeping of the BRIDGE driver loop, kept on the device:
//   * bound subset + its centre of mass     oc_nbody.py:60-61  (particles.bound_subset().center_of_mass())
//   * median-centred ejection cut           oc_code.py:231-246 (clean_ejections)
//   * stable compaction of the particle arrays after an ejection (system.particles.remove_particles)
// All FP64, O(N) per step next to the O(N^2) potential K4 already produced; fixed reduction trees, so results
// are run-to-run deterministic.
#include "ocg_internal.cuh"

#define CO_BLOCK 1024

// Block-wide sum of NV doubles per thread (fixed tree: warp shuffles, then warp 0 over the warp totals).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* [32][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < (int)(blockDim.x >> 5) ? smem[lane * NV + k] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) smem[k] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = smem[k];
  __syncthreads();
}

// One CTA per cluster (segment).  out[seg][0..2] = centre of mass of the bound stars, [3] their mass, [4] their
// number, [5..7] the velocity of the cluster's centre of mass the energies were taken in.
//   bound_i  <=>  0.5 |v_i - v_com|^2 + pot_to_v2 * phi_i < 0        (no bound star at all: every star counts)
__global__ void __launch_bounds__(CO_BLOCK) bound_com_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                                                             const double* __restrict__ mass, const double* __restrict__ pot,
                                                             long long n, const long long* __restrict__ seg, double pot_to_v2,
                                                             double* __restrict__ out, unsigned char* __restrict__ mask) {
  __shared__ double red[32 * 8];
  const long long a = seg ? seg[blockIdx.x] : 0, b = seg ? seg[blockIdx.x + 1] : n;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // m, m*v (3), m*x (3)
  for (long long i = a + threadIdx.x; i < b; i += CO_BLOCK) {
    const double m = mass[i];
    s[0] += m;
#pragma unroll
    for (int c = 0; c < 3; ++c) s[1 + c] += m * vel[c * n + i], s[4 + c] += m * pos[c * n + i];
  }
  block_sum<8>(s, red);
  const double mtot = s[0];
  const double vc[3] = {s[1] / mtot, s[2] / mtot, s[3] / mtot};
  const double call[3] = {s[4] / mtot, s[5] / mtot, s[6] / mtot};
  double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // bound mass, bound m*x (3), bound count
  for (long long i = a + threadIdx.x; i < b; i += CO_BLOCK) {
    const double dvx = vel[i] - vc[0], dvy = vel[n + i] - vc[1], dvz = vel[2 * n + i] - vc[2];
    const double e = 0.5 * (dvx * dvx + dvy * dvy + dvz * dvz) + pot_to_v2 * pot[i];
    const bool bd = e < 0.0;
    if (mask) mask[i] = bd ? 1 : 0;
    if (bd) {
      const double m = mass[i];
      t[0] += m, t[4] += 1.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) t[1 + c] += m * pos[c * n + i];
    }
  }
  block_sum<8>(t, red);
  if (threadIdx.x == 0) {
    double* o = out + 8 * (long long)blockIdx.x;
    if (t[4] > 0.0) {
      o[0] = t[1] / t[0], o[1] = t[2] / t[0], o[2] = t[3] / t[0], o[3] = t[0], o[4] = t[4];
    } else {
      o[0] = call[0], o[1] = call[1], o[2] = call[2], o[3] = mtot, o[4] = (double)(b - a);
    }
    o[5] = vc[0], o[6] = vc[1], o[7] = vc[2];
  }
  if (mask && t[4] == 0.0)
    for (long long i = a + threadIdx.x; i < b; i += CO_BLOCK) mask[i] = 1;
}

extern "C" int ocg_bound_com(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* mass_dev,
                             const double* pot_dev, int64_t n, const int64_t* seg_offsets_dev, int32_t n_seg,
                             double pot_to_v2, double* out_dev, uint8_t* bound_mask_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || n_seg < 1 || (n > 0 && (!pos_dev || !vel_dev || !mass_dev || !pot_dev || !out_dev)))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_bound_com: bad arguments");
  if (n_seg > 1 && !seg_offsets_dev) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_bound_com: n_seg > 1 needs segment offsets");
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  bound_com_kernel<<<n_seg, CO_BLOCK, 0, (cudaStream_t)stream>>>(pos_dev, vel_dev, mass_dev, pot_dev, n,
                                                                 reinterpret_cast<const long long*>(seg_offsets_dev),
                                                                 pot_to_v2, out_dev, bound_mask_dev);
  OCG_CHECK_LAUNCH(ctx, "bound_com_kernel");
  return OCG_OK;
}

// ---- exact order statistics by radix select on the order-preserving integer image of a double ----
__device__ __forceinline__ unsigned long long f64_key(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// grid = (3 axes, 2 ranks): out[axis][r] = the rank-th smallest of pos[axis][0..n), rank = (n-1)/2 and n/2.
// One CTA decides the key bit by bit from the top: 64 counting passes over the axis.
__global__ void __launch_bounds__(CO_BLOCK) median_select_kernel(const double* __restrict__ pos, long long n,
                                                                 double* __restrict__ out) {
  __shared__ long long wcount[32];
  __shared__ long long total;
  const double* x = pos + (long long)blockIdx.x * n;
  long long k = blockIdx.y == 0 ? (n - 1) / 2 : n / 2;
  unsigned long long prefix = 0, decided = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long bm = 1ull << bit;
    long long c = 0;
    for (long long i = threadIdx.x; i < n; i += CO_BLOCK) {
      const unsigned long long key = f64_key(x[i]);
      c += ((key & decided) == prefix) && !(key & bm);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wcount[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < CO_BLOCK / 32; ++w) t += wcount[w];
      total = t;
    }
    __syncthreads();
    const long long zeros = total;
    if (k >= zeros) {
      k -= zeros;
      prefix |= bm;
    }
    decided |= bm;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x * 2 + blockIdx.y] = key_f64(prefix);
}

// keep[i] = |len_scale * (x_i - median)| <= cut ; median per axis = (lo + hi) / 2 as numpy.median forms it
__global__ void eject_mask_kernel(const double* __restrict__ pos, long long n, const double* __restrict__ sel, double len_scale,
                                  double cut, unsigned char* __restrict__ keep, double* __restrict__ median_out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const double mx = (sel[0] + sel[1]) / 2.0, my = (sel[2] + sel[3]) / 2.0, mz = (sel[4] + sel[5]) / 2.0;
  if (i == 0 && median_out) median_out[0] = mx, median_out[1] = my, median_out[2] = mz;
  if (i >= n) return;
  const double dx = (pos[i] - mx) * len_scale, dy = (pos[n + i] - my) * len_scale, dz = (pos[2 * n + i] - mz) * len_scale;
  keep[i] = !(sqrt(dx * dx + dy * dy + dz * dz) > cut);
}

extern "C" int ocg_eject_mask(ocg_ctx* ctx, const double* pos_dev, int64_t n, double len_scale, double cut,
                              uint8_t* keep_mask_dev, double* median_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || (n > 0 && (!pos_dev || !keep_mask_dev))) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_eject_mask: bad arguments");
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  double* sel;
  int rc;
  if ((rc = ocg_scratch(ctx, OCG_SCR_MISC, 256, (void**)&sel))) return rc;
  median_select_kernel<<<dim3(3, 2), CO_BLOCK, 0, (cudaStream_t)stream>>>(pos_dev, n, sel);
  OCG_CHECK_LAUNCH(ctx, "median_select_kernel");
  eject_mask_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos_dev, n, sel, len_scale, cut, keep_mask_dev,
                                                                              median_out_dev);
  OCG_CHECK_LAUNCH(ctx, "eject_mask_kernel");
  return OCG_OK;
}

// Stable compaction of `rows` FP64 rows of length n by a byte mask: out[r][j] = in[r][i_j], i_j the j-th kept index.
// One CTA walks the mask in chunks with a running offset (n is a cluster's star count); n_keep_dev receives the total.
__global__ void __launch_bounds__(CO_BLOCK) compact_rows_kernel(const double* __restrict__ in, int rows, long long n,
                                                                const unsigned char* __restrict__ keep, double* __restrict__ out,
                                                                long long out_stride, long long* __restrict__ n_keep) {
  __shared__ int wtot[32];
  __shared__ long long base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (long long start = 0; start < n; start += CO_BLOCK) {
    const long long i = start + threadIdx.x;
    const bool k = i < n && keep[i];
    const unsigned b = __ballot_sync(0xffffffffu, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wtot[warp] = __popc(b);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < CO_BLOCK / 32; ++w) {
      const int t = wtot[w];
      if (w < warp) before += t;
      all += t;
    }
    if (k) {
      const long long j = base + before + __popc(b & ((1u << lane) - 1u));
      if (out)
        for (int r = 0; r < rows; ++r) out[r * out_stride + j] = in[r * n + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) base += all;
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_keep) *n_keep = base;
}

extern "C" int ocg_compact_rows(ocg_ctx* ctx, const double* in_dev, int32_t rows, int64_t n, const uint8_t* keep_mask_dev,
                                double* out_dev, int64_t out_stride, int64_t* n_keep_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || rows < 0 || (n > 0 && !keep_mask_dev) || (out_dev && !in_dev))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_compact_rows: bad arguments");
  if (n == 0) {
    if (n_keep_dev) cudaMemsetAsync(n_keep_dev, 0, sizeof(int64_t), (cudaStream_t)stream);
    return OCG_OK;
  }
  OcgDeviceGuard g(ctx->device);
  compact_rows_kernel<<<1, CO_BLOCK, 0, (cudaStream_t)stream>>>(in_dev, rows, n, keep_mask_dev, out_dev, out_stride,
                                                                reinterpret_cast<long long*>(n_keep_dev));
  OCG_CHECK_LAUNCH(ctx, "compact_rows_kernel");
  return OCG_OK;
}

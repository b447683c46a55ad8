// Source assembly on the device: what gizmo_interface.py does on the host with numpy before it hands
// the particles to pykdgrav — the Rmax cut of _clean_Rmag_ (gizmo_interface.py:297-304), the exclusion of the tracked star
// (:515), the per-species softening rules (:530-547) and the star | dark | gas concatenation (:518-528, :549) — fused with
// the FP64 recentring + FP32 rounding of SURVEY §7 H3.  The host uploads each species' raw snapshot arrays once; the FP32
// source records K1 consumes are written straight into their place in the concatenated arrays (stable order).
#include "ocg_internal.cuh"

#include <math.h>

#define AS_BLOCK 1024

struct AssembleParams {
  const double* pos;     // [n][3]
  const double* mass;    // [n]
  const long long* id;   // [n] or NULL
  const double* hsml;    // [n] or NULL (gas smoothing length, pc)
  long long n;
  long long exclude_id;
  double rmax;           // <= 0: no cut
  int rule;
  double soft_param, soft_scale;
  double cx, cy, cz;
};

__device__ __forceinline__ bool as_keep(const AssembleParams& p, long long i) {
  if (p.id && p.id[i] == p.exclude_id) return false;
  if (p.rmax > 0.0) {
    const double x = p.pos[3 * i], y = p.pos[3 * i + 1], z = p.pos[3 * i + 2];
    // host.distance.total < Rmax, the norm formed as numpy does: separately rounded squares, summed left to right
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    if (!(r < p.rmax)) return false;
  }
  return true;
}

__global__ void __launch_bounds__(AS_BLOCK) assemble_count_kernel(AssembleParams p, int* __restrict__ counts) {
  __shared__ int wsum[AS_BLOCK / 32];
  const long long i = blockIdx.x * (long long)AS_BLOCK + threadIdx.x;
  const bool k = i < p.n && as_keep(p, i);
  const unsigned b = __ballot_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) counts[blockIdx.x] = v;
  }
}

// single-block exclusive scan of the per-block counts (in place); total -> *total_out
__global__ void __launch_bounds__(1024) assemble_scan_kernel(int* counts, int nblocks, long long* total_out) {
  __shared__ long long carry;
  __shared__ int wtot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int w = wtot[threadIdx.x];
      int wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += u;
      }
      wtot[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const long long c = carry;
    const int excl = incl - v + wtot[threadIdx.x >> 5];
    if (i < nblocks) counts[i] = (int)(c + excl);  // n < 2^31 per call (checked on the host)
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(AS_BLOCK) assemble_scatter_kernel(AssembleParams p, const int* __restrict__ offs,
                                                                    float4* __restrict__ out_xyzm, float* __restrict__ out_soft,
                                                                    long long out_offset) {
  __shared__ int wbase[AS_BLOCK / 32];
  const long long i = blockIdx.x * (long long)AS_BLOCK + threadIdx.x;
  const bool k = i < p.n && as_keep(p, i);
  const unsigned b = __ballot_sync(0xffffffffu, k);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wbase[w] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int v = wbase[threadIdx.x];
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (threadIdx.x >= o) incl += u;
    }
    wbase[threadIdx.x] = incl - v;
  }
  __syncthreads();
  if (!k) return;
  const long long j = out_offset + offs[blockIdx.x] + wbase[w] + __popc(b & ((1u << lane) - 1u));
  const double m = p.mass[i];
  double s;
  if (p.rule == OCG_SOFT_CONSTANT) s = p.soft_param;                              // soft_pc / 1000   (gizmo_interface.py:536-545)
  else if (p.rule == OCG_SOFT_MASS_CUBE_ROOT) s = pow(m / p.soft_param, 1.0 / 3.0) / 1000.0;  // (m / m_char)^(1/3) / 1000 (:531-534)
  else s = 2.8 * p.hsml[i] / 1000.0;                                              // gas: 2.8 smooth.length / 1000 (:547)
  out_xyzm[j] = make_float4((float)(p.pos[3 * i] - p.cx), (float)(p.pos[3 * i + 1] - p.cy), (float)(p.pos[3 * i + 2] - p.cz), (float)m);
  out_soft[j] = (float)(s * p.soft_scale);
}

extern "C" int ocg_assemble_sources(ocg_ctx* ctx, const double* pos_dev, const double* mass_dev, const int64_t* id_dev,
                                    const double* hsml_dev, int64_t n, int64_t exclude_id, double rmax, int32_t soft_rule,
                                    double soft_param, double soft_scale, const double center[3], float* out_xyzm_dev,
                                    float* out_soft_dev, int64_t out_offset, int64_t* n_kept_host, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || out_offset < 0 || !center || !n_kept_host || (n > 0 && (!pos_dev || !mass_dev || !out_xyzm_dev || !out_soft_dev)))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_assemble_sources: bad arguments");
  if (soft_rule != OCG_SOFT_CONSTANT && soft_rule != OCG_SOFT_MASS_CUBE_ROOT && soft_rule != OCG_SOFT_GAS_SMOOTHING)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_assemble_sources: unknown softening rule %d", soft_rule);
  if (soft_rule == OCG_SOFT_GAS_SMOOTHING && n > 0 && !hsml_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_assemble_sources: the gas rule needs the smoothing lengths");
  if (soft_rule == OCG_SOFT_MASS_CUBE_ROOT && !(soft_param > 0.0))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_assemble_sources: characteristic mass %g must be > 0", soft_param);
  if (n >= (1ll << 31) - 2 * AS_BLOCK) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_assemble_sources: %lld particles per call; split the species", (long long)n);
  *n_kept_host = 0;
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long nb = (n + AS_BLOCK - 1) / AS_BLOCK;
  int* counts;
  int rc = ocg_scratch(ctx, OCG_SCR_COUNTS, (size_t)(nb + 2) * sizeof(int) + 16, (void**)&counts);
  if (rc) return rc;
  long long* total = reinterpret_cast<long long*>(counts + ((nb + 1) / 2) * 2);
  AssembleParams p;
  p.pos = pos_dev, p.mass = mass_dev, p.id = reinterpret_cast<const long long*>(id_dev), p.hsml = hsml_dev, p.n = n;
  p.exclude_id = exclude_id, p.rmax = rmax, p.rule = soft_rule, p.soft_param = soft_param, p.soft_scale = soft_scale;
  p.cx = center[0], p.cy = center[1], p.cz = center[2];
  assemble_count_kernel<<<(int)nb, AS_BLOCK, 0, st>>>(p, counts);
  OCG_CHECK_LAUNCH(ctx, "assemble_count_kernel");
  assemble_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)nb, total);
  OCG_CHECK_LAUNCH(ctx, "assemble_scan_kernel");
  assemble_scatter_kernel<<<(int)nb, AS_BLOCK, 0, st>>>(p, counts, reinterpret_cast<float4*>(out_xyzm_dev), out_soft_dev, out_offset);
  OCG_CHECK_LAUNCH(ctx, "assemble_scatter_kernel");
  long long h = 0;
  OCG_CUDA(ctx, cudaMemcpyAsync(&h, total, sizeof(h), cudaMemcpyDeviceToHost, st));
  OCG_CUDA(ctx, cudaStreamSynchronize(st));
  *n_kept_host = h;
  return OCG_OK;
}

// K2 / K3: grid planes and the per-kick evaluation behind get_gravity_at_point
// (gizmo_interface.py:677-717), as trilinear-in-space + linear-in-time interpolation on the
// regular lattice of grid_cartesian.py:16-32,59-69.
//
// Memory-bound gather: one float4 record (ax, ay, az, phi) per node per snapshot, z-adjacent corners
// share a 32-byte sector.  Every operation is individually rounded (no FMA contraction) in a fixed
// order — FP32 for the time blend of the FP32 records, FP64 for cell selection, weights and the
// trilinear blend — so the oracle (oracle/ocg_oracle.c: oracle_grid_interp) reproduces the bits.
#include "ocg_internal.cuh"

__global__ void pack_planes_kernel(const double* __restrict__ acc, const double* __restrict__ pot,
                                   long long n, float4* __restrict__ rec) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  rec[i] = make_float4((float)acc[i], (float)acc[n + i], (float)acc[2 * n + i], pot ? (float)pot[i] : 0.f);
}

__device__ __forceinline__ double lerp_rn(double a, double wa, double b, double wb) {
  return __dadd_rn(__dmul_rn(a, wa), __dmul_rn(b, wb));
}

// v = a*(1-w) + b*w on the FP32 records, every op rounded to FP32 — the same blend K3 fuses.
__global__ void time_blend_kernel(const float4* __restrict__ ra, const float4* __restrict__ rb, float wb,
                                  long long n, double* __restrict__ acc, double* __restrict__ pot) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float wa = __fsub_rn(1.0f, wb);
  float4 a = ra[i];
  if (rb) {
    const float4 b = rb[i];
    a.x = __fadd_rn(__fmul_rn(a.x, wa), __fmul_rn(b.x, wb));
    a.y = __fadd_rn(__fmul_rn(a.y, wa), __fmul_rn(b.y, wb));
    a.z = __fadd_rn(__fmul_rn(a.z, wa), __fmul_rn(b.z, wb));
    a.w = __fadd_rn(__fmul_rn(a.w, wa), __fmul_rn(b.w, wb));
  }
  acc[i] = (double)a.x;
  acc[n + i] = (double)a.y;
  acc[2 * n + i] = (double)a.z;
  if (pot) pot[i] = (double)a.w;
}

static inline int nblocks(long long n, int b) { return (int)((n + b - 1) / b); }

extern "C" int ocg_pack_planes(ocg_ctx* ctx, const double* acc_dev, const double* pot_dev, int64_t n_node,
                               float* rec_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n_node < 0 || (n_node > 0 && (!acc_dev || !rec_dev))) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_pack_planes: bad arguments");
  if (n_node == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  pack_planes_kernel<<<nblocks(n_node, 256), 256, 0, (cudaStream_t)stream>>>(acc_dev, pot_dev, n_node,
                                                                          reinterpret_cast<float4*>(rec_dev));
  OCG_CHECK_LAUNCH(ctx, "pack_planes_kernel");
  return OCG_OK;
}

extern "C" int ocg_grid_time_blend(ocg_ctx* ctx, const float* rec_a_dev, const float* rec_b_dev, double w_b,
                                   int64_t n_node, double* acc_out_dev, double* pot_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n_node < 0 || (n_node > 0 && (!rec_a_dev || !acc_out_dev))) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_time_blend: bad arguments");
  if (n_node == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  time_blend_kernel<<<nblocks(n_node, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(rec_a_dev), reinterpret_cast<const float4*>(rec_b_dev), (float)w_b, n_node,
      acc_out_dev, pot_out_dev);
  OCG_CHECK_LAUNCH(ctx, "time_blend_kernel");
  return OCG_OK;
}

// ------------------------------------------------------------------------------------ K3 ----
struct InterpParams {
  int n[3];
  int n_cluster;
  const double* node[3];
  const double* origin;
  const float4* rec[4];  // 1..4 record planes blended in time (2: linear bracket, 4: cubic B-spline coefficients)
  float w[4];            // their FP32 weights (the time blend is FP32 arithmetic)
  int n_rec;
  const double* sx;
  const double* sy;
  const double* sz;
  const int* scl;
  long long n_star;
  double* acc;
  double* pot;
  int* cell;
};

// Cell along one axis: i = searchsorted(node + o, x, side='right') - 1, clamped to [0, n-2].
// The arithmetic estimate only seeds the search; the result is decided by comparisons against
// the (node[i] + o) values, which the oracle forms with the same single FP64 addition.
__device__ __forceinline__ int find_cell(const double* __restrict__ node, int n, double o, double x) {
  const double lo = __dadd_rn(node[0], o);
  const double hi = __dadd_rn(node[n - 1], o);
  int i = 0;
  if (x == x && n > 2) {  // not NaN
    // seed only (any rounding will do): reciprocal via FP32 keeps the FP64 divide off this path
    double f = (x - lo) * (double)((float)(n - 1) / (float)(hi - lo));
    if (f >= (double)(n - 2)) i = n - 2;
    else if (f > 0.0) i = (int)f;
  }
  while (i > 0 && x < __dadd_rn(node[i], o)) --i;
  while (i < n - 2 && x >= __dadd_rn(node[i + 1], o)) ++i;
  return i;
}

__device__ __forceinline__ float lerp_rn_f(float a, float wa, float b, float wb) {
  return __fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
}

// One thread per star.  Arithmetic contract (identical in oracle/ocg_oracle.c: oracle_grid_interp):
//   cell      : FP64 comparisons against node[i] + origin (bit-exact searchsorted)
//   weight    : t = (x - (node[i] + origin)) * inv[i],  inv[i] = 1 / (node[i+1] - node[i])      FP64
//   time blend: v = a*(1-w) + b*w on the FP32 records, every op rounded to FP32               FP32
//   trilinear : z, then y, then x lerps of the 8 corner values, every op rounded to FP64        FP64
// SMEM_NODES: node and inverse-spacing tables staged in shared memory (else read from global).
template <bool SMEM_NODES, int MINB>
__global__ void __launch_bounds__(256, MINB) grid_interp_kernel(const InterpParams p) {
  extern __shared__ double s_tab[];
  const double* nd[3];
  const double* iv[3];
  if (SMEM_NODES) {
    int off = 0;
    for (int d = 0; d < 3; ++d) {
      double* sn = s_tab + off;
      double* si = sn + p.n[d];
      for (int i = threadIdx.x; i < p.n[d]; i += blockDim.x) {
        sn[i] = p.node[d][i];
        si[i] = i + 1 < p.n[d] ? __ddiv_rn(1.0, __dsub_rn(p.node[d][i + 1], p.node[d][i])) : 0.0;
      }
      nd[d] = sn, iv[d] = si;
      off += 2 * p.n[d];
    }
    __syncthreads();
  } else {
    nd[0] = p.node[0], nd[1] = p.node[1], nd[2] = p.node[2];
    iv[0] = iv[1] = iv[2] = nullptr;
  }
  // grid-stride over stars: the tables above are staged once per resident block, not once per 256 stars
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < p.n_star;
       s += (long long)gridDim.x * blockDim.x) {
  const int cl = p.scl ? p.scl[s] : 0;
  const double* org = p.origin + 3 * (long long)cl;
  const double pos[3] = {p.sx[s], p.sy[s], p.sz[s]};
  int c3[3];
  double t[3], u[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double o = org[d];
    const int i = find_cell(nd[d], p.n[d], o, pos[d]);
    const double inv = SMEM_NODES ? iv[d][i] : __ddiv_rn(1.0, __dsub_rn(nd[d][i + 1], nd[d][i]));
    t[d] = __dmul_rn(__dsub_rn(pos[d], __dadd_rn(nd[d][i], o)), inv);
    u[d] = __dsub_rn(1.0, t[d]);
    c3[d] = i;
  }

  const long long nyz = (long long)p.n[1] * p.n[2];
  const long long n_node = (long long)p.n[0] * nyz + 1;  // + appended origin row
  const long long base = (long long)cl * n_node + ((long long)c3[0] * p.n[1] + c3[1]) * p.n[2] + c3[2];
  // corner order: (di,dj,dk) = 000,001,010,011,100,101,110,111
  // time blend in FP32, left to right: v = ((r0*w0 + r1*w1) + r2*w2) + r3*w3 ; a single plane is taken as is
  float4 v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const long long off = base + (long long)(c >> 2) * nyz + (long long)((c >> 1) & 1) * p.n[2] + (c & 1);
    float4 a = __ldg(p.rec[0] + off);
    if (p.n_rec > 1) {
      const float4 b = __ldg(p.rec[1] + off);
      a = make_float4(lerp_rn_f(a.x, p.w[0], b.x, p.w[1]), lerp_rn_f(a.y, p.w[0], b.y, p.w[1]),
                      lerp_rn_f(a.z, p.w[0], b.z, p.w[1]), lerp_rn_f(a.w, p.w[0], b.w, p.w[1]));
      for (int r = 2; r < p.n_rec; ++r) {
        const float4 e = __ldg(p.rec[r] + off);
        const float wr = p.w[r];
        a = make_float4(__fadd_rn(a.x, __fmul_rn(e.x, wr)), __fadd_rn(a.y, __fmul_rn(e.y, wr)),
                        __fadd_rn(a.z, __fmul_rn(e.z, wr)), __fadd_rn(a.w, __fmul_rn(e.w, wr)));
      }
    }
    v[c] = a;
  }
  const int ncomp = p.pot ? 4 : 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q >= ncomp) break;
    double w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = (double)(q == 0 ? v[c].x : (q == 1 ? v[c].y : (q == 2 ? v[c].z : v[c].w)));
    const double c00 = lerp_rn(w[0], u[2], w[1], t[2]);
    const double c01 = lerp_rn(w[2], u[2], w[3], t[2]);
    const double c10 = lerp_rn(w[4], u[2], w[5], t[2]);
    const double c11 = lerp_rn(w[6], u[2], w[7], t[2]);
    const double c0 = lerp_rn(c00, u[1], c01, t[1]);
    const double c1 = lerp_rn(c10, u[1], c11, t[1]);
    const double r = lerp_rn(c0, u[0], c1, t[0]);
    if (q < 3) p.acc[(long long)q * p.n_star + s] = r;
    else p.pot[s] = r;
  }
  if (p.cell) {
    p.cell[s] = c3[0];
    p.cell[p.n_star + s] = c3[1];
    p.cell[2 * p.n_star + s] = c3[2];
  }
  }
}

static int g_interp_variant = 2;  // 0: <=128 regs (2 blocks/SM), 1: <=80 regs (3), 2: <=64 regs (4: fastest, latency-bound)
extern "C" int ocg_debug_set_interp_variant(int v) {
  if (v < 0 || v > 2) return OCG_ERR_INVALID;
  g_interp_variant = v;
  return OCG_OK;
}

static int interp_launch(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* const* rec, const float* w, int n_rec,
                         const double* sx, const double* sy, const double* sz, const int32_t* scl, int64_t n_star,
                         double* acc, double* pot, int32_t* cell, void* stream, const char* who) {
  if (!grid) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: grid is NULL", who);
  if (n_star < 0) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: negative n_star", who);
  if (n_star == 0) return OCG_OK;
  for (int d = 0; d < 3; ++d)
    if (grid->n[d] < 2 || !grid->node_dev[d])
      return ocg_fail(ctx, OCG_ERR_INVALID, "%s: axis %d needs >= 2 nodes (got %d)", who, d, grid->n[d]);
  if (grid->n_cluster < 1 || !grid->origin_dev || !sx || !sy || !sz || !acc)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: NULL argument", who);
  if (n_rec < 1 || n_rec > 4) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: n_rec = %d outside [1,4]", who, n_rec);
  OcgDeviceGuard g(ctx->device);
  InterpParams p;
  for (int d = 0; d < 3; ++d) p.n[d] = grid->n[d], p.node[d] = grid->node_dev[d];
  p.n_cluster = grid->n_cluster;
  p.origin = grid->origin_dev;
  for (int r = 0; r < 4; ++r) {
    p.rec[r] = r < n_rec ? reinterpret_cast<const float4*>(rec[r]) : nullptr;
    p.w[r] = r < n_rec ? w[r] : 0.f;
    if (r < n_rec && !rec[r]) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: record plane %d is NULL", who, r);
  }
  p.n_rec = n_rec;
  p.sx = sx, p.sy = sy, p.sz = sz;
  p.scl = scl;
  p.n_star = n_star;
  p.acc = acc, p.pot = pot, p.cell = cell;
  long long nn = (long long)grid->n[0] + grid->n[1] + grid->n[2];
  // persistent launch: as many blocks as are resident at once (a multiple of the SM count), grid-stride inside
  const bool in_smem = nn <= 2048;
  const size_t smem = in_smem ? (size_t)nn * 2 * sizeof(double) : 0;
  typedef void (*interp_fn)(const InterpParams);
  static const interp_fn fns[3][2] = {{grid_interp_kernel<false, 2>, grid_interp_kernel<true, 2>},
                                      {grid_interp_kernel<false, 3>, grid_interp_kernel<true, 3>},
                                      {grid_interp_kernel<false, 4>, grid_interp_kernel<true, 4>}};
  interp_fn fn = fns[g_interp_variant][in_smem ? 1 : 0];
  int occ = 0;
  OCG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, smem));
  if (occ < 1) occ = 1;
  int grid_blocks = nblocks(n_star, 256);
  if (grid_blocks > ctx->sm_count * occ) grid_blocks = ctx->sm_count * occ;
  fn<<<grid_blocks, 256, smem, (cudaStream_t)stream>>>(p);
  OCG_CHECK_LAUNCH(ctx, "grid_interp_kernel");
  return OCG_OK;
}

extern "C" int ocg_grid_interp(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* rec_a_dev,
                               const float* rec_b_dev, double w_b, const double* star_x_dev,
                               const double* star_y_dev, const double* star_z_dev,
                               const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                               double* pot_out_dev, int32_t* cell_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!rec_a_dev) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: NULL argument");
  if (!(w_b >= 0.0 && w_b <= 1.0))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: w_b = %g outside [0,1]", w_b);
  const float* rec[2] = {rec_a_dev, rec_b_dev};
  const float wb = (float)w_b;
  volatile float wa = 1.0f - wb;  // rounded to FP32, as the oracle does
  const float w[2] = {wa, wb};
  return interp_launch(ctx, grid, rec, w, rec_b_dev ? 2 : 1, star_x_dev, star_y_dev, star_z_dev, star_cluster_dev,
                       n_star, acc_out_dev, pot_out_dev, cell_out_dev, stream, "ocg_grid_interp");
}

extern "C" int ocg_grid_interp_multi(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* const* rec_dev,
                                     const double* weights, int32_t n_rec, const double* star_x_dev,
                                     const double* star_y_dev, const double* star_z_dev,
                                     const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                                     double* pot_out_dev, int32_t* cell_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!rec_dev || !weights || n_rec < 1 || n_rec > 4)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_multi: need 1..4 record planes and weights");
  float w[4];
  for (int r = 0; r < n_rec; ++r) w[r] = (float)weights[r];
  return interp_launch(ctx, grid, rec_dev, w, n_rec, star_x_dev, star_y_dev, star_z_dev, star_cluster_dev, n_star,
                       acc_out_dev, pot_out_dev, cell_out_dev, stream, "ocg_grid_interp_multi");
}

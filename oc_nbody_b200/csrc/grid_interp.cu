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
// Time-blend weights that a captured CUDA graph can pick up: kernel parameters are frozen at capture, constant memory is
// not.  ocg_set_interp_weight_slots() refreshes the slots (stream-ordered) before each replay; a K3 launch with
// w_slot >= 0 reads its weights from there instead of from its parameter block.
#define OCG_W_SLOTS 4
__constant__ float c_interp_w[OCG_W_SLOTS][4];

struct InterpParams {
  int n[3];
  int n_cluster;
  const double* node[3];
  const double* origin;
  const float4* rec[4];  // 1..4 record planes blended in time (2: linear bracket, 4: cubic B-spline coefficients)
  float w[4];            // their FP32 weights (the time blend is FP32 arithmetic)
  int w_slot;            // >= 0: take the weights from c_interp_w[w_slot] (graph replay), else from w[]
  int n_rec;
  // nested fine lattice (grid_cartesian.py:34-53,71-91) sharing origin and weights; n2[0] == 0: single level
  int n2[3];
  const double* node2[3];
  const float4* rec2[4];
  const double* sx;
  const double* sy;
  const double* sz;
  const int* scl;
  long long n_star;
  double* acc;
  double* pot;
  double* tensor;  // [9][n_star]: tensor[3*i + j] = d a_j / d x_i (gizmo_interface.py:719-756), or NULL
  int* cell;
  int* level;      // [n_star]: 0 coarse, 1 fine, or NULL
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

// One thread per star.  Arithmetic contract (identical in oracle/ocg_oracle.c: grid_interp_core):
//   level     : (nested only) fine iff node2[0] + o <= x <= node2[n2-1] + o on all three axes        FP64 compares
//   cell      : FP64 comparisons against node[i] + origin (bit-exact searchsorted)
//   weight    : t = (x - (node[i] + origin)) * inv[i],  inv[i] = 1 / (node[i+1] - node[i])      FP64
//   time blend: v = a*(1-w) + b*w on the FP32 records, every op rounded to FP32               FP32
//   trilinear : z, then y, then x lerps of the 8 corner values, every op rounded to FP64        FP64
//   tensor    : derivative of that trilinear form: differences of the z / y / x stage values times inv[i]
// SMEM_NODES: node and inverse-spacing tables staged in shared memory (else read from global).
template <bool SMEM_NODES, int MINB, bool NESTED, bool TENSOR>
__global__ void __launch_bounds__(256, MINB) grid_interp_kernel(const InterpParams p) {
  extern __shared__ double s_tab[];
  const double* nd[2][3];
  const double* iv[2][3];
  if (SMEM_NODES) {
    int off = 0;
    for (int l = 0; l < (NESTED ? 2 : 1); ++l)
      for (int d = 0; d < 3; ++d) {
        const int n = l ? p.n2[d] : p.n[d];
        const double* src = l ? p.node2[d] : p.node[d];
        double* sn = s_tab + off;
        double* si = sn + n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          sn[i] = src[i];
          si[i] = i + 1 < n ? __ddiv_rn(1.0, __dsub_rn(src[i + 1], src[i])) : 0.0;
        }
        nd[l][d] = sn, iv[l][d] = si;
        off += 2 * n;
      }
    __syncthreads();
  } else {
    for (int d = 0; d < 3; ++d) {
      nd[0][d] = p.node[d], nd[1][d] = NESTED ? p.node2[d] : p.node[d];
      iv[0][d] = iv[1][d] = nullptr;
    }
  }
  float wt[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) wt[r] = p.w_slot >= 0 ? c_interp_w[p.w_slot][r] : p.w[r];
  // grid-stride over stars: the tables above are staged once per resident block, not once per 256 stars
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < p.n_star;
       s += (long long)gridDim.x * blockDim.x) {
    const int cl = p.scl ? p.scl[s] : 0;
    const double* org = p.origin + 3 * (long long)cl;
    const double pos[3] = {p.sx[s], p.sy[s], p.sz[s]};
    int lv = 0;
    if (NESTED) {
      lv = 1;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double lo = __dadd_rn(nd[1][d][0], org[d]), hi = __dadd_rn(nd[1][d][p.n2[d] - 1], org[d]);
        if (!(pos[d] >= lo && pos[d] <= hi)) lv = 0;
      }
    }
    int nn[3];
    const float4* rec[4];
#pragma unroll
    for (int d = 0; d < 3; ++d) nn[d] = (NESTED && lv) ? p.n2[d] : p.n[d];
#pragma unroll
    for (int r = 0; r < 4; ++r) rec[r] = (NESTED && lv) ? p.rec2[r] : p.rec[r];
    int c3[3];
    double t[3], u[3], inv3[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double* ndd = (NESTED && lv) ? nd[1][d] : nd[0][d];
      const double o = org[d];
      const int i = find_cell(ndd, nn[d], o, pos[d]);
      const double inv = SMEM_NODES ? ((NESTED && lv) ? iv[1][d][i] : iv[0][d][i])
                                    : __ddiv_rn(1.0, __dsub_rn(ndd[i + 1], ndd[i]));
      t[d] = __dmul_rn(__dsub_rn(pos[d], __dadd_rn(ndd[i], o)), inv);
      u[d] = __dsub_rn(1.0, t[d]);
      if (TENSOR) inv3[d] = inv;
      c3[d] = i;
    }

    const long long nyz = (long long)nn[1] * nn[2];
    const long long n_node = (long long)nn[0] * nyz + 1;  // + appended origin row
    const long long base = (long long)cl * n_node + ((long long)c3[0] * nn[1] + c3[1]) * nn[2] + c3[2];
    // corner order: (di,dj,dk) = 000,001,010,011,100,101,110,111
    // time blend in FP32, left to right: v = ((r0*w0 + r1*w1) + r2*w2) + r3*w3 ; a single plane is taken as is
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const long long off = base + (long long)(c >> 2) * nyz + (long long)((c >> 1) & 1) * nn[2] + (c & 1);
      float4 a = __ldg(rec[0] + off);
      if (p.n_rec > 1) {
        const float4 b = __ldg(rec[1] + off);
        a = make_float4(lerp_rn_f(a.x, wt[0], b.x, wt[1]), lerp_rn_f(a.y, wt[0], b.y, wt[1]),
                        lerp_rn_f(a.z, wt[0], b.z, wt[1]), lerp_rn_f(a.w, wt[0], b.w, wt[1]));
#pragma unroll
        for (int r = 2; r < 4; ++r) {
          if (r >= p.n_rec) break;
          const float4 e = __ldg(rec[r] + off);
          const float wr = wt[r];
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
      if (TENSOR && q < 3) {
        // d/dx: difference of the two x-face values; d/dy: y-edge differences blended in x; d/dz: z-edge
        // differences blended in y then x; each times the inverse cell size of its axis
        const double gx = __dmul_rn(__dsub_rn(c1, c0), inv3[0]);
        const double gy = __dmul_rn(lerp_rn(__dsub_rn(c01, c00), u[0], __dsub_rn(c11, c10), t[0]), inv3[1]);
        const double z0 = lerp_rn(__dsub_rn(w[1], w[0]), u[1], __dsub_rn(w[3], w[2]), t[1]);
        const double z1 = lerp_rn(__dsub_rn(w[5], w[4]), u[1], __dsub_rn(w[7], w[6]), t[1]);
        const double gz = __dmul_rn(lerp_rn(z0, u[0], z1, t[0]), inv3[2]);
        p.tensor[(long long)(0 + q) * p.n_star + s] = gx;  // T[x][q] = d a_q / dx
        p.tensor[(long long)(3 + q) * p.n_star + s] = gy;
        p.tensor[(long long)(6 + q) * p.n_star + s] = gz;
      }
    }
    if (p.cell) {
      p.cell[s] = c3[0];
      p.cell[p.n_star + s] = c3[1];
      p.cell[2 * p.n_star + s] = c3[2];
    }
    if (p.level) p.level[s] = lv;
  }
}

// ctx->knobs.interp_variant: 0: <=128 regs (2 blocks/SM), 1: <=80 regs (3), 2: <=64 regs (4: fastest, latency-bound)

static int interp_launch(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* const* rec, const float* w, int n_rec,
                         const double* sx, const double* sy, const double* sz, const int32_t* scl, int64_t n_star,
                         double* acc, double* pot, int32_t* cell, void* stream, const char* who,
                         const ocg_grid_desc* fine = nullptr, const float* const* rec_fine = nullptr,
                         double* tensor = nullptr, int32_t* level = nullptr, int w_slot = -1) {
  if (!grid) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: grid is NULL", who);
  if (n_star < 0) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: negative n_star", who);
  if (n_star == 0) return OCG_OK;
  for (int d = 0; d < 3; ++d)
    if (grid->n[d] < 2 || !grid->node_dev[d])
      return ocg_fail(ctx, OCG_ERR_INVALID, "%s: axis %d needs >= 2 nodes (got %d)", who, d, grid->n[d]);
  if (grid->n_cluster < 1 || !grid->origin_dev || !sx || !sy || !sz || !acc)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: NULL argument", who);
  if (n_rec < 1 || n_rec > 4) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: n_rec = %d outside [1,4]", who, n_rec);
  if (fine) {
    for (int d = 0; d < 3; ++d)
      if (fine->n[d] < 2 || !fine->node_dev[d])
        return ocg_fail(ctx, OCG_ERR_INVALID, "%s: fine axis %d needs >= 2 nodes (got %d)", who, d, fine->n[d]);
    if (!rec_fine) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: fine lattice without record planes", who);
    if (fine->n_cluster != grid->n_cluster)
      return ocg_fail(ctx, OCG_ERR_INVALID, "%s: coarse and fine lattices must hold the same clusters", who);
  }
  OcgDeviceGuard g(ctx->device);
  InterpParams p;
  for (int d = 0; d < 3; ++d) {
    p.n[d] = grid->n[d], p.node[d] = grid->node_dev[d];
    p.n2[d] = fine ? fine->n[d] : 0, p.node2[d] = fine ? fine->node_dev[d] : nullptr;
  }
  p.n_cluster = grid->n_cluster;
  p.origin = grid->origin_dev;
  for (int r = 0; r < 4; ++r) {
    p.rec[r] = r < n_rec ? reinterpret_cast<const float4*>(rec[r]) : nullptr;
    p.rec2[r] = (fine && r < n_rec) ? reinterpret_cast<const float4*>(rec_fine[r]) : nullptr;
    p.w[r] = r < n_rec ? w[r] : 0.f;
    if (r < n_rec && (!rec[r] || (fine && !rec_fine[r])))
      return ocg_fail(ctx, OCG_ERR_INVALID, "%s: record plane %d is NULL", who, r);
  }
  p.n_rec = n_rec;
  p.w_slot = w_slot;
  p.sx = sx, p.sy = sy, p.sz = sz;
  p.scl = scl;
  p.n_star = n_star;
  p.acc = acc, p.pot = pot, p.cell = cell, p.tensor = tensor, p.level = level;
  long long nn = (long long)grid->n[0] + grid->n[1] + grid->n[2];
  if (fine) nn += (long long)fine->n[0] + fine->n[1] + fine->n[2];
  // persistent launch: as many blocks as are resident at once (a multiple of the SM count), grid-stride inside
  const bool in_smem = nn <= 2048;
  const size_t smem = in_smem ? (size_t)nn * 2 * sizeof(double) : 0;
  typedef void (*interp_fn)(const InterpParams);
  static const interp_fn fns[3][2] = {
      {grid_interp_kernel<false, 2, false, false>, grid_interp_kernel<true, 2, false, false>},
      {grid_interp_kernel<false, 3, false, false>, grid_interp_kernel<true, 3, false, false>},
      {grid_interp_kernel<false, 4, false, false>, grid_interp_kernel<true, 4, false, false>}};
  // [nested][tensor][in_smem]: the two-level and tensor forms (any combination) use <= 128-register builds
  static const interp_fn fns_x[2][2][2] = {
      {{nullptr, nullptr}, {grid_interp_kernel<false, 2, false, true>, grid_interp_kernel<true, 2, false, true>}},
      {{grid_interp_kernel<false, 3, true, false>, grid_interp_kernel<true, 3, true, false>},
       {grid_interp_kernel<false, 2, true, true>, grid_interp_kernel<true, 2, true, true>}}};
  interp_fn fn = (fine || tensor) ? fns_x[fine ? 1 : 0][tensor ? 1 : 0][in_smem ? 1 : 0] : fns[ctx->knobs.interp_variant][in_smem ? 1 : 0];
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

// Two-level form (the reference's nested fine grid, grid_cartesian.py:34-53,71-91) + optional tidal tensor
// (gizmo_interface.py:719-756) + optional level/cell outputs.  fine == NULL: single level.
extern "C" int ocg_grid_interp_nested(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                                      const float* const* rec_coarse_dev, const float* const* rec_fine_dev,
                                      const double* weights, int32_t n_rec, const double* star_x_dev,
                                      const double* star_y_dev, const double* star_z_dev,
                                      const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                                      double* pot_out_dev, double* tensor_out_dev, int32_t* level_out_dev,
                                      int32_t* cell_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!rec_coarse_dev || !weights || n_rec < 1 || n_rec > 4)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_nested: need 1..4 record planes and weights");
  float w[4];
  for (int r = 0; r < n_rec; ++r) w[r] = (float)weights[r];
  return interp_launch(ctx, coarse, rec_coarse_dev, w, n_rec, star_x_dev, star_y_dev, star_z_dev, star_cluster_dev,
                       n_star, acc_out_dev, pot_out_dev, cell_out_dev, stream, "ocg_grid_interp_nested", fine,
                       rec_fine_dev, tensor_out_dev, level_out_dev);
}

// rec[index[i]] = float4(acc[0][i], acc[1][i], acc[2][i], pot[i]) — K2 pack with a scatter: lays the rows of the
// reference's point list (kept coarse points | fine lattice | origin) out as full-lattice node records.
__global__ void pack_planes_indexed_kernel(const double* __restrict__ acc, const double* __restrict__ pot,
                                           const long long* __restrict__ index, long long n, float4* __restrict__ rec) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  rec[index[i]] = make_float4((float)acc[i], (float)acc[n + i], (float)acc[2 * n + i], pot ? (float)pot[i] : 0.f);
}

extern "C" int ocg_pack_planes_indexed(ocg_ctx* ctx, const double* acc_dev, const double* pot_dev, int64_t n,
                                       const int64_t* index_dev, float* rec_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || (n > 0 && (!acc_dev || !rec_dev || !index_dev)))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_pack_planes_indexed: bad arguments");
  if (n == 0) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  pack_planes_indexed_kernel<<<nblocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
      acc_dev, pot_dev, reinterpret_cast<const long long*>(index_dev), n, reinterpret_cast<float4*>(rec_dev));
  OCG_CHECK_LAUNCH(ctx, "pack_planes_indexed_kernel");
  return OCG_OK;
}

// ---- graph-replayable form ------------------------------------------------------------------------------------
extern "C" int ocg_set_interp_weight_slots(ocg_ctx* ctx, const double* weights_host, int32_t first_slot, int32_t n_slots,
                                           void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!weights_host || first_slot < 0 || n_slots < 1 || first_slot + n_slots > OCG_W_SLOTS)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_set_interp_weight_slots: slots [%d, %d) outside [0, %d)", first_slot,
                    first_slot + n_slots, OCG_W_SLOTS);
  OcgDeviceGuard g(ctx->device);
  // staged through a small pinned ring owned by the ctx: the copy is then truly asynchronous (a pageable source makes the
  // runtime wait for the stream — i.e. for the previous graph replay — before it returns), and the source outlives the call
  constexpr int RING = 16;
  if (!ctx->w_ring) {
    OCG_CUDA(ctx, cudaHostAlloc((void**)&ctx->w_ring, sizeof(float) * RING * OCG_W_SLOTS * 4, cudaHostAllocDefault));
    for (int i = 0; i < RING; ++i) OCG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->w_ring_ev[i], cudaEventDisableTiming));
    ctx->w_ring_head = 0;
  }
  const unsigned e = ctx->w_ring_head++ % RING;
  if (ctx->w_ring_head > RING) OCG_CUDA(ctx, cudaEventSynchronize(ctx->w_ring_ev[e]));  // 16 steps behind: long complete
  float* w = ctx->w_ring + (size_t)e * OCG_W_SLOTS * 4;
  for (int k = 0; k < n_slots; ++k)
    for (int r = 0; r < 4; ++r) w[4 * k + r] = (float)weights_host[4 * k + r];
  OCG_CUDA(ctx, cudaMemcpyToSymbolAsync(c_interp_w, w, sizeof(float) * 4 * n_slots, sizeof(float) * 4 * first_slot,
                                        cudaMemcpyHostToDevice, (cudaStream_t)stream));
  OCG_CUDA(ctx, cudaEventRecord(ctx->w_ring_ev[e], (cudaStream_t)stream));
  return OCG_OK;
}

extern "C" int ocg_grid_interp_slot(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                                    const float* const* rec_coarse_dev, const float* const* rec_fine_dev, int32_t w_slot,
                                    int32_t n_rec, const double* star_x_dev, const double* star_y_dev,
                                    const double* star_z_dev, const int32_t* star_cluster_dev, int64_t n_star,
                                    double* acc_out_dev, double* pot_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!rec_coarse_dev || n_rec < 1 || n_rec > 4 || w_slot < 0 || w_slot >= OCG_W_SLOTS)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_slot: need 1..4 record planes and a weight slot in [0, %d)", OCG_W_SLOTS);
  const float w[4] = {0.f, 0.f, 0.f, 0.f};
  return interp_launch(ctx, coarse, rec_coarse_dev, w, n_rec, star_x_dev, star_y_dev, star_z_dev, star_cluster_dev, n_star,
                       acc_out_dev, pot_out_dev, nullptr, stream, "ocg_grid_interp_slot", fine, rec_fine_dev, nullptr, nullptr,
                       w_slot);
}

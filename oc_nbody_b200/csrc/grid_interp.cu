// K2 / K3: grid planes and the per-kick evaluation behind get_gravity_at_point
// (gizmo_interface.py:677-717), as trilinear-in-space + linear-in-time interpolation on the
// regular lattice of grid_cartesian.py:16-32,59-69.
//
// HBM-bound gather: one float4 record (ax, ay, az, phi) per node per snapshot, z-adjacent corners
// share a 32-byte sector; all arithmetic in FP64 with separately rounded mul/add (no FMA
// contraction) in a fixed order so the oracle (oracle/ocg_oracle.c: oracle_grid_interp) reproduces
// the bits.
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

__global__ void time_blend_kernel(const float4* __restrict__ ra, const float4* __restrict__ rb, double wb,
                                  long long n, double* __restrict__ acc, double* __restrict__ pot) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double wa = __dsub_rn(1.0, wb);
  float4 a = ra[i], b = rb ? rb[i] : a;
  acc[i] = lerp_rn((double)a.x, wa, (double)b.x, wb);
  acc[n + i] = lerp_rn((double)a.y, wa, (double)b.y, wb);
  acc[2 * n + i] = lerp_rn((double)a.z, wa, (double)b.z, wb);
  if (pot) pot[i] = lerp_rn((double)a.w, wa, (double)b.w, wb);
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
      reinterpret_cast<const float4*>(rec_a_dev), reinterpret_cast<const float4*>(rec_b_dev), w_b, n_node,
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
  const float4* rec_a;
  const float4* rec_b;
  double wb;
  const double* sx;
  const double* sy;
  const double* sz;
  const int* scl;
  long long n_star;
  double* acc;
  double* pot;
  int* cell;
  int nodes_in_smem;
};

// Cell along one axis: i = searchsorted(node + o, x, side='right') - 1, clamped to [0, n-2].
// The arithmetic estimate only seeds the search; the result is decided by comparisons against
// the (node[i] + o) values, which the oracle forms with the same single FP64 addition.
__device__ __forceinline__ int find_cell(const double* __restrict__ node, int n, double o, double x,
                                         double& x0, double& x1) {
  const double lo = __dadd_rn(node[0], o);
  const double hi = __dadd_rn(node[n - 1], o);
  int i = 0;
  if (x == x && n > 2) {  // not NaN
    double f = (x - lo) / (hi - lo) * (double)(n - 1);
    if (f >= (double)(n - 2)) i = n - 2;
    else if (f > 0.0) i = (int)f;
  }
  while (i > 0 && x < __dadd_rn(node[i], o)) --i;
  while (i < n - 2 && x >= __dadd_rn(node[i + 1], o)) ++i;
  x0 = __dadd_rn(node[i], o);
  x1 = __dadd_rn(node[i + 1], o);
  return i;
}

__global__ void __launch_bounds__(256) grid_interp_kernel(const InterpParams p) {
  extern __shared__ double s_nodes[];
  const double* nd[3];
  if (p.nodes_in_smem) {
    int off = 0;
    for (int d = 0; d < 3; ++d) {
      for (int i = threadIdx.x; i < p.n[d]; i += blockDim.x) s_nodes[off + i] = p.node[d][i];
      nd[d] = s_nodes + off;
      off += p.n[d];
    }
    __syncthreads();
  } else {
    nd[0] = p.node[0], nd[1] = p.node[1], nd[2] = p.node[2];
  }
  long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= p.n_star) return;

  const int cl = p.scl ? p.scl[s] : 0;
  const double* org = p.origin + 3 * (long long)cl;
  const double x = p.sx[s], y = p.sy[s], z = p.sz[s];
  double x0, x1, y0, y1, z0, z1;
  const int i = find_cell(nd[0], p.n[0], org[0], x, x0, x1);
  const int j = find_cell(nd[1], p.n[1], org[1], y, y0, y1);
  const int k = find_cell(nd[2], p.n[2], org[2], z, z0, z1);
  const double tx = __ddiv_rn(__dsub_rn(x, x0), __dsub_rn(x1, x0));
  const double ty = __ddiv_rn(__dsub_rn(y, y0), __dsub_rn(y1, y0));
  const double tz = __ddiv_rn(__dsub_rn(z, z0), __dsub_rn(z1, z0));
  const double ux = __dsub_rn(1.0, tx), uy = __dsub_rn(1.0, ty), uz = __dsub_rn(1.0, tz);

  const long long nyz = (long long)p.n[1] * p.n[2];
  const long long n_node = (long long)p.n[0] * nyz + 1;  // + appended origin row
  const long long base = (long long)cl * n_node + ((long long)i * p.n[1] + j) * p.n[2] + k;
  const float4* A = p.rec_a + base;
  const float4* B = p.rec_b ? p.rec_b + base : nullptr;
  const double wb = p.wb, wa = __dsub_rn(1.0, wb);

  // corner order: (di,dj,dk) = 000,001,010,011,100,101,110,111
  double v[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const long long off = (long long)(c >> 2) * nyz + (long long)((c >> 1) & 1) * p.n[2] + (c & 1);
    const float4 a = __ldg(A + off);
    if (B) {
      const float4 b = __ldg(B + off);
      v[c][0] = lerp_rn((double)a.x, wa, (double)b.x, wb);
      v[c][1] = lerp_rn((double)a.y, wa, (double)b.y, wb);
      v[c][2] = lerp_rn((double)a.z, wa, (double)b.z, wb);
      v[c][3] = lerp_rn((double)a.w, wa, (double)b.w, wb);
    } else {
      v[c][0] = a.x, v[c][1] = a.y, v[c][2] = a.z, v[c][3] = a.w;
    }
  }
  const int ncomp = p.pot ? 4 : 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q >= ncomp) break;
    const double c00 = lerp_rn(v[0][q], uz, v[1][q], tz);
    const double c01 = lerp_rn(v[2][q], uz, v[3][q], tz);
    const double c10 = lerp_rn(v[4][q], uz, v[5][q], tz);
    const double c11 = lerp_rn(v[6][q], uz, v[7][q], tz);
    const double c0 = lerp_rn(c00, uy, c01, ty);
    const double c1 = lerp_rn(c10, uy, c11, ty);
    const double r = lerp_rn(c0, ux, c1, tx);
    if (q < 3) p.acc[(long long)q * p.n_star + s] = r;
    else p.pot[s] = r;
  }
  if (p.cell) {
    p.cell[s] = i;
    p.cell[p.n_star + s] = j;
    p.cell[2 * p.n_star + s] = k;
  }
}

extern "C" int ocg_grid_interp(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* rec_a_dev,
                               const float* rec_b_dev, double w_b, const double* star_x_dev,
                               const double* star_y_dev, const double* star_z_dev,
                               const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                               double* pot_out_dev, int32_t* cell_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!grid) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: grid is NULL");
  if (n_star < 0) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: negative n_star");
  if (n_star == 0) return OCG_OK;
  for (int d = 0; d < 3; ++d)
    if (grid->n[d] < 2 || !grid->node_dev[d])
      return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: axis %d needs >= 2 nodes (got %d)", d, grid->n[d]);
  if (grid->n_cluster < 1 || !grid->origin_dev || !rec_a_dev || !star_x_dev || !star_y_dev || !star_z_dev || !acc_out_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: NULL argument");
  if (!(w_b >= 0.0 && w_b <= 1.0))
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp: w_b = %g outside [0,1]", w_b);
  OcgDeviceGuard g(ctx->device);
  InterpParams p;
  for (int d = 0; d < 3; ++d) p.n[d] = grid->n[d], p.node[d] = grid->node_dev[d];
  p.n_cluster = grid->n_cluster;
  p.origin = grid->origin_dev;
  p.rec_a = reinterpret_cast<const float4*>(rec_a_dev);
  p.rec_b = reinterpret_cast<const float4*>(rec_b_dev);
  p.wb = rec_b_dev ? w_b : 0.0;
  p.sx = star_x_dev, p.sy = star_y_dev, p.sz = star_z_dev;
  p.scl = star_cluster_dev;
  p.n_star = n_star;
  p.acc = acc_out_dev, p.pot = pot_out_dev, p.cell = cell_out_dev;
  long long nn = (long long)grid->n[0] + grid->n[1] + grid->n[2];
  p.nodes_in_smem = nn <= 4096;
  size_t smem = p.nodes_in_smem ? (size_t)nn * sizeof(double) : 0;
  grid_interp_kernel<<<nblocks(n_star, 256), 256, smem, (cudaStream_t)stream>>>(p);
  OCG_CHECK_LAUNCH(ctx, "grid_interp_kernel");
  return OCG_OK;
}

// K7: get_gravity_at_point the way the reference itself evaluates it (gizmo_interface.py:651-717):
// per star, the `nclose` (150) nearest points of the evolved grid (cKDTree.query, :654,664) and a polyharmonic-spline
// RBF interpolant with an added polynomial (rbf.interpolate.RBFInterpolant(points, values, basis=phs3, order=5),
// :656-674; options.py:43-45) evaluated at the star; get_tidal_tensor_at_point (:719-756) differentiates the same
// interpolant.
//
//   s(x) = sum_m w_m phi(|x - y_m|) + sum_k c_k p_k(x),   phi(r) = r^3,  p_k = the 56 monomials of degree <= 5
//   [ K  P ] [w]   [d]
//   [ P' 0 ] [c] = [0]          K_mn = phi(|y_m - y_n|),  P_mk = p_k(y_m)            (206 x 206, symmetric)
//
// One CTA per star.  Coordinates are shifted to the star and scaled by the grid spacing (the interpolant is invariant
// under both; the system's condition number drops to ~7e4), so s(star) = u . [d; 0] with A u = [phi(|y_m|); e_0]:
// one solve serves every field component (ax, ay, az, phi), and three more right-hand sides give the gradient.
//   1. candidate window of lattice nodes around the star's cell (+ the appended origin row), exact nclose-nearest
//      selection by rank counting on (distance^2, point index), window grown until provably sufficient;
//   2. the matrix is factorised in FP32 in shared memory (LU, partial pivoting; 170 KB, no HBM traffic);
//   3. FP64 iterative refinement with the matrix entries regenerated on the fly from the coordinates
//      (residual in FP64, correction through the FP32 factors held by one warp per right-hand side in registers);
//   4. out_c = sum_m u_m field_c[id_m] in FP64.
#include "ocg_internal.cuh"

#include <math.h>

#define RBF_NMAX 206   /* nclose + number of monomials */
#define RBF_NMONO_MAX 56
#define RBF_NRHS_MAX 4
#define RBF_THREADS 256
#define RBF_SLOTS 7    /* ceil(RBF_NMAX / 32): rows per lane in the register-resident triangular solves */
#define RBF_MAXIT 12
#define RBF_CAND_MAX 4097 /* 16^3 lattice nodes + the origin row */

struct RbfParams {
  int n[3];
  int n_cluster;
  const double* node[3];
  const double* origin;   // [n_cluster][3]
  const double* field;    // [n_comp][n_cluster][n_node]
  long long comp_stride;  // n_cluster * n_node
  int n_comp, nclose, nmono, phs, include_origin, want_tensor;
  const double *sx, *sy, *sz;
  const int* scl;
  long long n_star;
  double* out;     // [n_comp][n_star]
  double* tensor;  // [3][n_comp][n_star]
  int* status;     // [n_star]
  long long* nb_out;  // [nclose][n_star] neighbour point indices (nullable)
  unsigned char pw[RBF_NMONO_MAX][3];
};

__device__ __forceinline__ int rbf_find_cell(const double* __restrict__ node, int n, double o, double x) {
  // searchsorted(node + o, x, side='right') - 1 clamped to [0, n-2] (the rule of K3)
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__dadd_rn(node[mid], o) <= x) lo = mid + 1;
    else hi = mid;
  }
  int i = lo - 1;
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  return i;
}

__device__ __forceinline__ double rbf_phi(double r2, int phs) {
  // r^phs for odd phs
  double r = sqrt(r2), v = r;
  for (int q = 1; q < phs; q += 2) v *= r2;
  return v;
}

__device__ __forceinline__ double rbf_mono(const double* __restrict__ Y, int ld, int m, const unsigned char* pw) {
  double v = 1.0;
  const double x = Y[m], y = Y[ld + m], z = Y[2 * ld + m];
  for (int a = 0; a < pw[0]; ++a) v *= x;
  for (int a = 0; a < pw[1]; ++a) v *= y;
  for (int a = 0; a < pw[2]; ++a) v *= z;
  return v;
}

// entry (i, j) of the saddle-point matrix in FP64
__device__ __forceinline__ double rbf_entry(int i, int j, int ncl, const double* __restrict__ Y, const RbfParams& p) {
  if (i < ncl && j < ncl) {
    const double dx = Y[i] - Y[j], dy = Y[RBF_NMAX + i] - Y[RBF_NMAX + j], dz = Y[2 * RBF_NMAX + i] - Y[2 * RBF_NMAX + j];
    return rbf_phi(dx * dx + dy * dy + dz * dz, p.phs);
  }
  if (i < ncl) return rbf_mono(Y, RBF_NMAX, i, p.pw[j - ncl]);
  if (j < ncl) return rbf_mono(Y, RBF_NMAX, j, p.pw[i - ncl]);
  return 0.0;
}

__global__ void __launch_bounds__(RBF_THREADS, 1) rbf_interp_kernel(const RbfParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int ncl = p.nclose, N = p.nclose + p.nmono;
  float* A = reinterpret_cast<float*>(smem);                 // [N][N] column-major; aliased by the candidate distances
  double* cand = reinterpret_cast<double*>(smem);            // [<= RBF_CAND_MAX]
  unsigned char* q = smem + sizeof(float) * RBF_NMAX * RBF_NMAX;
  double* Y = reinterpret_cast<double*>(q);                  // [3][RBF_NMAX] scaled coordinates relative to the star
  q += sizeof(double) * 3 * RBF_NMAX;
  double* B = reinterpret_cast<double*>(q);                  // [RBF_NRHS_MAX][RBF_NMAX] right-hand sides
  q += sizeof(double) * RBF_NRHS_MAX * RBF_NMAX;
  double* U = reinterpret_cast<double*>(q);                  // solutions
  q += sizeof(double) * RBF_NRHS_MAX * RBF_NMAX;
  double* R = reinterpret_cast<double*>(q);                  // residuals
  q += sizeof(double) * RBF_NRHS_MAX * RBF_NMAX;
  float* RP = reinterpret_cast<float*>(q);                   // reciprocal pivots
  q += sizeof(float) * RBF_NMAX;
  int* nbc = reinterpret_cast<int*>(q);                      // candidate ordinal of the m-th neighbour
  q += sizeof(int) * RBF_NMAX;
  int* perm = reinterpret_cast<int*>(q);                     // row permutation of the factorisation
  q += sizeof(int) * RBF_NMAX;
  long long* nb = reinterpret_cast<long long*>(q);           // point index of the m-th neighbour (within its cluster)
  q += sizeof(long long) * RBF_NMAX;
  __shared__ int s_lo[3], s_cnt[3], s_cell[3], s_piv, s_flag, s_conv[RBF_NRHS_MAX], s_C;
  __shared__ double s_p[3], s_o[3], s_h, s_r2max;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nrhs = p.want_tensor ? 4 : 1;
  const long long n_node = (long long)p.n[0] * p.n[1] * p.n[2] + 1;

  for (long long s = blockIdx.x; s < p.n_star; s += gridDim.x) {
    const int cl = p.scl ? p.scl[s] : 0;
    int status = 0;
    if (tid == 0) {
      s_p[0] = p.sx[s], s_p[1] = p.sy[s], s_p[2] = p.sz[s];
      double h = 0.0;
      for (int d = 0; d < 3; ++d) {
        s_o[d] = p.origin[3 * (long long)cl + d];
        s_cell[d] = rbf_find_cell(p.node[d], p.n[d], s_o[d], s_p[d]);
        const double hd = p.node[d][1] - p.node[d][0];
        h = hd > h ? hd : h;
      }
      s_h = h;
    }
    __syncthreads();

    // ---- 1. the nclose nearest grid points --------------------------------------------------------------
    for (int W = 8;; W += 4) {
      if (tid == 0) {
        int C = 1;
        for (int d = 0; d < 3; ++d) {
          const int cnt = p.n[d] < W ? p.n[d] : W;
          int lo = s_cell[d] - (W / 2 - 1);
          if (lo > p.n[d] - cnt) lo = p.n[d] - cnt;
          if (lo < 0) lo = 0;
          s_lo[d] = lo, s_cnt[d] = cnt;
          C *= cnt;
        }
        s_C = C + (p.include_origin ? 1 : 0);
      }
      __syncthreads();
      const int C = s_C, cy = s_cnt[1], cz = s_cnt[2], Clat = s_cnt[0] * cy * cz;
      for (int c = tid; c < C; c += RBF_THREADS) {
        double qx, qy, qz;
        if (c < Clat) {
          const int ix = c / (cy * cz), iy = (c / cz) % cy, iz = c % cz;
          qx = __dadd_rn(p.node[0][s_lo[0] + ix], s_o[0]);
          qy = __dadd_rn(p.node[1][s_lo[1] + iy], s_o[1]);
          qz = __dadd_rn(p.node[2][s_lo[2] + iz], s_o[2]);
        } else {
          qx = s_o[0], qy = s_o[1], qz = s_o[2];  // the appended origin row (grid_cartesian.py:66-67)
        }
        const double dx = __dadd_rn(qx, -s_p[0]), dy = __dadd_rn(qy, -s_p[1]), dz = __dadd_rn(qz, -s_p[2]);
        cand[c] = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      }
      __syncthreads();
      for (int c = tid; c < C; c += RBF_THREADS) {
        const double key = cand[c];
        int rank = 0;
        for (int j = 0; j < C; ++j) {
          const double v = cand[j];
          rank += (v < key || (v == key && j < c)) ? 1 : 0;
        }
        if (rank < ncl) nbc[rank] = c;
        if (rank == ncl - 1) s_r2max = key;
      }
      __syncthreads();
      // sufficient iff no excluded lattice node can be as close as the nclose-th neighbour
      bool ok = true, grown = false;
      for (int d = 0; d < 3; ++d) {
        if (s_lo[d] > 0) {
          const double b = s_p[d] - (p.node[d][s_lo[d] - 1] + s_o[d]);
          ok = ok && (b > 0.0 && b * b > s_r2max);
        }
        if (s_lo[d] + s_cnt[d] < p.n[d]) {
          const double b = (p.node[d][s_lo[d] + s_cnt[d]] + s_o[d]) - s_p[d];
          ok = ok && (b > 0.0 && b * b > s_r2max);
        }
        grown = grown || s_cnt[d] < p.n[d];
      }
      if (ok || !grown) break;
      if (W >= 16) {
        status |= 1;  // stencil truncated: the star is too far outside the grid for a 16-node window
        break;
      }
      __syncthreads();
    }
    // coordinates of the neighbours, shifted to the star and scaled by the spacing
    {
      const int cy = s_cnt[1], cz = s_cnt[2], Clat = s_cnt[0] * cy * cz;
      for (int m = tid; m < ncl; m += RBF_THREADS) {
        const int c = nbc[m];
        double qx, qy, qz;
        long long gi;
        if (c < Clat) {
          const int ix = s_lo[0] + c / (cy * cz), iy = s_lo[1] + (c / cz) % cy, iz = s_lo[2] + c % cz;
          qx = __dadd_rn(p.node[0][ix], s_o[0]), qy = __dadd_rn(p.node[1][iy], s_o[1]), qz = __dadd_rn(p.node[2][iz], s_o[2]);
          gi = ((long long)ix * p.n[1] + iy) * p.n[2] + iz;
        } else {
          qx = s_o[0], qy = s_o[1], qz = s_o[2];
          gi = n_node - 1;
        }
        nb[m] = gi;
        if (p.nb_out) p.nb_out[(long long)m * p.n_star + s] = gi;
        Y[m] = __ddiv_rn(__dadd_rn(qx, -s_p[0]), s_h);
        Y[RBF_NMAX + m] = __ddiv_rn(__dadd_rn(qy, -s_p[1]), s_h);
        Y[2 * RBF_NMAX + m] = __ddiv_rn(__dadd_rn(qz, -s_p[2]), s_h);
      }
    }
    __syncthreads();  // the candidate distances (aliased by A) are dead from here on

    // ---- 2. assemble (FP64 -> FP32) and factorise ----------------------------------------------------------
    for (int e = tid; e < N * N; e += RBF_THREADS) {
      const int j = e / N, i = e - j * N;
      A[e] = (float)rbf_entry(i, j, ncl, Y, p);
    }
    for (int i = tid; i < N; i += RBF_THREADS) perm[i] = i;
    __syncthreads();
    if (warp == 0) {  // pivot of column 0
      float best = -1.f;
      int bi = 0;
      for (int i = lane; i < N; i += 32) {
        const float v = fabsf(A[i]);
        if (v > best) best = v, bi = i;
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
      }
      if (lane == 0) s_piv = bi;
    }
    if (tid == 0) s_flag = 0;
    for (int k = 0; k < N; ++k) {
      __syncthreads();  // pivot row of column k known; trailing update of step k-1 complete
      const int pv = s_piv;
      if (pv != k) {
        for (int j = tid; j < N; j += RBF_THREADS) {
          const float t = A[j * N + k];
          A[j * N + k] = A[j * N + pv];
          A[j * N + pv] = t;
        }
        if (tid == 0) {
          const int t = perm[k];
          perm[k] = perm[pv];
          perm[pv] = t;
        }
      }
      __syncthreads();
      const float piv = A[k * N + k];
      float rp = 1.0f / piv;
      if (!(fabsf(piv) > 1e-30f)) rp = 0.f;
      if (tid == 0) {
        RP[k] = rp;
        if (rp == 0.f) s_flag = 1;
      }
      // A_ij -= L_ik U_kj with L_ik = A[k][i] * rp (column k keeps the unscaled values), 4 columns per pass
      const float* __restrict__ lk = A + k * N;
      for (int j0 = k + 1 + 4 * warp; j0 < N; j0 += 4 * (RBF_THREADS / 32)) {
        float u[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) u[c] = (j0 + c < N) ? A[(j0 + c) * N + k] * rp : 0.f;
        float best = -1.f;
        int bi = k + 1;
        for (int i = k + 1 + lane; i < N; i += 32) {
          const float l = lk[i];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (j0 + c < N) {
              const float v = fmaf(-l, u[c], A[(j0 + c) * N + i]);
              A[(j0 + c) * N + i] = v;
              if (c == 0 && j0 == k + 1 && fabsf(v) > best) best = fabsf(v), bi = i;
            }
        }
        if (j0 == k + 1) {  // this warp has just finished column k+1: choose its pivot
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
          }
          if (lane == 0) s_piv = bi;
        }
      }
    }
    __syncthreads();
    if (s_flag) status |= 4;

    // ---- 3. right-hand sides and FP64 iterative refinement -------------------------------------------------
    for (int e = tid; e < nrhs * N; e += RBF_THREADS) {
      const int r = e / N, i = e - r * N;
      double v;
      if (i < ncl) {
        const double x = Y[i], y = Y[RBF_NMAX + i], z = Y[2 * RBF_NMAX + i];
        const double r2 = x * x + y * y + z * z;
        if (r == 0) v = rbf_phi(r2, p.phs);
        else {
          // d/dx_a of phi(|x - y_m|) at x = 0: -phs |y|^(phs-2) y_a
          const double ya = r == 1 ? x : (r == 2 ? y : z);
          double g = p.phs == 1 ? (r2 > 0.0 ? 1.0 / sqrt(r2) : 0.0) : sqrt(r2);
          for (int t = 3; t < p.phs; t += 2) g *= r2;
          v = -(double)p.phs * g * ya;
        }
      } else {
        const unsigned char* pw = p.pw[i - ncl];
        if (r == 0) v = (pw[0] + pw[1] + pw[2] == 0) ? 1.0 : 0.0;
        else v = (pw[r - 1] == 1 && pw[0] + pw[1] + pw[2] == 1) ? 1.0 : 0.0;
      }
      B[r * RBF_NMAX + i] = v;
      U[r * RBF_NMAX + i] = 0.0;
      R[r * RBF_NMAX + i] = v;
    }
    if (tid < RBF_NRHS_MAX) s_conv[tid] = 0;
    __syncthreads();
    double prev_z = 0.0;
    for (int it = 0; it < RBF_MAXIT; ++it) {
      if (it > 0) {
        // R = B - A U in FP64, entries regenerated from the coordinates (row i per thread, all right-hand sides)
        if (tid < N) {
          double acc[RBF_NRHS_MAX];
#pragma unroll
          for (int r = 0; r < RBF_NRHS_MAX; ++r) acc[r] = 0.0;
          for (int j = 0; j < N; ++j) {
            const double e = rbf_entry(tid, j, ncl, Y, p);
#pragma unroll
            for (int r = 0; r < RBF_NRHS_MAX; ++r)
              if (r < nrhs) acc[r] += e * U[r * RBF_NMAX + j];
          }
#pragma unroll
          for (int r = 0; r < RBF_NRHS_MAX; ++r)
            if (r < nrhs) R[r * RBF_NMAX + tid] = B[r * RBF_NMAX + tid] - acc[r];
        }
        __syncthreads();
      }
      if (warp < nrhs && !s_conv[warp]) {
        // z = (LU)^-1 P R for right-hand side `warp`; z lives in registers, rows i = slot*32 + lane
        float z[RBF_SLOTS];
#pragma unroll
        for (int t = 0; t < RBF_SLOTS; ++t) {
          const int i = t * 32 + lane;
          z[t] = i < N ? (float)R[warp * RBF_NMAX + perm[i]] : 0.f;
        }
        // forward: unit lower, column-oriented
#pragma unroll
        for (int kt = 0; kt < RBF_SLOTS; ++kt) {
          for (int src = 0; src < 32; ++src) {
            const int k = kt * 32 + src;
            if (k >= N) break;
            const float zk = __shfl_sync(0xffffffffu, z[kt], src) * RP[k];
            const float* __restrict__ col = A + k * N;
#pragma unroll
            for (int t = 0; t < RBF_SLOTS; ++t) {
              const int i = t * 32 + lane;
              if (t >= kt && i > k && i < N) z[t] = fmaf(-col[i], zk, z[t]);
            }
          }
        }
        // backward: upper, column-oriented
#pragma unroll
        for (int kt = RBF_SLOTS - 1; kt >= 0; --kt) {
          for (int src = 31; src >= 0; --src) {
            const int k = kt * 32 + src;
            if (k >= N) continue;
            const float zk = __shfl_sync(0xffffffffu, z[kt], src) * RP[k];
            if (lane == src) z[kt] = zk;
            const float* __restrict__ col = A + k * N;
#pragma unroll
            for (int t = 0; t < RBF_SLOTS; ++t) {
              const int i = t * 32 + lane;
              if (t <= kt && i < k) z[t] = fmaf(-col[i], zk, z[t]);
            }
          }
        }
        double zmax = 0.0, umax = 0.0;
#pragma unroll
        for (int t = 0; t < RBF_SLOTS; ++t) {
          const int i = t * 32 + lane;
          if (i < N) {
            const double u = U[warp * RBF_NMAX + i] + (double)z[t];
            U[warp * RBF_NMAX + i] = u;
            zmax = fmax(zmax, fabs((double)z[t]));
            umax = fmax(umax, fabs(u));
          }
        }
        for (int o = 16; o > 0; o >>= 1) {
          zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
          umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        }
        // converged when the correction is below 1e-10 of the solution (the next correction would be ~1e-3 of that);
        // a correction that stopped shrinking ends the iteration too (accepted if already small)
        int conv = 0;
        if (!(zmax == zmax) || !(umax == umax)) conv = 2;
        else if (zmax <= 1e-10 * umax) conv = 1;
        else if (it > 0 && zmax > 0.5 * prev_z) conv = zmax <= 1e-7 * umax ? 1 : 2;
        prev_z = zmax;
        if (lane == 0 && conv) s_conv[warp] = conv;
      }
      __syncthreads();
      bool done = true;
      for (int r = 0; r < nrhs; ++r) done = done && s_conv[r] != 0;
      if (done) break;
    }
    for (int r = 0; r < nrhs; ++r)
      if (s_conv[r] != 1) status |= 2;  // refinement did not converge (ill-conditioned stencil, e.g. clipped by the grid edge)

    // ---- 4. out_c = sum_m u_m field_c[id_m] ------------------------------------------------------------------
    for (int pair = warp; pair < nrhs * p.n_comp; pair += RBF_THREADS / 32) {
      const int r = pair / p.n_comp, c = pair - r * p.n_comp;
      const double* __restrict__ f = p.field + (long long)c * p.comp_stride + (long long)cl * n_node;
      double acc = 0.0;
      for (int m = lane; m < ncl; m += 32) acc += U[r * RBF_NMAX + m] * f[nb[m]];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (r == 0) p.out[(long long)c * p.n_star + s] = acc;
        else p.tensor[((long long)(r - 1) * p.n_comp + c) * p.n_star + s] = acc / s_h;
      }
    }
    if (tid == 0 && p.status) p.status[s] = status;
    __syncthreads();
  }
}

extern "C" int ocg_grid_interp_rbf(ocg_ctx* ctx, const ocg_grid_desc* grid, const double* field_dev, int32_t n_comp,
                                   int32_t nclose, int32_t order, int32_t phs, int32_t include_origin,
                                   const double* star_x_dev, const double* star_y_dev, const double* star_z_dev,
                                   const int32_t* star_cluster_dev, int64_t n_star, double* out_dev, double* tensor_out_dev,
                                   int32_t* status_out_dev, int64_t* neighbors_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!grid || n_star < 0) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: bad arguments");
  if (n_star == 0) return OCG_OK;
  if (!field_dev || !star_x_dev || !star_y_dev || !star_z_dev || !out_dev || !grid->origin_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: NULL argument");
  if (n_comp < 1 || n_comp > 4) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: n_comp = %d outside [1,4]", n_comp);
  if (order < 0 || order > 5) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: order = %d outside [0,5]", order);
  if (phs != 1 && phs != 3 && phs != 5 && phs != 7)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: basis phs%d not supported (odd polyharmonic splines phs1/3/5/7)", phs);
  if (order < (phs - 1) / 2)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: phs%d needs order >= %d to be well posed", phs, (phs - 1) / 2);
  RbfParams p;
  int nm = 0;
  // monomials of total degree <= order, by degree
  for (int d = 0; d <= order; ++d)
    for (int a = d; a >= 0; --a)
      for (int b = d - a; b >= 0; --b) p.pw[nm][0] = (unsigned char)a, p.pw[nm][1] = (unsigned char)b, p.pw[nm][2] = (unsigned char)(d - a - b), ++nm;
  if (nclose < nm || nclose + nm > RBF_NMAX)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: nclose = %d must lie in [%d, %d] for order %d", nclose, nm,
                    RBF_NMAX - nm, order);
  long long n_lat = 1;
  for (int d = 0; d < 3; ++d) {
    if (grid->n[d] < 2 || !grid->node_dev[d]) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: bad grid axis %d", d);
    p.n[d] = grid->n[d], p.node[d] = grid->node_dev[d];
    n_lat *= grid->n[d];
  }
  if (n_lat + (include_origin ? 1 : 0) < nclose)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: the grid has fewer than nclose = %d points", nclose);
  p.n_cluster = grid->n_cluster < 1 ? 1 : grid->n_cluster;
  p.origin = grid->origin_dev, p.field = field_dev, p.comp_stride = (long long)p.n_cluster * (n_lat + 1);
  p.n_comp = n_comp, p.nclose = nclose, p.nmono = nm, p.phs = phs, p.include_origin = include_origin ? 1 : 0;
  p.want_tensor = tensor_out_dev ? 1 : 0;
  p.sx = star_x_dev, p.sy = star_y_dev, p.sz = star_z_dev, p.scl = star_cluster_dev, p.n_star = n_star;
  p.out = out_dev, p.tensor = tensor_out_dev, p.status = status_out_dev, p.nb_out = (long long*)neighbors_out_dev;
  OcgDeviceGuard g(ctx->device);
  size_t smem = sizeof(float) * RBF_NMAX * RBF_NMAX + sizeof(double) * 3 * RBF_NMAX + 3 * sizeof(double) * RBF_NRHS_MAX * RBF_NMAX +
                sizeof(float) * RBF_NMAX + 2 * sizeof(int) * RBF_NMAX + sizeof(long long) * RBF_NMAX;
  OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)rbf_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid_dim = (int)(n_star < ctx->sm_count ? n_star : ctx->sm_count);
  rbf_interp_kernel<<<grid_dim, RBF_THREADS, smem, (cudaStream_t)stream>>>(p);
  OCG_CHECK_LAUNCH(ctx, "rbf_interp_kernel");
  return OCG_OK;
}

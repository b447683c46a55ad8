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
#include "../../include/ocg_debug.h"

#include <math.h>

#define RBF_NMAX 206   /* nclose + number of monomials */
#define RBF_NMONO_MAX 56
#define RBF_NRHS_MAX 4
#define RBF_THREADS 256
#define RBF_MAXIT 12
#define RBF_CAND_MAX 4609 /* 16^3 lattice nodes + the origin row + an 8^3 window of the coarse level (nested grid) */
#define RBF_FAR 1e300     /* distance^2 of a candidate slot that holds no point */
#define RBF_NPHASE 6
#define RBF_LDA 212 /* column stride of the factor matrix: a multiple of 4 (16-byte aligned column segments for LDS.128), 4-way
                        bank conflicts only on the O(N^2) row-wise accesses */
#define RBF_NB 8    /* panel width of the blocked LU */

struct RbfParams {
  int n[3];
  int n_cluster;
  const double* node[3];
  const double* origin;   // [n_cluster][3]
  const double* field;    // [n_comp][n_cluster][n_node]
  long long comp_stride;  // n_cluster * n_node
  int n_comp, nclose, nmono, order, phs, include_origin, want_tensor, embedded;
  // nested grid, both levels searched (ocg_grid_interp_rbf_nested): `n`/`node` describe the FINE lattice; the kept coarse
  // points (grid_cartesian.py:71-81) are candidates too.  Rows of the reference's point list: kept coarse | fine | origin.
  int mixed;
  int cn[3];
  const double* cnode[3];
  const int* coarse_row;  // [cn0*cn1*cn2] row of a coarse lattice node in the point list, -1 = dropped (inside the fine box)
  long long fine_row0;    // first fine row
  long long n_point;      // rows of the point list (row stride of `field` per cluster)
  const double *sx, *sy, *sz;
  const int* scl;
  long long n_star;
  double* out;     // [n_comp][n_star]
  double* tensor;  // [3][n_comp][n_star]
  int* status;     // [n_star]
  long long* nb_out;  // [nclose][n_star] neighbour point indices (nullable)
  unsigned char pw[RBF_NMONO_MAX][3];
};

// cycles spent per phase by thread 0 of every CTA (select, assemble, factorise, residual, solve, output); debug only
__device__ unsigned long long g_rbf_cycles[RBF_NPHASE];

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

// Polyharmonic splines of rbf.basis (options.py:178-202): phs<k> = r^k for odd k, r^k log r for even k (0 at r = 0).
// (The package flips the sign of every other one to keep them conditionally positive definite; the interpolant does not
// depend on the sign.)  In coordinates scaled by h, r^k log r = h^k [y^k log y + log h * y^k]: the second part is a
// polynomial of degree k in the coordinates whose contribution the side conditions sum_j lambda_j q(y_j) = 0 reduce to a
// polynomial of degree <= k - order - 1, i.e. it is absorbed by the polynomial part when order >= k/2 — the same
// condition that makes the interpolant well posed — so y^k log y is the basis function used here.
// EVEN is a template parameter of the kernel: the odd splines (the reference's default phs3) keep their log-free code.
template <bool EVEN>
__device__ __forceinline__ double rbf_phi(double r2, int phs) {
  if (!EVEN) {
    double r = sqrt(r2), v = r;
    for (int q = 1; q < phs; q += 2) v *= r2;
    return v;
  }
  if (!(r2 > 0.0)) return 0.0;
  double v = r2;
  for (int q = 2; q < phs; q += 2) v *= r2;
  return 0.5 * v * log(r2);
}
// phi'(r) / r: the radial derivative over r, the factor of (x - y) in the gradient of phi(|x - y|)
template <bool EVEN>
__device__ __forceinline__ double rbf_dphi_over_r(double r2, int phs) {
  if (!EVEN) {
    double g = phs == 1 ? (r2 > 0.0 ? 1.0 / sqrt(r2) : 0.0) : sqrt(r2);
    for (int t = 3; t < phs; t += 2) g *= r2;
    return (double)phs * g;
  }
  if (!(r2 > 0.0)) return 0.0;
  double v = 1.0;
  for (int q = 2; q < phs; q += 2) v *= r2;
  return v * (0.5 * (double)phs * log(r2) + 1.0);
}

// Entries of the saddle-point matrix in FP64.  Y [3][ncl] scaled coordinates, PT [3][order+1][ncl] their powers.
template <bool EVEN>
__device__ __forceinline__ double rbf_k(int i, int j, int ncl, const double* __restrict__ Y, int phs) {
  const double dx = Y[i] - Y[j], dy = Y[ncl + i] - Y[ncl + j], dz = Y[2 * ncl + i] - Y[2 * ncl + j];
  return rbf_phi<EVEN>(dx * dx + dy * dy + dz * dz, phs);
}
__device__ __forceinline__ double rbf_p(int m, int k, int ncl, int np1, const double* __restrict__ PT, const RbfParams& p) {
  const unsigned char* pw = p.pw[k];
  return PT[(0 * np1 + pw[0]) * ncl + m] * PT[(1 * np1 + pw[1]) * ncl + m] * PT[(2 * np1 + pw[2]) * ncl + m];
}

template <bool EVEN>
__global__ void __launch_bounds__(RBF_THREADS, 1) rbf_interp_kernel(const RbfParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int ncl = p.nclose, N = p.nclose + p.nmono, np1 = p.order + 1;
  float* A = reinterpret_cast<float*>(smem);                 // [N][RBF_LDA] column-major; aliased by the candidate distances
  double* cand = reinterpret_cast<double*>(smem);            // [<= RBF_CAND_MAX]
  size_t a_bytes = sizeof(float) * (size_t)RBF_LDA * N;
  if (a_bytes < sizeof(double) * RBF_CAND_MAX) a_bytes = sizeof(double) * RBF_CAND_MAX;
  unsigned char* q = smem + ((a_bytes + 15) / 16) * 16;
  double* Y = reinterpret_cast<double*>(q);                  // [3][ncl] scaled coordinates relative to the star
  q += sizeof(double) * 3 * ncl;
  double* PT = reinterpret_cast<double*>(q);                 // [3][order+1][ncl] powers of the coordinates
  q += sizeof(double) * 3 * np1 * ncl;
  double* B = reinterpret_cast<double*>(q);                  // [RBF_NRHS_MAX][N] right-hand sides
  q += sizeof(double) * RBF_NRHS_MAX * N;
  double* U = reinterpret_cast<double*>(q);                  // solutions
  q += sizeof(double) * RBF_NRHS_MAX * N;
  double* R = reinterpret_cast<double*>(q);                  // residuals
  q += sizeof(double) * RBF_NRHS_MAX * N;
  long long* nb = reinterpret_cast<long long*>(q);           // point index of the m-th neighbour (within its cluster)
  q += sizeof(long long) * ncl;
  float* ZF = reinterpret_cast<float*>(q);                   // corrections (FP32 solves)
  q += sizeof(float) * RBF_NRHS_MAX * N;
  float* ZS = reinterpret_cast<float*>(q);                   // ... of the current block, scaled by the reciprocal pivots
  q += sizeof(float) * RBF_NRHS_MAX * N;
  float* RP = reinterpret_cast<float*>(q);                   // reciprocal pivots
  q += sizeof(float) * N;
  int* nbc = reinterpret_cast<int*>(q);                      // candidate ordinal of the m-th neighbour
  q += sizeof(int) * ncl;
  int* perm = reinterpret_cast<int*>(q);                     // row permutation of the factorisation
  __shared__ int s_lo[3], s_cnt[3], s_cell[3], s_flag, s_conv[RBF_NRHS_MAX], s_C, s_pv[RBF_NB], s_clo[3], s_ccnt[3];
  __shared__ unsigned s_wkey[RBF_THREADS / 32];
  __shared__ __align__(16) float s_prow[2][RBF_NB];
  __shared__ double s_p[3], s_o[3], s_h, s_r2max, s_zprev[RBF_NRHS_MAX];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nrhs = p.want_tensor ? 4 : 1;
  const long long n_node = p.mixed ? p.n_point : (long long)p.n[0] * p.n[1] * p.n[2] + 1;
#ifdef OCG_TUNING
  long long t_phase = clock64();
  auto phase_done = [&](int which) {
    if (tid == 0) {
      const long long t = clock64();
      atomicAdd(&g_rbf_cycles[which], (unsigned long long)(t - t_phase));
      t_phase = t;
    }
  };
#else
  auto phase_done = [&](int) {};
#endif

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
    // candidate order = point-list order (ties go to the lower row, as cKDTree's do): coarse window | fine window | origin

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
        int CC = p.mixed ? 1 : 0;
        for (int d = 0; d < 3; ++d) {
          s_clo[d] = 0, s_ccnt[d] = 0;
          if (p.mixed) {  // W/2 coarse nodes per axis around the star (4, 6, 8): grows with the fine window
            const int cc = rbf_find_cell(p.cnode[d], p.cn[d], s_o[d], s_p[d]);
            const int cnt = p.cn[d] < W / 2 ? p.cn[d] : W / 2;
            int lo = cc - (W / 4 - 1);
            if (lo > p.cn[d] - cnt) lo = p.cn[d] - cnt;
            if (lo < 0) lo = 0;
            s_clo[d] = lo, s_ccnt[d] = cnt;
            CC *= cnt;
          }
        }
        s_C = CC + C + (p.include_origin ? 1 : 0);
      }
      __syncthreads();
      const int Cc = s_ccnt[0] * s_ccnt[1] * s_ccnt[2];
      const int C = s_C, cy = s_cnt[1], cz = s_cnt[2], Clat = s_cnt[0] * cy * cz;
      for (int c = tid; c < C; c += RBF_THREADS) {
        double qx, qy, qz;
        bool present = true;
        if (c < Cc) {  // a node of the coarse level: a point of the list only if it was kept
          const int ccy = s_ccnt[1], ccz = s_ccnt[2];
          const int ix = s_clo[0] + c / (ccy * ccz), iy = s_clo[1] + (c / ccz) % ccy, iz = s_clo[2] + c % ccz;
          present = p.coarse_row[((long long)ix * p.cn[1] + iy) * p.cn[2] + iz] >= 0;
          qx = __dadd_rn(p.cnode[0][ix], s_o[0]), qy = __dadd_rn(p.cnode[1][iy], s_o[1]), qz = __dadd_rn(p.cnode[2][iz], s_o[2]);
        } else if (c < Cc + Clat) {
          const int f = c - Cc;
          const int ix = f / (cy * cz), iy = (f / cz) % cy, iz = f % cz;
          qx = __dadd_rn(p.node[0][s_lo[0] + ix], s_o[0]);
          qy = __dadd_rn(p.node[1][s_lo[1] + iy], s_o[1]);
          qz = __dadd_rn(p.node[2][s_lo[2] + iz], s_o[2]);
        } else {
          qx = s_o[0], qy = s_o[1], qz = s_o[2];  // the appended origin row (grid_cartesian.py:66-67)
        }
        const double dx = __dadd_rn(qx, -s_p[0]), dy = __dadd_rn(qy, -s_p[1]), dz = __dadd_rn(qz, -s_p[2]);
        cand[c] = present ? __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) : RBF_FAR;
      }
      if (tid == 0) s_r2max = RBF_FAR;  // stays there when the window holds fewer than nclose points: the window grows
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
      bool ok = true;
      for (int d = 0; d < 3; ++d) {
        if (s_lo[d] > 0) {
          const double b = s_p[d] - (p.node[d][s_lo[d] - 1] + s_o[d]);
          ok = ok && (b > 0.0 && b * b > s_r2max);
        }
        if (s_lo[d] + s_cnt[d] < p.n[d]) {
          const double b = (p.node[d][s_lo[d] + s_cnt[d]] + s_o[d]) - s_p[d];
          ok = ok && (b > 0.0 && b * b > s_r2max);
        }
      }
      if (p.mixed) {  // ... nor a coarse node beyond its window
        for (int d = 0; d < 3; ++d) {
          if (s_clo[d] > 0) {
            const double b = s_p[d] - (p.cnode[d][s_clo[d] - 1] + s_o[d]);
            ok = ok && (b > 0.0 && b * b > s_r2max);
          }
          if (s_clo[d] + s_ccnt[d] < p.cn[d]) {
            const double b = (p.cnode[d][s_clo[d] + s_ccnt[d]] + s_o[d]) - s_p[d];
            ok = ok && (b > 0.0 && b * b > s_r2max);
          }
        }
      }
      if (ok) break;
      if (W >= 16) {
        status |= 1;  // stencil truncated: the star is too far outside the grid for a 16-node window
        break;
      }
      __syncthreads();
    }
    if (!(s_r2max < RBF_FAR)) {
      // fewer than nclose points within reach (a star far outside the grid): no stencil, no value
      for (int e = tid; e < p.n_comp; e += RBF_THREADS) {
        p.out[(long long)e * p.n_star + s] = __longlong_as_double(0x7ff8000000000000ll);
        if (p.tensor)
          for (int r = 0; r < 3; ++r) p.tensor[((long long)r * p.n_comp + e) * p.n_star + s] = __longlong_as_double(0x7ff8000000000000ll);
      }
      if (tid == 0 && p.status) p.status[s] = status | 1;
      __syncthreads();
      continue;
    }
    if (p.embedded && !p.mixed) {
      // the lattice is the fine level of the reference's nested grid: every kept coarse point lies on or outside the
      // fine box, so the neighbours are all fine points iff the nclose-th one is closer than the box surface
      bool inside = true;
      for (int d = 0; d < 3; ++d) {
        const double b0 = s_p[d] - (p.node[d][0] + s_o[d]), b1 = (p.node[d][p.n[d] - 1] + s_o[d]) - s_p[d];
        const double b = b0 < b1 ? b0 : b1;
        inside = inside && (b > 0.0 && b * b > s_r2max);
      }
      if (!inside) status |= 8;
    }
    // coordinates of the neighbours, shifted to the star and scaled by the spacing, and their powers
    {
      const int cy = s_cnt[1], cz = s_cnt[2], Clat = s_cnt[0] * cy * cz, Cc = s_ccnt[0] * s_ccnt[1] * s_ccnt[2];
      for (int m = tid; m < ncl; m += RBF_THREADS) {
        const int c = nbc[m];
        double qd[3];
        long long gi;
        if (c < Cc) {
          const int ccy = s_ccnt[1], ccz = s_ccnt[2];
          const int ix = s_clo[0] + c / (ccy * ccz), iy = s_clo[1] + (c / ccz) % ccy, iz = s_clo[2] + c % ccz;
          qd[0] = __dadd_rn(p.cnode[0][ix], s_o[0]), qd[1] = __dadd_rn(p.cnode[1][iy], s_o[1]), qd[2] = __dadd_rn(p.cnode[2][iz], s_o[2]);
          gi = p.coarse_row[((long long)ix * p.cn[1] + iy) * p.cn[2] + iz];
        } else if (c < Cc + Clat) {
          const int f = c - Cc;
          const int ix = s_lo[0] + f / (cy * cz), iy = s_lo[1] + (f / cz) % cy, iz = s_lo[2] + f % cz;
          qd[0] = __dadd_rn(p.node[0][ix], s_o[0]), qd[1] = __dadd_rn(p.node[1][iy], s_o[1]), qd[2] = __dadd_rn(p.node[2][iz], s_o[2]);
          gi = (p.mixed ? p.fine_row0 : 0) + ((long long)ix * p.n[1] + iy) * p.n[2] + iz;
        } else {
          qd[0] = s_o[0], qd[1] = s_o[1], qd[2] = s_o[2];
          gi = n_node - 1;
        }
        nb[m] = gi;
        if (p.nb_out) p.nb_out[(long long)m * p.n_star + s] = gi;
        for (int d = 0; d < 3; ++d) {
          const double y = __ddiv_rn(__dadd_rn(qd[d], -s_p[d]), s_h);
          Y[d * ncl + m] = y;
          double v = 1.0;
          for (int a = 0; a < np1; ++a) {
            PT[(d * np1 + a) * ncl + m] = v;
            v *= y;
          }
        }
      }
    }
    __syncthreads();  // the candidate distances (aliased by A) are dead from here on
    phase_done(0);

    // ---- 2. assemble (FP64 -> FP32) and factorise ----------------------------------------------------------
    if (tid < ncl) {  // row of a data point: K | P
#pragma unroll 4
      for (int j = 0; j < ncl; ++j) A[j * RBF_LDA + tid] = (float)rbf_k<EVEN>(tid, j, ncl, Y, p.phs);
#pragma unroll 4
      for (int k = 0; k < p.nmono; ++k) A[(ncl + k) * RBF_LDA + tid] = (float)rbf_p(tid, k, ncl, np1, PT, p);
    } else if (tid < N) {  // row of a monomial: P' | 0
#pragma unroll 4
      for (int j = 0; j < ncl; ++j) A[j * RBF_LDA + tid] = (float)rbf_p(j, tid - ncl, ncl, np1, PT, p);
      for (int k = 0; k < p.nmono; ++k) A[(ncl + k) * RBF_LDA + tid] = 0.f;
    }
    for (int i = tid; i < N; i += RBF_THREADS) perm[i] = i;
    __syncthreads();
    phase_done(1);
    if (tid == 0) s_flag = 0;
    // Blocked right-looking LU with partial pivoting, panels of RBF_NB columns.  Thread `tid` owns row `tid`.
    //   panel   : the row's RBF_NB panel entries live in registers; per column one block-wide arg-max (warp shuffles +
    //             one value per warp), the pivot row is published through shared memory (which also performs the row
    //             exchange), the elimination inside the panel never touches shared memory;
    //   swaps   : the panel's row exchanges are applied to the other columns by one thread per column;
    //   U rows  : U[kb..kb+NB)[j] = L_panel^-1 A[kb..kb+NB)[j], one thread per column;
    //   trailing: A[i][j] -= sum_c L[i][kb+c] U[kb+c][j]: the RBF_NB multipliers of a row stay in registers and the
    //             U column segment is two broadcast LDS.128, so an element is read and written once per RBF_NB pivots
    //             (4 shared-memory wavefronts per 32 elements per panel instead of 3 per pivot).
    // Column k keeps the unscaled values A[k][i] below the diagonal; L_ik = A[k][i] * RP[k].
    for (int kb = 0; kb < N; kb += RBF_NB) {
      const int nb = N - kb < RBF_NB ? N - kb : RBF_NB;
      const bool act = tid >= kb && tid < N;
      float a[RBF_NB];
#pragma unroll
      for (int c = 0; c < RBF_NB; ++c) a[c] = (act && c < nb) ? A[(kb + c) * RBF_LDA + tid] : 0.f;
#pragma unroll
      for (int c = 0; c < RBF_NB; ++c) {
        if (c < nb) {  // uniform
          const int k = kb + c;
          // pivot = arg-max |a[c]| over rows >= k, on the key (magnitude bits without their low byte | 255 - row): one
          // REDUX per warp and eight keys per CTA instead of a shuffle tree (partial pivoting only needs a near-maximal
          // pivot; ties go to the lowest row)
          unsigned key = (act && tid >= k) ? ((__float_as_uint(fabsf(a[c])) & 0xffffff00u) | (255u - (unsigned)tid)) : 0u;
          key = __reduce_max_sync(0xffffffffu, key);
          if (lane == 0) s_wkey[warp] = key;
          __syncthreads();
#pragma unroll
          for (int w = 0; w < RBF_THREADS / 32; ++w) key = max(key, s_wkey[w]);
          const int pv = key ? 255 - (int)(key & 0xffu) : k;
          // slot 0 <- the pivot row (it becomes row k), slot 1 <- the old row k (it moves to row pv)
          if (tid == pv) {
#pragma unroll
            for (int c2 = 0; c2 < RBF_NB; ++c2) s_prow[0][c2] = a[c2];
          } else if (tid == k) {
#pragma unroll
            for (int c2 = 0; c2 < RBF_NB; ++c2) s_prow[1][c2] = a[c2];
          }
          __syncthreads();
          if (pv != k) {
            if (tid == k) {
#pragma unroll
              for (int c2 = 0; c2 < RBF_NB; ++c2) a[c2] = s_prow[0][c2];
            } else if (tid == pv) {
#pragma unroll
              for (int c2 = 0; c2 < RBF_NB; ++c2) a[c2] = s_prow[1][c2];
            }
          }
          const float piv = s_prow[0][c];
          float rp = 1.0f / piv;
          if (!(fabsf(piv) > 1e-30f)) rp = 0.f;
          if (tid == 0) {
            RP[k] = rp;
            if (rp == 0.f) s_flag = 1;
            s_pv[c] = pv;
            const int t = perm[k];
            perm[k] = perm[pv];
            perm[pv] = t;
          }
          if (act && tid > k) {
            const float l = a[c] * rp;
#pragma unroll
            for (int c2 = 0; c2 < RBF_NB; ++c2)
              if (c2 > c) a[c2] = fmaf(-l, s_prow[0][c2], a[c2]);
          }
        }
      }
      if (act) {
#pragma unroll
        for (int c = 0; c < RBF_NB; ++c)
          if (c < nb) A[(kb + c) * RBF_LDA + tid] = a[c];
      }
      __syncthreads();  // panel written back, RP / s_pv of the whole panel visible
      if (tid < N && (tid < kb || tid >= kb + nb)) {
        float* colp = A + tid * RBF_LDA;  // this thread's column: apply the panel's row exchanges in order
        for (int c = 0; c < nb; ++c) {
          const int k = kb + c, pv = s_pv[c];
          if (pv != k) {
            const float t = colp[k];
            colp[k] = colp[pv];
            colp[pv] = t;
          }
        }
        if (tid >= kb + nb) {  // ... and turn rows kb..kb+nb of it into U: forward substitution with the panel's unit-lower block
          float u[RBF_NB];
#pragma unroll
          for (int c = 0; c < RBF_NB; ++c) u[c] = c < nb ? colp[kb + c] : 0.f;
#pragma unroll
          for (int c = 1; c < RBF_NB; ++c)
#pragma unroll
            for (int c1 = 0; c1 < c; ++c1)
              if (c < nb) u[c] = fmaf(-(A[(kb + c1) * RBF_LDA + kb + c] * RP[kb + c1]), u[c1], u[c]);
#pragma unroll
          for (int c = 0; c < RBF_NB; ++c)
            if (c < nb) colp[kb + c] = u[c];
        }
      }
      __syncthreads();
      if (nb == RBF_NB && tid >= kb + RBF_NB && tid < N) {
        float l[RBF_NB];
#pragma unroll
        for (int c = 0; c < RBF_NB; ++c) l[c] = A[(kb + c) * RBF_LDA + tid] * RP[kb + c];
        int j = kb + RBF_NB;
        for (; j + 2 <= N; j += 2) {
          // (j * RBF_LDA + kb) * 4 bytes is a multiple of 16: RBF_LDA and kb are multiples of 4
          const float4 u0 = *reinterpret_cast<const float4*>(A + j * RBF_LDA + kb);
          const float4 u1 = *reinterpret_cast<const float4*>(A + j * RBF_LDA + kb + 4);
          const float4 v0 = *reinterpret_cast<const float4*>(A + (j + 1) * RBF_LDA + kb);
          const float4 v1 = *reinterpret_cast<const float4*>(A + (j + 1) * RBF_LDA + kb + 4);
          float x = A[j * RBF_LDA + tid], y = A[(j + 1) * RBF_LDA + tid];
          x = fmaf(-l[0], u0.x, x), y = fmaf(-l[0], v0.x, y);
          x = fmaf(-l[1], u0.y, x), y = fmaf(-l[1], v0.y, y);
          x = fmaf(-l[2], u0.z, x), y = fmaf(-l[2], v0.z, y);
          x = fmaf(-l[3], u0.w, x), y = fmaf(-l[3], v0.w, y);
          x = fmaf(-l[4], u1.x, x), y = fmaf(-l[4], v1.x, y);
          x = fmaf(-l[5], u1.y, x), y = fmaf(-l[5], v1.y, y);
          x = fmaf(-l[6], u1.z, x), y = fmaf(-l[6], v1.z, y);
          x = fmaf(-l[7], u1.w, x), y = fmaf(-l[7], v1.w, y);
          A[j * RBF_LDA + tid] = x, A[(j + 1) * RBF_LDA + tid] = y;
        }
        for (; j < N; ++j) {
          const float4 u0 = *reinterpret_cast<const float4*>(A + j * RBF_LDA + kb);
          const float4 u1 = *reinterpret_cast<const float4*>(A + j * RBF_LDA + kb + 4);
          float x = A[j * RBF_LDA + tid];
          x = fmaf(-l[0], u0.x, x), x = fmaf(-l[1], u0.y, x), x = fmaf(-l[2], u0.z, x), x = fmaf(-l[3], u0.w, x);
          x = fmaf(-l[4], u1.x, x), x = fmaf(-l[5], u1.y, x), x = fmaf(-l[6], u1.z, x), x = fmaf(-l[7], u1.w, x);
          A[j * RBF_LDA + tid] = x;
        }
      }
      __syncthreads();  // trailing update complete before the next panel is loaded
    }
    __syncthreads();
    if (s_flag) status |= 4;
    phase_done(2);

    // ---- 3. right-hand sides and FP64 iterative refinement -------------------------------------------------
    for (int e = tid; e < nrhs * N; e += RBF_THREADS) {
      const int r = e / N, i = e - r * N;
      double v;
      if (i < ncl) {
        const double x = Y[i], y = Y[ncl + i], z = Y[2 * ncl + i];
        const double r2 = x * x + y * y + z * z;
        if (r == 0) v = rbf_phi<EVEN>(r2, p.phs);
        else {
          // d/dx_a of phi(|x - y_m|) at x = 0: -(phi'(|y|) / |y|) y_a
          const double ya = r == 1 ? x : (r == 2 ? y : z);
          v = -rbf_dphi_over_r<EVEN>(r2, p.phs) * ya;
        }
      } else {
        const unsigned char* pw = p.pw[i - ncl];
        if (r == 0) v = (pw[0] + pw[1] + pw[2] == 0) ? 1.0 : 0.0;
        else v = (pw[r - 1] == 1 && pw[0] + pw[1] + pw[2] == 1) ? 1.0 : 0.0;
      }
      B[r * N + i] = v;
      U[r * N + i] = 0.0;
      R[r * N + i] = v;
    }
    if (tid < RBF_NRHS_MAX) s_conv[tid] = 0, s_zprev[tid] = 0.0;
    __syncthreads();
    int iters = 0;
    for (int it = 0; it < RBF_MAXIT; ++it) {
      iters = it + 1;
      if (it > 0) {
        // R = B - A U in FP64, entries regenerated from the coordinates (row i per thread, all right-hand sides)
        if (tid < N) {
          double acc[RBF_NRHS_MAX];
#pragma unroll
          for (int r = 0; r < RBF_NRHS_MAX; ++r) acc[r] = 0.0;
          if (tid < ncl) {
#pragma unroll 4
            for (int j = 0; j < ncl; ++j) {
              const double e = rbf_k<EVEN>(tid, j, ncl, Y, p.phs);
#pragma unroll
              for (int r = 0; r < RBF_NRHS_MAX; ++r)
                if (r < nrhs) acc[r] += e * U[r * N + j];
            }
#pragma unroll 4
            for (int k = 0; k < p.nmono; ++k) {
              const double e = rbf_p(tid, k, ncl, np1, PT, p);
#pragma unroll
              for (int r = 0; r < RBF_NRHS_MAX; ++r)
                if (r < nrhs) acc[r] += e * U[r * N + ncl + k];
            }
          } else {
#pragma unroll 4
            for (int j = 0; j < ncl; ++j) {
              const double e = rbf_p(j, tid - ncl, ncl, np1, PT, p);
#pragma unroll
              for (int r = 0; r < RBF_NRHS_MAX; ++r)
                if (r < nrhs) acc[r] += e * U[r * N + j];
            }
          }
#pragma unroll
          for (int r = 0; r < RBF_NRHS_MAX; ++r)
            if (r < nrhs) R[r * N + tid] = B[r * N + tid] - acc[r];
        }
        __syncthreads();
        phase_done(3);
      }
      // z = (LU)^-1 P R through the FP32 factors: 32-column blocks, the triangular diagonal block by one warp per
      // right-hand side (values in registers, one shuffle per column), the rows outside the block by all threads
      for (int e = tid; e < nrhs * N; e += RBF_THREADS) {
        const int r = e / N, i = e - r * N;
        ZF[e] = (float)R[r * N + perm[i]];
      }
      __syncthreads();
      for (int k0 = 0; k0 < N; k0 += 32) {  // forward, unit lower: L_ik = A[k][i] * RP[k]
        const int kw = N - k0 < 32 ? N - k0 : 32;
        if (warp < nrhs) {
          float z = lane < kw ? ZF[warp * N + k0 + lane] : 0.f;
          float lc[32];  // this lane's row of the diagonal block, fetched before the dependent chain starts
#pragma unroll
          for (int c = 0; c < 32; ++c) lc[c] = (c < kw && lane > c && lane < kw) ? A[(k0 + c) * RBF_LDA + k0 + lane] * RP[k0 + c] : 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) z = fmaf(-lc[c], __shfl_sync(0xffffffffu, z, c), z);
          if (lane < kw) ZF[warp * N + k0 + lane] = z, ZS[warp * N + k0 + lane] = z * RP[k0 + lane];
        }
        __syncthreads();
        const int m = N - k0 - kw;
        for (int e = tid; e < nrhs * m; e += RBF_THREADS) {
          const int r = e / m, i = k0 + kw + (e - r * m);
          float acc = 0.f;
          for (int c = 0; c < kw; ++c) acc = fmaf(A[(k0 + c) * RBF_LDA + i], ZS[r * N + k0 + c], acc);
          ZF[r * N + i] -= acc;
        }
        __syncthreads();
      }
      for (int k0 = ((N - 1) / 32) * 32; k0 >= 0; k0 -= 32) {  // backward, upper: U_kk = 1 / RP[k], U_ik = A[k][i]
        const int kw = N - k0 < 32 ? N - k0 : 32;
        if (warp < nrhs) {
          float z = lane < kw ? ZF[warp * N + k0 + lane] : 0.f;
          float uc[32];
          const float rpl = lane < kw ? RP[k0 + lane] : 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) uc[c] = (c < kw && lane < c) ? A[(k0 + c) * RBF_LDA + k0 + lane] : 0.f;
#pragma unroll
          for (int c = 31; c >= 0; --c) {
            if (lane == c) z *= rpl;  // z_c = z'_c / U_cc
            z = fmaf(-uc[c], __shfl_sync(0xffffffffu, z, c), z);
          }
          if (lane < kw) ZF[warp * N + k0 + lane] = z;
        }
        __syncthreads();
        for (int e = tid; e < nrhs * k0; e += RBF_THREADS) {
          const int r = e / k0, i = e - r * k0;
          float acc = 0.f;
          for (int c = 0; c < kw; ++c) acc = fmaf(A[(k0 + c) * RBF_LDA + i], ZF[r * N + k0 + c], acc);
          ZF[r * N + i] -= acc;
        }
        __syncthreads();
      }
      if (warp < nrhs) {
        double zmax = 0.0, umax = 0.0;
        for (int i = lane; i < N; i += 32) {
          const double z = (double)ZF[warp * N + i];
          const double u = U[warp * N + i] + z;
          U[warp * N + i] = u;
          zmax = fmax(zmax, fabs(z));
          umax = fmax(umax, fabs(u));
          if (!(z == z)) zmax = z;
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double oz = __shfl_xor_sync(0xffffffffu, zmax, o), ou = __shfl_xor_sync(0xffffffffu, umax, o);
          zmax = (oz > zmax || !(oz == oz)) ? oz : zmax;
          umax = fmax(umax, ou);
        }
        // converged when the correction is below 1e-10 of the solution (the next one would be ~1e-3 of that); a
        // correction that stopped shrinking ends the iteration too (accepted if already small)
        if (lane == 0) {
          int conv = 0;
          if (!(zmax == zmax) || !(umax == umax)) conv = 2;
          else if (zmax <= 1e-10 * umax) conv = 1;
          else if (it > 0 && zmax > 0.5 * s_zprev[warp]) conv = zmax <= 1e-7 * umax ? 1 : 2;
          s_zprev[warp] = zmax;
          s_conv[warp] = conv;
        }
      }
      __syncthreads();
      phase_done(4);
      bool done = true;
      for (int r = 0; r < nrhs; ++r) done = done && s_conv[r] != 0;
      if (done) break;
    }
    for (int r = 0; r < nrhs; ++r)
      if (s_conv[r] != 1) status |= 2;  // refinement did not converge (ill-conditioned stencil, e.g. clipped by the grid edge)
    status |= iters << 8;

    // ---- 4. out_c = sum_m u_m field_c[id_m] ------------------------------------------------------------------
    for (int pair = warp; pair < nrhs * p.n_comp; pair += RBF_THREADS / 32) {
      const int r = pair / p.n_comp, c = pair - r * p.n_comp;
      const double* __restrict__ f = p.field + (long long)c * p.comp_stride + (long long)cl * n_node;
      double acc = 0.0;
      for (int m = lane; m < ncl; m += 32) acc += U[r * N + m] * f[nb[m]];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (r == 0) p.out[(long long)c * p.n_star + s] = acc;
        else p.tensor[((long long)(r - 1) * p.n_comp + c) * p.n_star + s] = acc / s_h;
      }
    }
    if (tid == 0 && p.status) p.status[s] = status;
    __syncthreads();
    phase_done(5);
  }
}

static size_t rbf_smem_bytes(int N, int ncl, int np1) {
  size_t a = sizeof(float) * (size_t)RBF_LDA * N;
  if (a < sizeof(double) * RBF_CAND_MAX) a = sizeof(double) * RBF_CAND_MAX;
  a = ((a + 15) / 16) * 16;
  return a + sizeof(double) * 3 * ncl + sizeof(double) * 3 * np1 * ncl + 3 * sizeof(double) * RBF_NRHS_MAX * N +
         sizeof(long long) * ncl + 2 * sizeof(float) * RBF_NRHS_MAX * N + sizeof(float) * N + sizeof(int) * ncl + sizeof(int) * N;
}

// cycles per phase summed over CTAs since the last call (select, assemble, factorise, residual, solve, output)
extern "C" int ocg_debug_rbf_phase_cycles(ocg_ctx* ctx, double* out6) {
  if (!ctx || !out6) return OCG_ERR_INVALID;
  OcgDeviceGuard g(ctx->device);
  unsigned long long h[RBF_NPHASE], z[RBF_NPHASE] = {0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(h, g_rbf_cycles, sizeof(h)) != cudaSuccess) return OCG_ERR_CUDA;
  if (cudaMemcpyToSymbol(g_rbf_cycles, z, sizeof(z)) != cudaSuccess) return OCG_ERR_CUDA;
  for (int i = 0; i < RBF_NPHASE; ++i) out6[i] = (double)h[i];
  return OCG_OK;
}

static int rbf_launch(ocg_ctx* ctx, const char* who, RbfParams& p, const ocg_grid_desc* grid, const double* field_dev, int32_t n_comp,
                      int32_t nclose, int32_t order, int32_t phs, int32_t include_origin, const double* star_x_dev,
                      const double* star_y_dev, const double* star_z_dev, const int32_t* star_cluster_dev, int64_t n_star,
                      double* out_dev, double* tensor_out_dev, int32_t* status_out_dev, int64_t* neighbors_out_dev, void* stream) {
  if (!field_dev || !star_x_dev || !star_y_dev || !star_z_dev || !out_dev || !grid->origin_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: NULL argument", who);
  if (n_comp < 1 || n_comp > 4) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: n_comp = %d outside [1,4]", who, n_comp);
  if (order < 0 || order > 5) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: order = %d outside [0,5]", who, order);
  if (phs < 1 || phs > 8)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: basis phs%d not supported (polyharmonic splines phs1 .. phs8)", who, phs);
  if (order < phs / 2)  // odd k: (k - 1) / 2, even k: k / 2 (also what makes the even ones independent of the length unit)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: phs%d needs order >= %d to be well posed", who, phs, phs / 2);
  int nm = 0;
  // monomials of total degree <= order, by degree
  for (int d = 0; d <= order; ++d)
    for (int a = d; a >= 0; --a)
      for (int b = d - a; b >= 0; --b) p.pw[nm][0] = (unsigned char)a, p.pw[nm][1] = (unsigned char)b, p.pw[nm][2] = (unsigned char)(d - a - b), ++nm;
  if (nclose < nm || nclose + nm > RBF_NMAX)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: nclose = %d must lie in [%d, %d] for order %d", who, nclose, nm, RBF_NMAX - nm, order);
  long long n_lat = 1;
  for (int d = 0; d < 3; ++d) {
    if (grid->n[d] < 2 || !grid->node_dev[d]) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: bad grid axis %d", who, d);
    p.n[d] = grid->n[d], p.node[d] = grid->node_dev[d];
    n_lat *= grid->n[d];
  }
  if (!p.mixed && n_lat + (include_origin ? 1 : 0) < nclose)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: the grid has fewer than nclose = %d points", who, nclose);
  p.n_cluster = grid->n_cluster < 1 ? 1 : grid->n_cluster;
  p.origin = grid->origin_dev, p.field = field_dev;
  p.comp_stride = (long long)p.n_cluster * (p.mixed ? p.n_point : n_lat + 1);
  p.n_comp = n_comp, p.nclose = nclose, p.nmono = nm, p.order = order, p.phs = phs, p.include_origin = include_origin ? 1 : 0;
  p.want_tensor = tensor_out_dev ? 1 : 0;
  p.sx = star_x_dev, p.sy = star_y_dev, p.sz = star_z_dev, p.scl = star_cluster_dev, p.n_star = n_star;
  p.out = out_dev, p.tensor = tensor_out_dev, p.status = status_out_dev, p.nb_out = (long long*)neighbors_out_dev;
  OcgDeviceGuard g(ctx->device);
  const size_t smem = rbf_smem_bytes(nclose + nm, nclose, order + 1);
  if (smem > 227 * 1024)
    return ocg_fail(ctx, OCG_ERR_INVALID, "%s: nclose = %d, order = %d need %zu bytes of shared memory", who, nclose, order, smem);
  void (*kern)(const RbfParams) = (phs & 1) ? rbf_interp_kernel<false> : rbf_interp_kernel<true>;
  OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid_dim = (int)(n_star < ctx->sm_count ? n_star : ctx->sm_count);
  kern<<<grid_dim, RBF_THREADS, smem, (cudaStream_t)stream>>>(p);
  OCG_CHECK_LAUNCH(ctx, "rbf_interp_kernel");
  return OCG_OK;
}

extern "C" int ocg_grid_interp_rbf(ocg_ctx* ctx, const ocg_grid_desc* grid, const double* field_dev, int32_t n_comp,
                                   int32_t nclose, int32_t order, int32_t phs, int32_t include_origin, int32_t embedded,
                                   const double* star_x_dev, const double* star_y_dev, const double* star_z_dev,
                                   const int32_t* star_cluster_dev, int64_t n_star, double* out_dev, double* tensor_out_dev,
                                   int32_t* status_out_dev, int64_t* neighbors_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!grid || n_star < 0) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf: bad arguments");
  if (n_star == 0) return OCG_OK;
  RbfParams p;
  memset(&p, 0, sizeof(p));
  p.embedded = embedded ? 1 : 0;
  return rbf_launch(ctx, "ocg_grid_interp_rbf", p, grid, field_dev, n_comp, nclose, order, phs, include_origin, star_x_dev, star_y_dev,
                    star_z_dev, star_cluster_dev, n_star, out_dev, tensor_out_dev, status_out_dev, neighbors_out_dev, stream);
}

extern "C" int ocg_grid_interp_rbf_nested(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                                          const int32_t* coarse_row_dev, int64_t fine_row0, int64_t n_point, const double* field_dev,
                                          int32_t n_comp, int32_t nclose, int32_t order, int32_t phs, int32_t include_origin,
                                          const double* star_x_dev, const double* star_y_dev, const double* star_z_dev,
                                          const int32_t* star_cluster_dev, int64_t n_star, double* out_dev, double* tensor_out_dev,
                                          int32_t* status_out_dev, int64_t* neighbors_out_dev, void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!coarse || !fine || !coarse_row_dev || n_star < 0 || fine_row0 < 0 || n_point < 1)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf_nested: bad arguments");
  if (n_star == 0) return OCG_OK;
  RbfParams p;
  memset(&p, 0, sizeof(p));
  p.mixed = 1, p.embedded = 1;
  long long n_fine = 1;
  for (int d = 0; d < 3; ++d) {
    if (coarse->n[d] < 2 || !coarse->node_dev[d]) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf_nested: bad coarse axis %d", d);
    p.cn[d] = coarse->n[d], p.cnode[d] = coarse->node_dev[d];
    n_fine *= fine->n[d];
  }
  if (fine_row0 + n_fine + 1 != n_point)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_grid_interp_rbf_nested: point list of %lld rows is not %lld kept coarse + %lld fine + origin",
                    (long long)n_point, (long long)fine_row0, n_fine);
  p.coarse_row = coarse_row_dev, p.fine_row0 = fine_row0, p.n_point = n_point;
  return rbf_launch(ctx, "ocg_grid_interp_rbf_nested", p, fine, field_dev, n_comp, nclose, order, phs, include_origin, star_x_dev,
                    star_y_dev, star_z_dev, star_cluster_dev, n_star, out_dev, tensor_out_dev, status_out_dev, neighbors_out_dev, stream);
}

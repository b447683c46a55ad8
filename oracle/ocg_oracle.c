/*
 * ocg_oracle.c — CPU ORACLE for the oceanic gravity hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; the product (oc_nbody_b200/) never does and has no CPU fallback.
 *
 * PARITY UNPINNED BY THE REFERENCE: gusbeane/oc_nbody ships no tests, golden vectors or fixtures
 * (SURVEY.md §4, §8c), cannot be imported (analysis.py:573 SyntaxError; amuse/pykdgrav/rbf absent)
 * and the arithmetic of its field build lives in un-vendored, un-pinned pykdgrav.  What IS pinned:
 *   - the grid layout, single lattice and nested fine grid, against the real grid_cartesian.py imported
 *     in the build container (tests/golden/grid_reference.npz, grid_nested_reference.npz, made by
 *     tests/golden/make_golden.py);
 *   - the reference's time interpolation (scipy splrep/splev per grid point, runnable as-is):
 *     tests/golden/time_spline_reference.npz;
 *   - the spline kernel, restated from pykdgrav's published ForceKernel/PotentialKernel
 *     (M. Grudic, pykdgrav/pytreegrav kernel.py; cubic spline of Springel et al. 2001 with
 *     support radius h), checked for continuity/Newtonian limits and against scipy quadrature of
 *     the spline density in tests/test_cpu_oracle.py;
 *   - analytic known answers (two-body, shell theorem, homogeneous sphere, affine fields);
 *   - the reference's kNN + RBF-PHS spatial interpolation (oracle/__init__.py: rbf_interp*): against scipy's cKDTree and
 *     RBFInterpolator (the rbf author's port) live and through tests/golden/rbf_reference.npz;
 *   - the Hermite force loop and step (ph4 is absent: the scheme is restated from Makino & Aarseth 1992):
 *     jerk against a finite difference of the acceleration along the flow, 4th-order convergence and energy
 *     conservation on a Kepler orbit.
 *
 * Every function is the FP64 restatement of one step of the reference with the north_star's
 * algorithm substitutions (direct sum for the theta=0.5 tree, trilinear for RBF, linear-in-time
 * for the cubic spline, leapfrog for ph4) — see SURVEY.md §0 table.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off (no FMA contraction: oracle_grid_interp and the
 * kick/drift must round every mul and add separately, like the CUDA kernels do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KERNEL_PLUMMER 0
#define KERNEL_SPLINE 1

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The timed CPU baseline wants every host core even when the launcher exported OMP_NUM_THREADS=1 (torchrun does). */
int oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* (float)(pos - center): the FP32 rounding the GPU path applies after recentring in FP64
 * (SURVEY §7 H3).  pos [n][3] fp64 -> out [n][4] fp32 (x,y,z,m). */
void oracle_recentre(const double* pos, const double* mass, int64_t n, const double* center, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    out[4 * i + 0] = (float)(pos[3 * i + 0] - center[0]);
    out[4 * i + 1] = (float)(pos[3 * i + 1] - center[1]);
    out[4 * i + 2] = (float)(pos[3 * i + 2] - center[2]);
    out[4 * i + 3] = mass ? (float)mass[i] : 0.0f;
  }
}

/* pykdgrav ForceKernel(r, h): (mass fraction enclosed)/r^3 of a cubic-spline blob of support h. */
static inline double spline_force(double r, double h) {
  if (r >= h) return 1.0 / (r * r * r);
  double hinv = 1.0 / h, q = r * hinv;
  if (q <= 0.5) return (32.0 / 3.0 + q * q * (32.0 * q - 38.4)) * hinv * hinv * hinv;
  return (64.0 / 3.0 - 48.0 * q + 38.4 * q * q - (32.0 / 3.0) * q * q * q - (1.0 / 15.0) / (q * q * q)) * hinv * hinv * hinv;
}
/* pykdgrav PotentialKernel(r, h) (negative). */
static inline double spline_pot(double r, double h) {
  if (r >= h) return -1.0 / r;
  double hinv = 1.0 / h, q = r * hinv;
  if (q <= 0.5) return (-2.8 + q * q * (16.0 / 3.0 + q * q * (6.4 * q - 9.6))) * hinv;
  return (-3.2 + (1.0 / 15.0) / q + q * q * (32.0 / 3.0 + q * (-16.0 + q * (9.6 - (32.0 / 15.0) * q)))) * hinv;
}
double oracle_spline_force(double r, double h) { return spline_force(r, h); }
double oracle_spline_pot(double r, double h) { return spline_pot(r, h); }

/* Field build, direct sum (gizmo_interface.py:561,564,566 at theta -> 0):
 *   acc[c][t] = G sum_s m_s K(|x_s - x_t|, soft_s) (x_s - x_t)[c]     pot[t] = G sum_s m_s P(...)
 * Inputs are the SAME FP32-rounded arrays the GPU consumes; all arithmetic FP64.
 * Pair rule: Plummer: skipped iff r^2 + soft^2 == 0; spline: skipped iff r == 0. */
void oracle_field_direct(const float* src_xyzm, const float* src_soft, int64_t n_src, const float* tgt_xyzw,
                         int64_t n_tgt, int kernel, double G, double* acc, double* pot) {
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n_tgt; ++t) {
    const double tx = tgt_xyzw[4 * t], ty = tgt_xyzw[4 * t + 1], tz = tgt_xyzw[4 * t + 2];
    double ax = 0, ay = 0, az = 0, ph = 0;
    if (kernel == KERNEL_PLUMMER) {
      for (int64_t s = 0; s < n_src; ++s) {
        const double dx = src_xyzm[4 * s] - tx, dy = src_xyzm[4 * s + 1] - ty, dz = src_xyzm[4 * s + 2] - tz;
        const double m = src_xyzm[4 * s + 3];
        const double e = src_soft ? (double)src_soft[s] : 0.0;
        const double r2 = dx * dx + dy * dy + dz * dz + e * e;
        if (r2 > 0.0) {
          const double ri = 1.0 / sqrt(r2);
          const double f = m * ri * ri * ri;
          ax += f * dx, ay += f * dy, az += f * dz;
          ph -= m * ri;
        }
      }
    } else {
      for (int64_t s = 0; s < n_src; ++s) {
        const double dx = src_xyzm[4 * s] - tx, dy = src_xyzm[4 * s + 1] - ty, dz = src_xyzm[4 * s + 2] - tz;
        const double m = src_xyzm[4 * s + 3];
        const double h = src_soft ? (double)src_soft[s] : 0.0;
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > 0.0) {
          const double r = sqrt(r2);
          double f, p;
          if (r >= h) {
            const double ri = 1.0 / r;
            f = ri * ri * ri, p = -ri;
          } else {
            f = spline_force(r, h), p = spline_pot(r, h);
          }
          ax += m * f * dx, ay += m * f * dy, az += m * f * dz;
          ph += m * p;
        }
      }
    }
    acc[t] = G * ax, acc[n_tgt + t] = G * ay, acc[2 * n_tgt + t] = G * az;
    if (pot) pot[t] = G * ph;
  }
}

/* Condition numbers of the sums above: abs_out[c][t] = G * sum_s |term_c(s, t)| — the sum of the MAGNITUDES of the pair
 * terms that oracle_field_direct adds with signs.  FP32 pair arithmetic delivers every pair term to a few ulp of ITSELF
 * (north_star: "fp32-source, fp64-accumulate"), so |a_gpu - a_ref| <= eps_pair * abs_out is the backward-error bound the
 * parity tests use where terms cancel (tests/util.py::rel_err). */
void oracle_field_direct_abs(const float* src_xyzm, const float* src_soft, int64_t n_src, const float* tgt_xyzw,
                             int64_t n_tgt, int kernel, double G, double* abs_out) {
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n_tgt; ++t) {
    const double tx = tgt_xyzw[4 * t], ty = tgt_xyzw[4 * t + 1], tz = tgt_xyzw[4 * t + 2];
    double ax = 0, ay = 0, az = 0;
    for (int64_t s = 0; s < n_src; ++s) {
      const double dx = src_xyzm[4 * s] - tx, dy = src_xyzm[4 * s + 1] - ty, dz = src_xyzm[4 * s + 2] - tz;
      const double m = fabs((double)src_xyzm[4 * s + 3]);
      const double h = src_soft ? (double)src_soft[s] : 0.0;
      double r2 = dx * dx + dy * dy + dz * dz, f = 0.0;
      if (kernel == KERNEL_PLUMMER) {
        r2 += h * h;
        if (r2 > 0.0) {
          const double ri = 1.0 / sqrt(r2);
          f = ri * ri * ri;
        }
      } else if (r2 > 0.0) {
        const double r = sqrt(r2);
        f = r >= h ? 1.0 / (r2 * r) : fabs(spline_force(r, h));
      }
      ax += m * f * fabs(dx), ay += m * f * fabs(dy), az += m * f * fabs(dz);
    }
    abs_out[t] = G * ax, abs_out[n_tgt + t] = G * ay, abs_out[2 * n_tgt + t] = G * az;
  }
}

/* Throughput-oriented variant of the Plummer sum for the timed CPU baseline: identical maths, sources
 * pre-split into SoA FP64 so gcc can vectorise the inner loop.  Used by bench.py only. */
void oracle_field_direct_fast(const double* sx, const double* sy, const double* sz, const double* sm,
                              const double* se2, int64_t n_src, const double* tgt_xyz, int64_t n_tgt, double G,
                              double* acc) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t t = 0; t < n_tgt; ++t) {
    const double tx = tgt_xyz[3 * t], ty = tgt_xyz[3 * t + 1], tz = tgt_xyz[3 * t + 2];
    double ax = 0, ay = 0, az = 0;
#pragma omp simd reduction(+ : ax, ay, az)
    for (int64_t s = 0; s < n_src; ++s) {
      const double dx = sx[s] - tx, dy = sy[s] - ty, dz = sz[s] - tz;
      const double r2 = dx * dx + dy * dy + dz * dz + se2[s];
      const double ri = 1.0 / sqrt(r2);
      const double f = sm[s] * ri * ri * ri;
      ax += f * dx, ay += f * dy, az += f * dz;
    }
    acc[t] = G * ax, acc[n_tgt + t] = G * ay, acc[2 * n_tgt + t] = G * az;
  }
}

/* Frame subtraction (gizmo_interface.py:566,569-571). acc [3][n]. */
void oracle_frame_subtract(double* acc, int64_t n, int64_t row) {
  for (int c = 0; c < 3; ++c) {
    const double v = acc[c * n + row];
    for (int64_t i = 0; i < n; ++i) acc[c * n + i] = acc[c * n + i] - v;
  }
}

/* Cluster self-gravity (ph4 force loop, oc_code.py:218-229, eps2 = oc_code.py:225), batch of
 * segments.  pos [3][n] fp64, mass [n]; each segment recentred on its first particle then rounded
 * to FP32 exactly as the GPU path does; pair sums in FP64; self pair excluded by index; a pair
 * with r^2 + eps2 == 0 contributes nothing. Outputs only for [t0, t1). */
void oracle_self_gravity(const double* pos, const double* mass, int64_t n, const int64_t* seg_off, int32_t n_seg,
                         double eps2, double G, int64_t t0, int64_t t1, double* acc, double* pot) {
  float* r = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    if (b <= a) continue;
    const double cx = pos[a], cy = pos[n + a], cz = pos[2 * n + a];
    for (int64_t i = a; i < b; ++i) {
      r[4 * i] = (float)(pos[i] - cx), r[4 * i + 1] = (float)(pos[n + i] - cy), r[4 * i + 2] = (float)(pos[2 * n + i] - cz);
      r[4 * i + 3] = (float)mass[i];
    }
  }
  const double e2 = (double)(float)eps2;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    const int64_t lo = a > t0 ? a : t0, hi = b < t1 ? b : t1;
#pragma omp parallel for schedule(static)
    for (int64_t i = lo; i < hi; ++i) {
      double ax = 0, ay = 0, az = 0, ph = 0;
      const double tx = r[4 * i], ty = r[4 * i + 1], tz = r[4 * i + 2];
      for (int64_t j = a; j < b; ++j) {
        if (j == i) continue;
        const double dx = r[4 * j] - tx, dy = r[4 * j + 1] - ty, dz = r[4 * j + 2] - tz;
        const double r2 = dx * dx + dy * dy + dz * dz + e2;
        if (r2 > 0.0) {
          const double ri = 1.0 / sqrt(r2), m = r[4 * j + 3];
          const double f = m * ri * ri * ri;
          ax += f * dx, ay += f * dy, az += f * dz;
          ph -= m * ri;
        }
      }
      acc[i] = G * ax, acc[n + i] = G * ay, acc[2 * n + i] = G * az;
      if (pot) pot[i] = G * ph;
    }
  }
  free(r);
}

/* Hermite force loop: acceleration AND jerk (and potential) of the cluster's self-gravity — the arithmetic of
 * the ph4 worker itself (4th-order Hermite; oc_code.py:218-229, eps2 = oc_code.py:225).  Positions and
 * velocities of each segment recentred on its first particle, rounded to FP32 exactly as the GPU path does;
 * pair terms and sums in FP64:  jerk = G vel_to_len sum m [ w/r^3 - 3 (d.w) d / r^5 ],  r^2 = |d|^2 + eps2. */
void oracle_self_gravity_hermite(const double* pos, const double* vel, const double* mass, int64_t n, const int64_t* seg_off,
                                 int32_t n_seg, double eps2, double G, double vel_to_len, int64_t t0, int64_t t1, double* acc,
                                 double* jerk, double* pot) {
  float* r = (float*)malloc(sizeof(float) * 7 * (size_t)n);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    for (int64_t i = a; i < b; ++i) {
      for (int c = 0; c < 3; ++c) {
        r[7 * i + c] = (float)(pos[c * n + i] - pos[c * n + a]);
        r[7 * i + 4 + c] = (float)(vel[c * n + i] - vel[c * n + a]);
      }
      r[7 * i + 3] = (float)mass[i];
    }
  }
  const double e2 = (double)(float)eps2;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    const int64_t lo = a > t0 ? a : t0, hi = b < t1 ? b : t1;
#pragma omp parallel for schedule(static)
    for (int64_t i = lo; i < hi; ++i) {
      double A[3] = {0, 0, 0}, J[3] = {0, 0, 0}, ph = 0;
      const float* ti = r + 7 * i;
      for (int64_t j = a; j < b; ++j) {
        if (j == i) continue;
        const float* sj = r + 7 * j;
        const double d[3] = {(double)sj[0] - ti[0], (double)sj[1] - ti[1], (double)sj[2] - ti[2]};
        const double w[3] = {(double)sj[4] - ti[4], (double)sj[5] - ti[5], (double)sj[6] - ti[6]};
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + e2;
        if (r2 > 0.0) {
          const double ri = 1.0 / sqrt(r2), m = sj[3];
          const double f = m * ri * ri * ri;
          const double al = -3.0 * (d[0] * w[0] + d[1] * w[1] + d[2] * w[2]) * ri * ri;
          for (int c = 0; c < 3; ++c) A[c] += f * d[c], J[c] += f * (w[c] + al * d[c]);
          ph -= m * ri;
        }
      }
      for (int c = 0; c < 3; ++c) acc[c * n + i] = G * A[c], jerk[c * n + i] = G * vel_to_len * J[c];
      if (pot) pot[i] = G * ph;
    }
  }
  free(r);
}

/* Condition numbers of the cluster sums (see oracle_field_direct_abs): abs_acc[c][i] = G sum_j |m d_c / r^3| and, when
 * vel != NULL, abs_jerk[c][i] = G vel_to_len sum_j (|m w_c / r^3| + |3 m (d.w) d_c / r^5|); same inputs and roundings as
 * oracle_self_gravity[_hermite]. */
void oracle_self_gravity_abs(const double* pos, const double* vel, const double* mass, int64_t n, const int64_t* seg_off,
                             int32_t n_seg, double eps2, double G, double vel_to_len, int64_t t0, int64_t t1,
                             double* abs_acc, double* abs_jerk) {
  float* r = (float*)malloc(sizeof(float) * 7 * (size_t)n);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    for (int64_t i = a; i < b; ++i) {
      for (int c = 0; c < 3; ++c) {
        r[7 * i + c] = (float)(pos[c * n + i] - pos[c * n + a]);
        r[7 * i + 4 + c] = vel ? (float)(vel[c * n + i] - vel[c * n + a]) : 0.f;
      }
      r[7 * i + 3] = (float)mass[i];
    }
  }
  const double e2 = (double)(float)eps2;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    const int64_t lo = a > t0 ? a : t0, hi = b < t1 ? b : t1;
#pragma omp parallel for schedule(static)
    for (int64_t i = lo; i < hi; ++i) {
      double A[3] = {0, 0, 0}, J[3] = {0, 0, 0};
      const float* ti = r + 7 * i;
      for (int64_t j = a; j < b; ++j) {
        if (j == i) continue;
        const float* sj = r + 7 * j;
        const double d[3] = {(double)sj[0] - ti[0], (double)sj[1] - ti[1], (double)sj[2] - ti[2]};
        const double w[3] = {(double)sj[4] - ti[4], (double)sj[5] - ti[5], (double)sj[6] - ti[6]};
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + e2;
        if (r2 > 0.0) {
          const double ri = 1.0 / sqrt(r2), m = fabs((double)sj[3]);
          const double f = m * ri * ri * ri;
          const double al = 3.0 * fabs(d[0] * w[0] + d[1] * w[1] + d[2] * w[2]) * ri * ri;
          for (int c = 0; c < 3; ++c) A[c] += f * fabs(d[c]), J[c] += f * (fabs(w[c]) + al * fabs(d[c]));
        }
      }
      for (int c = 0; c < 3; ++c) {
        abs_acc[c * n + i] = G * A[c];
        if (abs_jerk) abs_jerk[c * n + i] = G * vel_to_len * J[c];
      }
    }
  }
  free(r);
}

/* Cell along one axis = searchsorted(node + o, x, side='right') - 1 clamped to [0, n-2]
 * (the bit-exact cell-selection rule; nodes = np.linspace of grid_cartesian.py:29-31,
 * shifted by evolve_grid's position, gizmo_interface.py:640-642 / grid_cartesian.py:55-57). */
static int find_cell(const double* node, int n, double o, double x) {
  int lo = 0, hi = n; /* first index with node+o > x */
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (node[mid] + o > x) hi = mid;
    else lo = mid + 1;
  }
  int i = lo - 1;
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  return i;
}

static inline double lerp2(double a, double wa, double b, double wb) { return a * wa + b * wb; }
/* FP32 blend, every operation rounded to FP32 (this file is compiled with -ffp-contract=off; the volatile
 * stores keep x87-style excess precision out of the picture on any host). */
static inline float lerp2f(float a, float wa, float b, float wb) {
  volatile float p = a * wa, q = b * wb;
  volatile float r = p + q;
  return r;
}

/* get_gravity_at_point (gizmo_interface.py:677-717) as trilinear-in-space interpolation of a time blend of
 * 1..4 record planes.  rec[r]: [n_cluster][nx*ny*nz+1][4] FP32 node records (ax,ay,az,phi); w[r] FP32 weights.
 * Arithmetic contract (same as the CUDA kernel):
 *   level  : (nested, grid_cartesian.py:34-53,71-91) fine iff node2[0]+o <= x <= node2[n2-1]+o on all axes
 *   cell   : searchsorted on node + origin (FP64)
 *   weight : t = (x - (node[i] + origin)) * inv[i], inv[i] = 1/(node[i+1] - node[i])            FP64
 *   time   : v = ((r0*w0 + r1*w1) + r2*w2) + r3*w3 on the FP32 records, each op rounded to FP32
 *   space  : z, y, x lerps of the 8 corner values, each op rounded to FP64
 *   tensor : T[3*i + j] = d a_j / d x_i of that trilinear form (gizmo_interface.py:719-756 T[i][j]) */
typedef struct {
  const int32_t* nn;
  const double* node[3];
  const float* const* rec;
} lattice_t;

static void grid_interp_core2(const lattice_t* coarse, const lattice_t* fine, const double* origin, const float* w,
                              int n_rec, const double* sx, const double* sy, const double* sz, const int32_t* scl,
                              int64_t n_star, double* acc, double* pot, double* tensor, int32_t* level, int32_t* cell) {
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < n_star; ++s) {
    const int cl = scl ? scl[s] : 0;
    const double* o = origin + 3 * (int64_t)cl;
    const double pos[3] = {sx[s], sy[s], sz[s]};
    int lv = 0;
    if (fine) {
      lv = 1;
      for (int d = 0; d < 3; ++d) {
        const double lo = fine->node[d][0] + o[d], hi = fine->node[d][fine->nn[d] - 1] + o[d];
        if (!(pos[d] >= lo && pos[d] <= hi)) lv = 0;
      }
    }
    const lattice_t* L = lv ? fine : coarse;
    const int nx = L->nn[0], ny = L->nn[1], nz = L->nn[2];
    const int64_t nyz = (int64_t)ny * nz, n_node = (int64_t)nx * nyz + 1;
    int c3[3];
    double t[3], u[3], inv[3];
    for (int d = 0; d < 3; ++d) {
      const double* nd = L->node[d];
      const int i = find_cell(nd, L->nn[d], o[d], pos[d]);
      inv[d] = 1.0 / (nd[i + 1] - nd[i]);
      t[d] = (pos[d] - (nd[i] + o[d])) * inv[d];
      u[d] = 1.0 - t[d];
      c3[d] = i;
    }
    const int64_t base = (int64_t)cl * n_node + ((int64_t)c3[0] * ny + c3[1]) * nz + c3[2];
    double v[8][4];
    for (int c = 0; c < 8; ++c) {
      const int64_t off = base + (int64_t)(c >> 2) * nyz + (int64_t)((c >> 1) & 1) * nz + (c & 1);
      for (int q = 0; q < 4; ++q) {
        float a = L->rec[0][4 * off + q];
        if (n_rec > 1) {
          a = lerp2f(a, w[0], L->rec[1][4 * off + q], w[1]);
          for (int r = 2; r < n_rec; ++r) {
            volatile float pr = L->rec[r][4 * off + q] * w[r];
            volatile float sm = a + pr;
            a = sm;
          }
        }
        v[c][q] = (double)a;
      }
    }
    const int ncomp = pot ? 4 : 3;
    for (int q = 0; q < ncomp; ++q) {
      const double c00 = lerp2(v[0][q], u[2], v[1][q], t[2]), c01 = lerp2(v[2][q], u[2], v[3][q], t[2]);
      const double c10 = lerp2(v[4][q], u[2], v[5][q], t[2]), c11 = lerp2(v[6][q], u[2], v[7][q], t[2]);
      const double c0 = lerp2(c00, u[1], c01, t[1]), c1 = lerp2(c10, u[1], c11, t[1]);
      const double r = lerp2(c0, u[0], c1, t[0]);
      if (q < 3) acc[(int64_t)q * n_star + s] = r;
      else pot[s] = r;
      if (tensor && q < 3) {
        const double gx = (c1 - c0) * inv[0];
        const double gy = lerp2(c01 - c00, u[0], c11 - c10, t[0]) * inv[1];
        const double z0 = lerp2(v[1][q] - v[0][q], u[1], v[3][q] - v[2][q], t[1]);
        const double z1 = lerp2(v[5][q] - v[4][q], u[1], v[7][q] - v[6][q], t[1]);
        const double gz = lerp2(z0, u[0], z1, t[0]) * inv[2];
        tensor[(int64_t)(0 + q) * n_star + s] = gx;
        tensor[(int64_t)(3 + q) * n_star + s] = gy;
        tensor[(int64_t)(6 + q) * n_star + s] = gz;
      }
    }
    if (cell) cell[s] = c3[0], cell[n_star + s] = c3[1], cell[2 * n_star + s] = c3[2];
    if (level) level[s] = lv;
  }
}

static void grid_interp_core(const int32_t* nn, const double* nodex, const double* nodey, const double* nodez,
                             const double* origin, const float* const* rec, const float* w, int n_rec,
                             const double* sx, const double* sy, const double* sz, const int32_t* scl, int64_t n_star,
                             double* acc, double* pot, int32_t* cell) {
  lattice_t L = {nn, {nodex, nodey, nodez}, rec};
  grid_interp_core2(&L, NULL, origin, w, n_rec, sx, sy, sz, scl, n_star, acc, pot, NULL, NULL, cell);
}

/* Two-level form + tidal tensor + level output (ocg_grid_interp_nested). nn2 == NULL: single level. */
void oracle_grid_interp_nested(const int32_t* nn, const double* nodex, const double* nodey, const double* nodez,
                               const int32_t* nn2, const double* node2x, const double* node2y, const double* node2z,
                               const double* origin, const float* const* rec, const float* const* rec2,
                               const double* weights, int32_t n_rec, const double* sx, const double* sy,
                               const double* sz, const int32_t* scl, int64_t n_star, double* acc, double* pot,
                               double* tensor, int32_t* level, int32_t* cell) {
  float w[4] = {0, 0, 0, 0};
  for (int r = 0; r < n_rec; ++r) w[r] = (float)weights[r];
  lattice_t C = {nn, {nodex, nodey, nodez}, rec};
  lattice_t F = {nn2, {node2x, node2y, node2z}, rec2};
  grid_interp_core2(&C, nn2 ? &F : NULL, origin, w, n_rec, sx, sy, sz, scl, n_star, acc, pot, tensor, level, cell);
}

/* Linear-in-time form (north_star): planes a, b and the weight wb of b; wa = fl32(1 - fl32(wb)). */
void oracle_grid_interp(const int32_t* nn, int32_t n_cluster, const double* nodex, const double* nodey,
                        const double* nodez, const double* origin, const float* rec_a, const float* rec_b, double wb,
                        const double* sx, const double* sy, const double* sz, const int32_t* scl, int64_t n_star,
                        double* acc, double* pot, int32_t* cell) {
  (void)n_cluster;
  const float wbf = rec_b ? (float)wb : 0.0f;
  volatile float waf_v = 1.0f - wbf;
  const float w[2] = {waf_v, wbf};
  const float* rec[2] = {rec_a, rec_b};
  grid_interp_core(nn, nodex, nodey, nodez, origin, rec, w, rec_b ? 2 : 1, sx, sy, sz, scl, n_star, acc, pot, cell);
}

/* General form: n_rec planes with FP64 weights rounded to FP32 (cubic B-spline in time: 4 coefficient planes). */
void oracle_grid_interp_multi(const int32_t* nn, int32_t n_cluster, const double* nodex, const double* nodey,
                              const double* nodez, const double* origin, const float* const* rec, const double* weights,
                              int32_t n_rec, const double* sx, const double* sy, const double* sz, const int32_t* scl,
                              int64_t n_star, double* acc, double* pot, int32_t* cell) {
  (void)n_cluster;
  float w[4] = {0, 0, 0, 0};
  for (int r = 0; r < n_rec; ++r) w[r] = (float)weights[r];
  grid_interp_core(nn, nodex, nodey, nodez, origin, rec, w, n_rec, sx, sy, sz, scl, n_star, acc, pot, cell);
}

/* rec[i] = (float)(ax, ay, az, phi) — the FP32 node records K3 reads. acc [3][n]. */
void oracle_pack_planes(const double* acc, const double* pot, int64_t n, float* rec) {
  for (int64_t i = 0; i < n; ++i) {
    rec[4 * i] = (float)acc[i], rec[4 * i + 1] = (float)acc[n + i], rec[4 * i + 2] = (float)acc[2 * n + i];
    rec[4 * i + 3] = pot ? (float)pot[i] : 0.0f;
  }
}

/* Linear time blend of node records (gizmo_interface.py:607-620 in linear form). */
void oracle_time_blend(const float* ra, const float* rb, double wb, int64_t n, double* acc, double* pot) {
  const float wbf = rb ? (float)wb : 0.0f;
  volatile float waf_v = 1.0f - wbf;
  const float waf = waf_v;
  for (int64_t i = 0; i < n; ++i) {
    const float* a = ra + 4 * i;
    float v[4];
    for (int q = 0; q < 4; ++q) v[q] = rb ? lerp2f(a[q], waf, rb[4 * i + q], wbf) : a[q];
    acc[i] = (double)v[0], acc[n + i] = (double)v[1], acc[2 * n + i] = (double)v[2];
    if (pot) pot[i] = (double)v[3];
  }
}

/* BRIDGE kick and leapfrog drift (oc_nbody.py:56 via amuse.couple.bridge): separately rounded. */
void oracle_kick(double* vel, const double* acc, int64_t n, double dt) {
  for (int64_t i = 0; i < 3 * n; ++i) vel[i] = vel[i] + acc[i] * dt;
}
void oracle_drift(double* pos, const double* vel, int64_t n, double dt, double vel_to_len) {
  for (int64_t i = 0; i < 3 * n; ++i) pos[i] = pos[i] + (vel[i] * dt) * vel_to_len;
}

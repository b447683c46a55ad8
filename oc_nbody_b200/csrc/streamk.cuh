// Stream-K work decomposition of the target-paired streaming kernels (K1/K4 direct_sum_tp_kernel, K6 hermite_tp_kernel).
//
// The work of a launch is a list of ROWS — one row per tile of targets — each against a run of source tiles
// (K1: every row against all fast tiles; K4/K6: a row against the tiles of its own cluster).  Counting one
// (target tile x source tile) product as a unit, CTA c of G takes the units [c*U/G, (c+1)*U/G): every CTA streams the
// same number of source tiles (+-1) whatever the number of rows, instead of whole (row x chunk) items dealt round robin
// (which left N = 65 536 self-gravity at 128 rows x 37 chunks over 296 CTAs = 16 rounds of 4 tiles for 55.4 tiles of
// ideal work: 15 % idle).  A row is therefore shared by at most ceil(G / rows) + 1 consecutive CTAs; each writes its
// FP64 partial sums into its own slot, takes a ticket, and the CTA that takes the last ticket of a row adds the slots
// IN SLOT ORDER (run-to-run deterministic) and writes the final field — no finish kernel, no (chunks x targets)
// partial buffer (K1 at configs[1]: 667 MB of partials down to 19 MB).
//
// PASSES (uniform mode, K1): with more source tiles than the L2 holds (configs[1]: 200 MB of tiles, 86 rows) CTAs that
// each stream their own 0.58 rows sit at 148 different places of the source array, and every tile is fetched from HBM
// again for nearly every row (measured: 9.1 GB of DRAM reads per launch for 0.2 GB of tiles).  The source tiles are
// therefore cut into passes of `tile_cap` tiles (32 MB); the CTAs split the units of pass 0, then those of pass 1, ...:
// at any time they all stream the same 32 MB, which stays in L2.  A row then has sharers in every pass; they all write
// slots (pass-major, CTA order within a pass) and take tickets of the same counter, and the last one adds the slots in
// slot order as before.  No barrier between passes: the alignment only matters for locality, not for correctness.
#pragma once
#include "ocg_internal.cuh"

__device__ __forceinline__ long long sk_nst(const StreamKParams& k) { return k.nst_uniform ? (long long)(*k.nst_uniform) : (long long)k.nst_value; }
__device__ __forceinline__ long long sk_units(const StreamKParams& k) {
  return k.rows ? k.row_prefix[k.n_rows] : (long long)k.n_rows * sk_nst(k);
}
__device__ __forceinline__ long long sk_row_start(const StreamKParams& k, int r) {
  return k.rows ? k.row_prefix[r] : (long long)r * sk_nst(k);
}
// the row holding unit u
__device__ __forceinline__ int sk_find_row(const StreamKParams& k, long long u) {
  if (!k.rows) return (int)(u / sk_nst(k));
  int lo = 0, hi = k.n_rows;  // largest r with prefix[r] <= u
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (k.row_prefix[mid] <= u) lo = mid;
    else hi = mid;
  }
  return lo;
}
// ---- uniform mode with an explicit tile count per row (the tiles of one pass) ----
__device__ __forceinline__ long long sk_pass_cap(const StreamKParams& k, long long nst_total) {
  return (k.rows || k.tile_cap <= 0 || k.tile_cap >= nst_total) ? (nst_total > 0 ? nst_total : 1) : (long long)k.tile_cap;
}
__device__ __forceinline__ long long sk_units(const StreamKParams& k, long long nst) {
  return k.rows ? k.row_prefix[k.n_rows] : (long long)k.n_rows * nst;
}
__device__ __forceinline__ long long sk_row_start(const StreamKParams& k, int r, long long nst) {
  return k.rows ? k.row_prefix[r] : (long long)r * nst;
}
__device__ __forceinline__ int sk_find_row(const StreamKParams& k, long long u, long long nst) {
  return k.rows ? sk_find_row(k, u) : (int)(u / nst);
}
// CTAs that take part: with fewer units than CTAs the surplus CTAs idle, so that the sharers of a row are consecutive
__device__ __forceinline__ long long sk_ctas(long long U, long long grid) { return U < grid ? U : grid; }
// first unit of CTA c, and the CTA that owns unit x:  b(c) = floor(c U / G);  c(x) = ceil((x + 1) G / U) - 1
__device__ __forceinline__ long long sk_first_unit(long long c, long long U, long long G) { return (c * U) / G; }
__device__ __forceinline__ int sk_cta_of(long long x, long long U, long long G) { return (int)(((x + 1) * G + U - 1) / U - 1); }

__device__ __forceinline__ int sk_sharers(long long rs, long long re, long long U, long long G) {
  return sk_cta_of(re - 1, U, G) - sk_cta_of(rs, U, G) + 1;
}

__device__ __forceinline__ void sk_row(const StreamKParams& k, int r, long long& tgt_begin, int& tgt_count, long long& tile_begin) {
  if (k.rows) {
    const OcgRow w = k.rows[r];
    tgt_begin = w.tgt_begin, tgt_count = w.tgt_count, tile_begin = w.tile_begin;
  } else {
    tgt_begin = (long long)r * k.ct;
    const long long rem = k.n_tgt - tgt_begin;
    tgt_count = rem < k.ct ? (int)rem : k.ct;
    tile_begin = 0;
  }
}

// Take a ticket of row r after this CTA's partial slot is written; true for the CTA that holds the last of `n_sharers`
// tickets (it then sees every other sharer's slot).  All threads of the CTA call it; s_flag is a shared int.
__device__ __forceinline__ bool sk_last_of_row(const StreamKParams& k, int r, int n_sharers, int* s_flag) {
  __threadfence();  // this thread's slot writes are visible device-wide before the ticket is taken
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(&k.tickets[r], 1u);
    const int last = old == (unsigned int)(n_sharers - 1);
    if (last) k.tickets[r] = 0u;  // leave the counter ready for the next launch (CUDA-graph replay included)
    *s_flag = last;
  }
  __syncthreads();
  const bool last = *s_flag != 0;
  if (last) __threadfence();
  return last;
}

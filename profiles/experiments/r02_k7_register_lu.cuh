// EXPERIMENT (round 2, not compiled into the library): register-resident LU for K7, measured SLOWER than the blocked
// shared-memory LU it was meant to replace (profiles/r02_bench_rbf_register_lu_experiment.json: 2.75 ms per 1 024-star kick
// against 2.56 ms; factorisation 178 us per star against 154 us).  Both are bound by the serial pivot chain (arg-max ->
// barrier -> row exchange -> barrier -> reciprocal -> update: ~1 500 cycles per pivot at 8 warps per SM), not by the
// shared-memory wavefronts of the trailing update that this version removes.  Kept as a record; see DESIGN.md, K7.
// Lesson recorded with it: __restrict__ on shared-memory pointers through which threads exchange data lets the compiler
// forward a thread's own store to its later load across __syncthreads(), silently dropping another thread's update.

// ---- register-resident LU ------------------------------------------------------------------------------------------
// The blocked LU below keeps the matrix in shared memory and is bound by its wavefronts (4 per 32 elements per panel).
// Here the 256 threads form a 16 x 16 grid and thread (ti, tj) holds the elements (ti + 16 a, tj + 16 b), a, b < 13, of the
// matrix in REGISTERS (2-D cyclic: 169 per thread, 208 x 208 >= 206 x 206), so that a pivot step is a rank-1 update of
// register tiles with 13 + 13 broadcast operands from shared memory.  The pivot loop runs over blocks of 16 pivots, the
// block index KB being a template parameter: within block KB the local index of the pivot column is KB (a compile-time
// register index), tile rows and columns below KB are finished, and only the (13 - KB)^2 live elements are updated —
// 13.1 k FMAs per thread for the whole factorisation instead of 206 x 169.  The only run-time register index is the
// local row of the pivot (partial pivoting picks it), resolved by an if-chain executed by the 16 threads that hold it.
// The result is written to shared memory as it is produced, in the layout the blocked LU leaves behind (U on and above
// the diagonal, the unscaled column below it, reciprocal pivots in RP, the row permutation in perm), so the triangular
// solves and the refinement are shared.  Two barriers per pivot.
#define RBF_TG 16
#define RBF_TR 13
static_assert(RBF_TG * RBF_TG == RBF_THREADS && RBF_TG * RBF_TR >= RBF_NMAX, "16 x 16 threads x 13 x 13 elements cover the matrix");

template <int KB>
// (no __restrict__ on the shared-memory pointers: the threads exchange data through them, and with it the compiler
// forwards a thread's own store to its later load of the same address across the barriers, missing the row exchange
// another thread applied in between)
__device__ __forceinline__ void rbf_lu_block(float (&c)[RBF_TR][RBF_TR], float* A, float* RP, int* perm, float* rowbuf, float* rowbuf2,
                                             unsigned* s_wkey, int* s_flag, const int N, const int tid, const int dbg) {
  const int ti = tid >> 4, tj = tid & 15, lane = tid & 31, warp = tid >> 5;
  const int kt_end = N - RBF_TG * KB < RBF_TG ? N - RBF_TG * KB : RBF_TG;
#pragma unroll 1
  for (int kt = 0; kt < kt_end; ++kt) {
    const int k = RBF_TG * KB + kt;
    if (dbg & 8) __syncthreads();
    // 1. pivot = arg-max over rows >= k, searched in the registers of the 16 holders of column k (same key as the blocked
    //    LU: magnitude without its low byte | 255 - row; ties go to the lowest row)
    const bool v1_store = (dbg & 32) != 0, v1_key = (dbg & 16) != 0;
    if (v1_store && tj == kt) {
#pragma unroll
      for (int a = KB; a < RBF_TR; ++a) {
        const int i = ti + RBF_TG * a;
        if (i >= k && i < N) A[k * RBF_LDA + i] = c[a][KB];
      }
    }
    if (v1_key) __syncthreads();
    unsigned key = 0u;
    if (v1_key) {
      key = (tid >= k && tid < N) ? ((__float_as_uint(fabsf(A[k * RBF_LDA + tid])) & 0xffffff00u) | (255u - (unsigned)tid)) : 0u;
    } else if (tj == kt) {
#pragma unroll
      for (int a = KB; a < RBF_TR; ++a) {
        const int i = ti + RBF_TG * a;
        const bool live = (a > KB || ti >= kt) && i < N;  // i >= k
        const unsigned kk = (__float_as_uint(fabsf(c[a][KB])) & 0xffffff00u) | (255u - (unsigned)i);
        key = max(key, live ? kk : 0u);
      }
    }
    key = __reduce_max_sync(0xffffffffu, key);
    if (lane == 0) s_wkey[warp] = key;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < RBF_THREADS / 32; ++w) key = max(key, s_wkey[w]);
    const int pv = key ? 255 - (int)(key & 0xffu) : k;
    const int pa = pv >> 4, pt = pv & 15;
    // 2. column k goes to the factor matrix with the exchange k <-> pv already applied (it stays there as the unscaled L
    //    column; the pivot itself is stored by thread 0 below)
    if (!v1_store && tj == kt) {
#pragma unroll
      for (int a = KB; a < RBF_TR; ++a) {
        const int i = ti + RBF_TG * a;
        const bool live = a > KB || ti >= kt;
        if (live && i != pv) A[k * RBF_LDA + (i == k ? pv : i)] = c[a][KB];
      }
    }
    // 3. the pivot row goes to rowbuf (it becomes row k = the U row), the old row k to rowbuf2 (it moves to row pv);
    //    the exchange is applied to the columns already in shared memory (0 .. k)
    // (a real switch: an if-chain is if-converted into 13 x 13 predicated stores that every warp issues at every pivot)
#define RBF_ROW_OUT(a)                                                          \
  case a:                                                                       \
    if (a >= KB) {                                                              \
      _Pragma("unroll") for (int b = KB; b < RBF_TR; ++b) rowbuf[tj + RBF_TG * b] = c[a][b]; \
    }                                                                           \
    break;
    if (ti == pt) {
      switch (pa) {
        RBF_ROW_OUT(0) RBF_ROW_OUT(1) RBF_ROW_OUT(2) RBF_ROW_OUT(3) RBF_ROW_OUT(4) RBF_ROW_OUT(5) RBF_ROW_OUT(6)
        RBF_ROW_OUT(7) RBF_ROW_OUT(8) RBF_ROW_OUT(9) RBF_ROW_OUT(10) RBF_ROW_OUT(11) RBF_ROW_OUT(12)
        default: break;
      }
    }
#undef RBF_ROW_OUT
    if (pv != k) {
      if (ti == kt) {
#pragma unroll
        for (int b = KB; b < RBF_TR; ++b) rowbuf2[tj + RBF_TG * b] = c[KB][b];
      }
      if (tid < k || (v1_store && tid == k)) {
        const float t = A[tid * RBF_LDA + k];
        A[tid * RBF_LDA + k] = A[tid * RBF_LDA + pv];
        A[tid * RBF_LDA + pv] = t;
      }
      if (tid == 0) {
        const int t = perm[k];
        perm[k] = perm[pv];
        perm[pv] = t;
      }
    }
    __syncthreads();
    // 4. rank-1 update of the live register tiles
#define RBF_ROW_IN(a)                                                           \
  case a:                                                                       \
    if (a >= KB) {                                                              \
      _Pragma("unroll") for (int b = KB; b < RBF_TR; ++b) c[a][b] = rowbuf2[tj + RBF_TG * b]; \
    }                                                                           \
    break;
    if (pv != k && ti == pt) {
      switch (pa) {
        RBF_ROW_IN(0) RBF_ROW_IN(1) RBF_ROW_IN(2) RBF_ROW_IN(3) RBF_ROW_IN(4) RBF_ROW_IN(5) RBF_ROW_IN(6)
        RBF_ROW_IN(7) RBF_ROW_IN(8) RBF_ROW_IN(9) RBF_ROW_IN(10) RBF_ROW_IN(11) RBF_ROW_IN(12)
        default: break;
      }
    }
#undef RBF_ROW_IN
    const float piv = rowbuf[k];
    float rp = (dbg & 4) ? 1.0f / piv : __fdividef(1.0f, piv);
    if (!(fabsf(piv) > 1e-30f)) rp = 0.f;
    if (tid == 0) {
      if (!v1_store) A[k * RBF_LDA + k] = piv;
      RP[k] = rp;
      if (rp == 0.f) *s_flag = 1;
    }
    if (tid > k && tid < N) A[tid * RBF_LDA + k] = rowbuf[tid];  // U row k
    // no predicates: rows and columns <= k (and the padding >= N) are dead registers, updating them is harmless
    float l[RBF_TR], u[RBF_TR];
#pragma unroll
    for (int a = KB; a < RBF_TR; ++a) {
      l[a] = A[k * RBF_LDA + ti + RBF_TG * a] * rp;
      if ((dbg & 2) && !(ti + RBF_TG * a > k && ti + RBF_TG * a < N)) l[a] = 0.f;
    }
#pragma unroll
    for (int b = KB; b < RBF_TR; ++b) {
      u[b] = rowbuf[tj + RBF_TG * b];
      if ((dbg & 2) && !(tj + RBF_TG * b > k && tj + RBF_TG * b < N)) u[b] = 0.f;
    }
#pragma unroll
    for (int a = KB; a < RBF_TR; ++a)
#pragma unroll
      for (int b = KB; b < RBF_TR; ++b) c[a][b] = fmaf(-l[a], u[b], c[a][b]);
  }
}


// ---- call site inside rbf_interp_kernel, in place of the blocked LU loop ----
/*
    if (p.lu_reg) {
      float c[RBF_TR][RBF_TR];
      {
        const int ti = tid >> 4, tj = tid & 15;
#pragma unroll
        for (int a = 0; a < RBF_TR; ++a)
#pragma unroll
          for (int b = 0; b < RBF_TR; ++b) {
            const int i = ti + RBF_TG * a, j = tj + RBF_TG * b;
            c[a][b] = (i < N && j < N) ? A[j * RBF_LDA + i] : 0.f;
          }
      }
      __syncthreads();  // every element is in registers before the first column is written back
#define RBF_LU_BLOCK(KB) \
  if (N > RBF_TG * KB) rbf_lu_block<KB>(c, A, RP, perm, ZF, ZS, s_wkey, &s_flag, N, tid, p.lu_reg);
      RBF_LU_BLOCK(0) RBF_LU_BLOCK(1) RBF_LU_BLOCK(2) RBF_LU_BLOCK(3) RBF_LU_BLOCK(4) RBF_LU_BLOCK(5) RBF_LU_BLOCK(6)
      RBF_LU_BLOCK(7) RBF_LU_BLOCK(8) RBF_LU_BLOCK(9) RBF_LU_BLOCK(10) RBF_LU_BLOCK(11) RBF_LU_BLOCK(12)
#undef RBF_LU_BLOCK
    }
*/

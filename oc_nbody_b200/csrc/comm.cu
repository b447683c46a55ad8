// Multi-GPU exchange over NVLink / NVSwitch peer memory: one process per GPU on one node, every rank's exchange WINDOW
// (cudaMalloc'd here, exported with cudaIpcGetMemHandle) mapped into every other rank (cudaIpcOpenMemHandle), kernels
// that load / store the peers' windows directly and synchronise through epoch flags in those windows.
//
// The reference has no multi-GPU path (its only parallel construct is a multiprocessing.Pool, gizmo_interface.py:600-605);
// the partition is BASELINE.json north_star's / SURVEY §8(e): K1 source-sharded + all-reduce of the partial fields,
// K4 star-sharded + all-gather of the positions every kick.
//
//   ocg_comm_allreduce_f64     fp64 sum over the ranks in RANK ORDER: deterministic, bit-identical on every rank
//                              (one persistent kernel: publish | reduce my slice from all peers | gather all slices)
//   ocg_self_gravity_sharded   K4 for this rank's block of stars: ONE kernel publishes the rank's FP64 positions, waits for
//                              the peers', reads their blocks over NVLink and writes the recentred FP32 source tiles
//                              (what all_gather + strided copy + pack did in three launches and an NCCL call), then the
//                              stream-K force kernel for the rank's target rows.
//
// Flag protocol.  Window header: flags[channel][rank] (64-bit epochs).  Rank q "signals" rank p on a channel by storing
// the call's epoch into p's flags[channel][q] (st.release.sys through the mapped pointer) after a system-scope fence that
// orders its data stores; rank p "waits" by polling its own flags (ld.acquire.sys).  Epochs are device-resident counters
// advanced by the kernels themselves, so a captured CUDA graph replays correctly.  Every rank issues the same sequence of
// calls (SPMD), each on its own GPU, so the polls terminate; a poll that does not within ~2 s (a peer died) sets the
// window's status word, the kernel finishes with whatever it has, and the next host call reports OCG_ERR_CUDA.
#include "ocg_internal.cuh"
#include "streamk.cuh"

#include <math.h>
#include <stdlib.h>

#define COMM_MAGIC 0x4f43474357494e31ull /* "OCGCWIN1" */
#define COMM_MAX_RANKS 16
#define COMM_CHANNELS 4 /* 0: position gather; 1, 2: all-reduce phases; 3: velocity gather (Hermite) */
#define COMM_HEADER_BYTES 4096
#define COMM_POLL_LIMIT_CYCLES (4000000000ll) /* ~2 s at 1.9 GHz */

struct CommHeader {  // lives at the start of every window, device memory
  unsigned long long flags[COMM_CHANNELS][COMM_MAX_RANKS];  // written by the peers
  unsigned long long epoch[COMM_CHANNELS];                  // local: last completed epoch per channel
  unsigned int ticket[8];                                   // local: CTA tickets of the exchange kernels
  unsigned int status;                                      // 0 ok, 1 a poll timed out
};

struct CommHandleBlob {  // OCG_COMM_HANDLE_BYTES = 128
  unsigned long long magic;
  int rank, nranks;
  long long window_bytes;
  int device, pad;
  cudaIpcMemHandle_t ipc;  // 64 bytes
  char reserve[32];
};
static_assert(sizeof(CommHandleBlob) == OCG_COMM_HANDLE_BYTES, "handle blob must be 128 bytes");
static_assert(sizeof(CommHeader) <= COMM_HEADER_BYTES, "header too large");

struct ocg_comm {
  int rank, nranks;
  long long window_bytes;            // data bytes after the header
  char* base[COMM_MAX_RANKS];        // mapped window base of every rank (base[rank] = own allocation)
  bool connected;
};

struct CommPeers {  // kernel argument: the mapped windows
  int rank, nranks;
  char* base[COMM_MAX_RANKS];
};

__device__ __forceinline__ CommHeader* hdr(char* base) { return reinterpret_cast<CommHeader*>(base); }
__device__ __forceinline__ double* win_data(char* base) { return reinterpret_cast<double*>(base + COMM_HEADER_BYTES); }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer(const double* p) {  // peer windows change between calls: never from a stale L1 line
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Signal every rank (self included) on `channel` with `epoch`; called by ONE thread after the CTA-level completion logic.
__device__ __forceinline__ void comm_signal_all(const CommPeers& pr, int channel, unsigned long long epoch) {
  __threadfence_system();
  for (int q = 0; q < pr.nranks; ++q) st_release_sys(&hdr(pr.base[q])->flags[channel][pr.rank], epoch);
}
// Wait (one thread) until every rank has signalled `epoch` on `channel`; false on timeout.
__device__ __forceinline__ bool comm_wait_all(const CommPeers& pr, int channel, unsigned long long epoch) {
  CommHeader* me = hdr(pr.base[pr.rank]);
  const long long t0 = clock64();
  for (int q = 0; q < pr.nranks; ++q) {
    while (ld_acquire_sys(&me->flags[channel][q]) < epoch) {
      if (clock64() - t0 > COMM_POLL_LIMIT_CYCLES) {
        me->status = 1u;
        return false;
      }
      __nanosleep(64);
    }
  }
  return true;
}
// CTA-level: all threads call; the LAST CTA of the grid to arrive (ticket slot `t`) returns true in every thread.
__device__ __forceinline__ bool comm_last_cta(CommHeader* me, int t, int* s_flag) {
  __threadfence_system();  // this thread's window stores are visible system-wide before the ticket is taken
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(&me->ticket[t], 1u);
    const int last = old == gridDim.x - 1;
    if (last) me->ticket[t] = 0u;
    *s_flag = last;
  }
  __syncthreads();
  const bool last = *s_flag != 0;
  __syncthreads();  // s_flag may be rewritten by the caller
  return last;
}

// ------------------------------------------------------------------------------ all-reduce ----
// In place fp64 sum over the ranks of buf[0..m), m <= cap.  The all-reduce works in the SECOND half of the window's data
// area (`half` doubles in), the force exchanges below in the first: a rank that runs ahead into an all-reduce must not
// overwrite the positions a slower peer is still gathering.  Second half: [0, cap) this rank's contribution, [cap, 2 cap)
// the slice this rank reduced.  Slice of rank q: [q*m/P, (q+1)*m/P).
__global__ void __launch_bounds__(512) comm_allreduce_kernel(CommPeers pr, double* __restrict__ buf, long long m, long long cap,
                                                             long long half) {
  __shared__ int s_flag;
  __shared__ unsigned long long s_epoch;
  CommHeader* me = hdr(pr.base[pr.rank]);
  if (threadIdx.x == 0) s_epoch = me->epoch[1] + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  const int P = pr.nranks, r = pr.rank;
  // phase 1: publish my contribution
  double* mine = win_data(pr.base[r]) + half;
  for (long long i = tid; i < m; i += nthr) mine[i] = buf[i];
  if (comm_last_cta(me, 0, &s_flag) && threadIdx.x == 0) comm_signal_all(pr, 1, e);
  // phase 2: reduce my slice over the ranks, in rank order
  if (threadIdx.x == 0) s_flag = comm_wait_all(pr, 1, e) ? 1 : 0;
  __syncthreads();
  const bool ok1 = s_flag != 0;
  __syncthreads();
  const long long s0 = (long long)r * m / P, s1 = (long long)(r + 1) * m / P;
  double* res = mine + cap;
  if (ok1)
    for (long long i = s0 + tid; i < s1; i += nthr) {
      double s = 0.0;
      for (int q = 0; q < P; ++q) s += ld_peer(win_data(pr.base[q]) + half + i);
      res[i - s0] = s;
    }
  if (comm_last_cta(me, 1, &s_flag) && threadIdx.x == 0) comm_signal_all(pr, 2, e);
  // phase 3: gather every rank's reduced slice
  if (threadIdx.x == 0) s_flag = comm_wait_all(pr, 2, e) ? 1 : 0;
  __syncthreads();
  const bool ok2 = s_flag != 0;
  __syncthreads();
  if (ok1 && ok2)
    for (int q = 0; q < P; ++q) {
      const long long q0 = (long long)q * m / P, q1 = (long long)(q + 1) * m / P;
      const double* src = win_data(pr.base[q]) + half + cap;
      for (long long i = q0 + tid; i < q1; i += nthr) buf[i] = ld_peer(src + (i - q0));
    }
  // the last CTA to finish advances the epoch (every CTA has read it by now)
  if (comm_last_cta(me, 2, &s_flag) && threadIdx.x == 0) me->epoch[1] = e;
}

// ------------------------------------------------------------- gather + pack (K4, sharded) ----
// Block partition of n stars over P ranks, sizes differing by at most one (distributed.shard_bounds).
__host__ __device__ __forceinline__ long long shard_begin(long long n, int P, int r) {
  const long long base = n / P, extra = n % P;
  return r * base + (r < extra ? r : extra);
}
__device__ __forceinline__ int shard_owner(long long n, int P, long long i) {
  const long long base = n / P, extra = n % P, cut = extra * (base + 1);
  return i < cut ? (int)(i / (base + 1)) : (int)(extra + (i - cut) / base);
}

// One kernel: (A) copy this rank's FP64 block [NARR_IN][n_local] into its window (buffer epoch & 1) and signal;
// (B) wait for every rank, then read ALL blocks through the mapped windows and lay out the FP32 source tiles and float4
// targets exactly as pack_cluster_kernel (self_gravity.cu) does for one segment: recentred on star 0 in FP64, rounded to
// FP32, positions scaled by the power of two `scale`.  mass_all is replicated on every rank.
__global__ void __launch_bounds__(256) comm_gather_pack_kernel(CommPeers pr, const double* __restrict__ pos_local,
                                                               const double* __restrict__ mass_all, long long n, long long max_local,
                                                               float e2, float scale, float* __restrict__ tiles,
                                                               float4* __restrict__ tgt) {
  __shared__ int s_flag;
  __shared__ unsigned long long s_epoch;
  __shared__ double s_c[3];
  CommHeader* me = hdr(pr.base[pr.rank]);
  if (threadIdx.x == 0) s_epoch = me->epoch[0] + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const int P = pr.nranks, r = pr.rank;
  const long long a = shard_begin(n, P, r), n_local = shard_begin(n, P, r + 1) - a;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  const long long buf_off = (long long)(e & 1ull) * 3 * max_local;  // double buffering: a fast rank may publish epoch e+1
                                                                    // while a slow one still reads epoch e
  double* mine = win_data(pr.base[r]) + buf_off;
  for (long long i = tid; i < 3 * n_local; i += nthr) {
    const long long c = i / n_local, k = i - c * n_local;
    mine[c * max_local + k] = pos_local[i];
  }
  if (comm_last_cta(me, 3, &s_flag) && threadIdx.x == 0) comm_signal_all(pr, 0, e);
  if (threadIdx.x == 0) s_flag = comm_wait_all(pr, 0, e) ? 1 : 0;
  __syncthreads();
  const bool ok = s_flag != 0;
  if (threadIdx.x < 3) s_c[threadIdx.x] = ok ? ld_peer(win_data(pr.base[0]) + buf_off + threadIdx.x * max_local) : 0.0;  // star 0
  __syncthreads();
  const long long total_tiles = (n + OCG_TS - 1) / OCG_TS;
  for (long long slot = tid; slot < total_tiles * OCG_TS; slot += nthr) {
    const long long tile = slot / OCG_TS;
    const int j = (int)(slot - tile * OCG_TS);
    float* T = tiles + tile * (long long)OCG_TILE_FLOATS;
    if (slot < n && ok) {
      const int q = shard_owner(n, P, slot);
      const double* src = win_data(pr.base[q]) + buf_off + (slot - shard_begin(n, P, q));
      const float x = (float)(ld_peer(src) - s_c[0]), y = (float)(ld_peer(src + max_local) - s_c[1]);
      const float z = (float)(ld_peer(src + 2 * max_local) - s_c[2]);
      const float m = (float)mass_all[slot];
      T[j] = x * scale, T[OCG_TS + j] = y * scale, T[2 * OCG_TS + j] = z * scale, T[3 * OCG_TS + j] = m;
      T[4 * OCG_TS + j] = e2 * scale * scale;
      tgt[slot] = make_float4(x, y, z, m);
    } else {
      T[j] = 0.f, T[OCG_TS + j] = 0.f, T[2 * OCG_TS + j] = 0.f, T[3 * OCG_TS + j] = 0.f, T[4 * OCG_TS + j] = 1.f;
    }
  }
  if (comm_last_cta(me, 4, &s_flag) && threadIdx.x == 0) me->epoch[0] = e;
}

// The same for K6 (hermite.cu): positions AND velocities are published ([6][n_local] per rank, channel 3 of the window
// header), and the tiles are the 7-array Hermite tiles, recentred on star 0 in position and velocity exactly as
// pack_hermite_kernel does for resident arrays.
__global__ void __launch_bounds__(256) comm_gather_pack_hermite_kernel(CommPeers pr, const double* __restrict__ pos_local,
                                                                       const double* __restrict__ vel_local,
                                                                       const double* __restrict__ mass_all, long long n,
                                                                       long long max_local, float scale, long long total_tiles,
                                                                       float* __restrict__ tiles, float4* __restrict__ tgt_pos,
                                                                       float4* __restrict__ tgt_vel) {
  __shared__ int s_flag;
  __shared__ unsigned long long s_epoch;
  __shared__ double s_c[6];
  CommHeader* me = hdr(pr.base[pr.rank]);
  if (threadIdx.x == 0) s_epoch = me->epoch[3] + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const int P = pr.nranks, r = pr.rank;
  const long long a = shard_begin(n, P, r), n_local = shard_begin(n, P, r + 1) - a;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  // its own region of the window, behind the two position buffers of the K4 exchange, double-buffered by epoch parity
  const long long buf_off = 6 * max_local + (long long)(e & 1ull) * 6 * max_local;
  double* mine = win_data(pr.base[r]) + buf_off;
  for (long long i = tid; i < 3 * n_local; i += nthr) {
    const long long c = i / n_local, k = i - c * n_local;
    mine[c * max_local + k] = pos_local[i];
    mine[(3 + c) * max_local + k] = vel_local[i];
  }
  if (comm_last_cta(me, 5, &s_flag) && threadIdx.x == 0) comm_signal_all(pr, 3, e);
  if (threadIdx.x == 0) s_flag = comm_wait_all(pr, 3, e) ? 1 : 0;
  __syncthreads();
  const bool ok = s_flag != 0;
  if (threadIdx.x < 6) s_c[threadIdx.x] = ok ? ld_peer(win_data(pr.base[0]) + buf_off + threadIdx.x * max_local) : 0.0;  // star 0
  __syncthreads();
  for (long long slot = tid; slot < total_tiles * HM_TS; slot += nthr) {
    const long long tile = slot / HM_TS;
    const int j = (int)(slot - tile * HM_TS);
    float* T = tiles + tile * (long long)HM_TILE_FLOATS;
    float v[HM_NARR] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (slot < n && ok) {
      const int q = shard_owner(n, P, slot);
      const double* src = win_data(pr.base[q]) + buf_off + (slot - shard_begin(n, P, q));
      const float x = (float)(ld_peer(src) - s_c[0]), y = (float)(ld_peer(src + max_local) - s_c[1]);
      const float z = (float)(ld_peer(src + 2 * max_local) - s_c[2]);
      const float vx = (float)(ld_peer(src + 3 * max_local) - s_c[3]), vy = (float)(ld_peer(src + 4 * max_local) - s_c[4]);
      const float vz = (float)(ld_peer(src + 5 * max_local) - s_c[5]);
      const float m = (float)mass_all[slot];
      v[0] = x * scale, v[1] = y * scale, v[2] = z * scale, v[3] = m, v[4] = vx, v[5] = vy, v[6] = vz;
      tgt_pos[slot] = make_float4(x, y, z, m);
      tgt_vel[slot] = make_float4(vx, vy, vz, 0.f);
    }
#pragma unroll
    for (int c = 0; c < HM_NARR; ++c) T[c * HM_TS + j] = v[c];
  }
  if (comm_last_cta(me, 6, &s_flag) && threadIdx.x == 0) me->epoch[3] = e;
}

// ------------------------------------------------------------------------------- host side ----
static int comm_check(ocg_ctx* ctx, const char* who) {
  if (!ctx) return OCG_ERR_INVALID;
  if (!ctx->comm) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: no communicator on this ctx (ocg_comm_create / ocg_comm_connect first)", who);
  if (!ctx->comm->connected) return ocg_fail(ctx, OCG_ERR_INVALID, "%s: ocg_comm_connect has not been called", who);
  return OCG_OK;
}

static CommPeers comm_peers(const ocg_comm* c) {
  CommPeers pr;
  pr.rank = c->rank, pr.nranks = c->nranks;
  for (int q = 0; q < COMM_MAX_RANKS; ++q) pr.base[q] = q < c->nranks ? c->base[q] : nullptr;
  return pr;
}

extern "C" int ocg_comm_create(ocg_ctx* ctx, int32_t rank, int32_t nranks, int64_t window_bytes, void* handle_out) {
  if (!ctx) return OCG_ERR_INVALID;
  if (ctx->comm) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_create: this ctx already has a communicator");
  if (nranks < 1 || nranks > COMM_MAX_RANKS || rank < 0 || rank >= nranks || window_bytes < 0 || !handle_out)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_create: rank %d of %d (at most %d ranks), window %lld bytes", rank, nranks,
                    COMM_MAX_RANKS, (long long)window_bytes);
  OcgDeviceGuard g(ctx->device);
  ocg_comm* c = (ocg_comm*)calloc(1, sizeof(ocg_comm));
  if (!c) return ocg_fail(ctx, OCG_ERR_NOMEM, "calloc failed");
  c->rank = rank, c->nranks = nranks;
  c->window_bytes = ((window_bytes + 511) / 512) * 512;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, COMM_HEADER_BYTES + (size_t)c->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, COMM_HEADER_BYTES + (size_t)c->window_bytes);
  CommHandleBlob blob;
  memset(&blob, 0, sizeof(blob));
  if (e == cudaSuccess && nranks > 1) e = cudaIpcGetMemHandle(&blob.ipc, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (p) cudaFree(p);
    free(c);
    return ocg_fail(ctx, OCG_ERR_CUDA, "ocg_comm_create: window of %lld bytes: %s", (long long)window_bytes, cudaGetErrorString(e));
  }
  c->base[rank] = (char*)p;
  blob.magic = COMM_MAGIC, blob.rank = rank, blob.nranks = nranks, blob.window_bytes = c->window_bytes, blob.device = ctx->device;
  memcpy(handle_out, &blob, sizeof(blob));
  c->connected = false;
  ctx->comm = c;
  return OCG_OK;
}

extern "C" int ocg_comm_connect(ocg_ctx* ctx, const void* all_handles) {
  if (!ctx) return OCG_ERR_INVALID;
  ocg_comm* c = ctx->comm;
  if (!c || !all_handles) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_connect: ocg_comm_create first, handles must not be NULL");
  if (c->connected) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_connect: already connected");
  OcgDeviceGuard g(ctx->device);
  const CommHandleBlob* h = (const CommHandleBlob*)all_handles;
  for (int q = 0; q < c->nranks; ++q) {
    if (h[q].magic != COMM_MAGIC || h[q].rank != q || h[q].nranks != c->nranks || h[q].window_bytes != c->window_bytes)
      return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_connect: handle %d is not rank %d's window of this communicator "
                      "(all ranks must pass the same nranks and window size, handles in rank order)", q, q);
  }
  for (int q = 0; q < c->nranks; ++q) {
    if (q == c->rank) continue;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h[q].ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return ocg_fail(ctx, OCG_ERR_CUDA, "ocg_comm_connect: mapping rank %d's window (device %d) failed: %s — the ranks must be "
                      "processes on one node whose GPUs have peer access (NVLink / NVSwitch)", q, h[q].device, cudaGetErrorString(e));
    }
    c->base[q] = (char*)p;
  }
  c->connected = true;
  return OCG_OK;
}

extern "C" int ocg_comm_destroy(ocg_ctx* ctx) {
  if (!ctx || !ctx->comm) return OCG_OK;
  OcgDeviceGuard g(ctx->device);
  ocg_comm* c = ctx->comm;
  cudaDeviceSynchronize();
  for (int q = 0; q < c->nranks; ++q) {
    if (!c->base[q]) continue;
    if (q == c->rank) cudaFree(c->base[q]);
    else cudaIpcCloseMemHandle(c->base[q]);
  }
  free(c);
  ctx->comm = nullptr;
  return OCG_OK;
}

extern "C" int ocg_comm_info(const ocg_ctx* ctx, int32_t* rank, int32_t* nranks, int64_t* window_bytes) {
  if (!ctx || !ctx->comm) return OCG_ERR_INVALID;
  if (rank) *rank = ctx->comm->rank;
  if (nranks) *nranks = ctx->comm->nranks;
  if (window_bytes) *window_bytes = ctx->comm->window_bytes;
  return OCG_OK;
}

// status word of the own window: a poll timed out in an earlier kernel (the stream is synchronised to read it)
extern "C" int ocg_comm_status(ocg_ctx* ctx, void* stream) {
  int rc = comm_check(ctx, "ocg_comm_status");
  if (rc) return rc;
  OcgDeviceGuard g(ctx->device);
  unsigned int st = 0;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaMemcpy(&st, &((CommHeader*)ctx->comm->base[ctx->comm->rank])->status, sizeof(st), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return ocg_fail(ctx, OCG_ERR_CUDA, "ocg_comm_status: %s", cudaGetErrorString(e));
  if (st) return ocg_fail(ctx, OCG_ERR_CUDA, "a peer did not arrive at an exchange within the poll limit (rank died or call sequences differ)");
  return OCG_OK;
}

extern "C" int ocg_comm_allreduce_f64(ocg_ctx* ctx, double* buf_dev, int64_t n, void* stream) {
  int rc = comm_check(ctx, "ocg_comm_allreduce_f64");
  if (rc) return rc;
  if (n < 0 || (n > 0 && !buf_dev)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_allreduce_f64: bad arguments");
  if (n == 0) return OCG_OK;
  ocg_comm* c = ctx->comm;
  OcgDeviceGuard g(ctx->device);
  // second half of the window data: [0, cap) contribution, [cap, cap + ceil(cap / P)) reduced slice
  const long long half = (c->window_bytes / 8) / 2, cap = half / 2;
  if (cap < c->nranks) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_comm_allreduce_f64: window too small (%lld bytes)", c->window_bytes);
  const CommPeers pr = comm_peers(c);
  int grid = ctx->sm_count;  // all CTAs co-resident: they poll the peers inside the kernel
  for (int64_t off = 0; off < n; off += cap) {
    const long long m = n - off < cap ? n - off : cap;
    comm_allreduce_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(pr, buf_dev + off, m, cap, half);
    OCG_CHECK_LAUNCH(ctx, "comm_allreduce_kernel");
  }
  return OCG_OK;
}

extern "C" int ocg_self_gravity_sharded(ocg_ctx* ctx, const double* pos_local_dev, const double* mass_all_dev, int64_t n,
                                        double eps2, double G, double* acc_local_dev, double* pot_local_dev, void* stream) {
  int rc = comm_check(ctx, "ocg_self_gravity_sharded");
  if (rc) return rc;
  if (n < 1 || !pos_local_dev || !mass_all_dev || !acc_local_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_sharded: bad arguments");
  if (!(eps2 > 0.0)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_sharded: eps2 = %g must be > 0", eps2);
  ocg_comm* c = ctx->comm;
  const int P = c->nranks, r = c->rank;
  const long long a = shard_begin(n, P, r), b = shard_begin(n, P, r + 1), n_local = b - a;
  const long long max_local = (n + P - 1) / P;
  if (2 * 3 * max_local * 8 > c->window_bytes / 2)  // the force exchanges own the first half of the window
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_sharded: window of %lld bytes too small for %lld stars over %d ranks (need %lld)",
                    c->window_bytes, (long long)n, P, 2 * 2 * 3 * max_local * 8);
  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_pot = pot_local_dev != nullptr;
  const float e2f = (float)eps2;
  int ex;
  frexpf(sqrtf(e2f), &ex);
  float scale = ldexpf(1.0f, -8 - ex);  // eps at ~2^-8, as ocg_self_gravity
  if (!(scale > 0.f) || !isfinite(scale) || !(e2f * scale * scale > 0.f)) scale = 1.0f;
  // a target-paired stream-K shape (wide or mid rows by the shard's size, ocg_pick_variant); the exchange and the tile pack
  // do not depend on it, so ranks whose block sizes differ by a star may even pick differently
  int variant = ocg_pick_variant(ctx, b - a, n, /*guard=*/false, /*allow_mf=*/false, (n + OCG_TS - 1) / OCG_TS, /*fine_tiles=*/true);
  if (!ocg_variant_is_tp(variant)) variant = ocg_variant_cluster_tp();
  const int CT = ocg_variant_threads(variant) * ocg_variant_tpt(variant);
  const long long grid = ocg_variant_slots(ctx, variant);
  int64_t seg[2] = {0, n};
  OcgClusterRows plan;
  if ((rc = ocg_plan_cluster_rows(ctx, n, seg, 1, a, b, CT, OCG_TS, grid, st, &plan))) return rc;
  float* tiles;
  float4* tgt;
  double* partial;
  unsigned int* tickets;
  rc = ocg_scratch(ctx, OCG_SCR_TILES, (size_t)plan.total_tiles * OCG_TILE_BYTES, (void**)&tiles);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TGT, sizeof(float4) * (size_t)n, (void**)&tgt);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_PARTIAL, sizeof(double) * (size_t)plan.n_slots * 4 * (size_t)n, (void**)&partial);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TICKETS, sizeof(unsigned int) * (size_t)(plan.n_rows > 0 ? plan.n_rows : 1), (void**)&tickets, true);
  if (rc) return rc;
  {
    const long long nslots = plan.total_tiles * OCG_TS;
    long long nb = (nslots + 255) / 256;
    if (nb > 2ll * ctx->sm_count) nb = 2ll * ctx->sm_count;  // all CTAs co-resident: they poll the peers inside the kernel
    comm_gather_pack_kernel<<<(int)nb, 256, 0, st>>>(comm_peers(c), pos_local_dev, mass_all_dev, n, max_local, e2f, scale, tiles, tgt);
    OCG_CHECK_LAUNCH(ctx, "comm_gather_pack_kernel");
  }
  if (plan.n_rows == 0) return OCG_OK;
  DirectParams p{};
  p.out_stride = n, p.n_tgt = n, p.scale_val = scale;
  p.tiles = tiles, p.tgt = tgt, p.partial = partial;
  p.sk.rows = plan.d_rows, p.sk.row_prefix = plan.d_prefix, p.sk.n_rows = plan.n_rows, p.sk.n_slots = plan.n_slots;
  p.sk.n_tgt = n, p.sk.ct = CT, p.sk.nst_uniform = nullptr, p.sk.tickets = tickets;
  // the kernel indexes its outputs with GLOBAL star indices: shift the block arrays so that star a lands on element 0
  p.out_acc = acc_local_dev - a, p.out_pot = want_pot ? pot_local_dev - a : nullptr, p.out_n = n_local;
  p.G = G, p.accumulate = 0, p.self_e2s = want_pot ? e2f * scale * scale : -1.f, p.m0_ptr = nullptr;
  return ocg_launch_direct(ctx, p, variant, want_pot, /*guard=*/false, st);
}

// ---- K6 sharded: the position + velocity gather fused into the Hermite tile pack -------------------------------------
struct HermiteGatherArgs {
  const double *pos_local, *vel_local, *mass_all;
  long long n, max_local;
};
static int hermite_gather_pack(ocg_ctx* ctx, void* user, float scale, long long total_tiles, float* tiles, float4* tgt_pos,
                               float4* tgt_vel, cudaStream_t st) {
  const HermiteGatherArgs* g = (const HermiteGatherArgs*)user;
  long long nb = (total_tiles * HM_TS + 255) / 256;
  if (nb > 2ll * ctx->sm_count) nb = 2ll * ctx->sm_count;  // all CTAs co-resident: they poll the peers inside the kernel
  comm_gather_pack_hermite_kernel<<<(int)nb, 256, 0, st>>>(comm_peers(ctx->comm), g->pos_local, g->vel_local, g->mass_all, g->n,
                                                          g->max_local, scale, total_tiles, tiles, tgt_pos, tgt_vel);
  OCG_CHECK_LAUNCH(ctx, "comm_gather_pack_hermite_kernel");
  return OCG_OK;
}

extern "C" int ocg_self_gravity_hermite_sharded(ocg_ctx* ctx, const double* pos_local_dev, const double* vel_local_dev,
                                                const double* mass_all_dev, int64_t n, double eps2, double G, double vel_to_len,
                                                double* acc_dev, double* jerk_dev, double* pot_dev, void* stream) {
  int rc = comm_check(ctx, "ocg_self_gravity_hermite_sharded");
  if (rc) return rc;
  if (n < 1 || !pos_local_dev || !vel_local_dev || !mass_all_dev || !acc_dev || !jerk_dev)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite_sharded: bad arguments");
  if (!(eps2 >= 0.0)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite_sharded: eps2 = %g must be >= 0", eps2);
  ocg_comm* c = ctx->comm;
  const int P = c->nranks, r = c->rank;
  const long long a = shard_begin(n, P, r), b = shard_begin(n, P, r + 1);
  const long long max_local = (n + P - 1) / P;
  if (18 * max_local * 8 > c->window_bytes / 2)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity_hermite_sharded: window of %lld bytes too small for %lld stars over %d ranks (need %lld)",
                    c->window_bytes, (long long)n, P, 2 * 18 * max_local * 8);
  OcgDeviceGuard g(ctx->device);
  HermiteGatherArgs args = {pos_local_dev, vel_local_dev, mass_all_dev, (long long)n, max_local};
  // every rank must launch the exchange kernel, also one whose block is empty
  return ocg_hermite_force_packed(ctx, n, a, b, eps2, G, vel_to_len, acc_dev, jerk_dev, pot_dev, (cudaStream_t)stream,
                                  hermite_gather_pack, &args);
}

// Internal declarations shared by the translation units of liboc_nbody_b200.
// Not part of the ABI (see include/ocg.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ocg.h"

// ---- geometry of the direct-sum kernel (K1/K4) ------------------------------------------------
// A source tile is TS sources stored component-major: x[TS] | y[TS] | z[TS] | m[TS] | e2[TS].
// One tile = one cp.async.bulk (TMA 1-D bulk copy) of OCG_TILE_BYTES into one pipeline stage.
#define OCG_TS 512
#define OCG_TILE_FLOATS (5 * OCG_TS)
#define OCG_TILE_BYTES (OCG_TILE_FLOATS * 4)
/* mass-folded tiles that also serve the potential carry a 6th array, 1/w */
#define OCG_TILE_ARRAYS(MF, POT) (((MF) && (POT)) ? 6 : 5)
#define OCG_NSTAGE 4
#define OCG_CONSUMER_WARPS 8
#define OCG_CONSUMER_THREADS (OCG_CONSUMER_WARPS * 32)
#define OCG_CTA_THREADS (OCG_CONSUMER_THREADS + 32) /* + 1 TMA producer warp */

// Work item of the direct-sum kernel: a tile of targets against a run of source tiles.
struct OcgWorkItem {
  long long tgt_begin;   // first target (global index into tgt array / output)
  int tgt_count;         // targets in this item (<= CTA targets)
  int tile_count;        // source tiles to stream
  long long tile_begin;  // first source tile
  long long out_slot;    // partial-sum slot: partial[(out_slot*NC + c)*out_stride + tgt]
};

enum { OCG_SCR_TILES = 0, OCG_SCR_PARTIAL, OCG_SCR_NEAR, OCG_SCR_ITEMS, OCG_SCR_MISC, OCG_SCR_TGT,
       OCG_SCR_SRC, OCG_SCR_SOFT, OCG_SCR_F64A, OCG_SCR_F64B, OCG_SCR_F64C, OCG_SCR_OUT, OCG_SCR_COUNTS,
       OCG_SCR_ITEMS_HM, OCG_SCR_TILES_HM, OCG_SCR_TGT_HM, OCG_SCR_PARTIAL_HM, OCG_SCR_TICKETS, OCG_SCR_TICKETS_HM, OCG_SCR_NEARPART, OCG_SCR_BLOCK,
       OCG_SCR_N };

// Tuning / test knobs of one ctx (include/ocg_debug.h: ocg_debug_set).  Defaults are the production behaviour.
struct OcgKnobs {
  int direct_variant;       // -1 = heuristic, else index into the K1/K4 shape table (direct_sum.cu)
  int precise_near;         // 1 = sources inside the precision radius take the FP64 pair path
  int mass_fold;            // 1 = K1 without potential uses mass-folded tiles
  int small_cluster_path;   // 1 = one small cluster takes the fused single-launch K4 kernel
  long long host_chunk;     // particles staged at a time by ocg_field_build_host
  int hermite_variant;      // -1 = production shape of the Hermite force kernel
  int hermite_small_path;   // 1 = one small cluster takes the fused single-launch K6 kernel
  int interp_variant;       // register bound of K3: 0 <=128, 1 <=80, 2 <=64 (production)
  int field_precision;      // 0 = FP32 pair arithmetic + FP64 accumulation (north_star), 1 = every pair in FP64
  long long pass_bytes;     // K1: bytes of source tiles per stream-K pass (0 = one pass); default 32 MB, a quarter of the L2
  long long near_cap;       // > 0: size limit of K1's FP64 precision-radius set (0 = max(n_src / 512, 2^36 / n_src))
};

struct ocg_comm;  // comm.cu: the rank's exchange window and the mapped windows of its peers

struct ocg_ctx {
  int device;
  ocg_comm* comm;
  OcgKnobs knobs;
  int sm_count;
  int sm_clock_khz;
  size_t global_mem;
  char err[1024];
  void* scratch[OCG_SCR_N];
  size_t scratch_bytes[OCG_SCR_N];
  unsigned long long scratch_generation;  // bumped whenever a scratch buffer is (re)allocated: its address changed
  long long launches;
  // pinned staging ring of ocg_set_interp_weight_slots (a pageable source would make the copy synchronise the stream)
  float* w_ring;                    // [OCG_W_RING][16] floats, cudaHostAlloc'd on first use
  cudaEvent_t w_ring_ev[16];        // recorded after the copy out of entry i; waited on before entry i is rewritten
  unsigned w_ring_head;
  int source_shards;  // ocg_set_source_shards: this ctx sees 1/n of the sources of a source-sharded field build
  int timing;
  cudaEvent_t ev0, ev1;
  int ev_valid;
  long long last_traffic_bytes;  // model of the last K1 launch's HBM traffic (ocg_last_direct_traffic_bytes)
  // cached host copy of the last uploaded item list, to skip re-upload: [0] K4, [1] the Hermite force loop.  Each has
  // its own device buffer (OCG_SCR_ITEMS / OCG_SCR_ITEMS_HM): a captured CUDA graph of one must not see the other's plan.
  struct PlanCache {
    OcgWorkItem* items_host;
    size_t items_host_cap;
    size_t items_uploaded;  // number of items currently in device list
    unsigned long long items_hash;
  } plan[2];
};

// ---- error helpers ------------------------------------------------------------------------------
int ocg_fail(ocg_ctx* ctx, int code, const char* fmt, ...);

#define OCG_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ocg_fail((ctx), OCG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)

#define OCG_CHECK_LAUNCH(ctx, name)                                                      \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ocg_fail((ctx), OCG_ERR_CUDA, "launch of %s failed: %s", name,             \
                      cudaGetErrorString(_e));                                           \
    (ctx)->launches++;                                                                   \
  } while (0)

// Grow-only scratch buffer tied to the ctx.
int ocg_scratch(ocg_ctx* ctx, int which, size_t bytes, void** out, bool zero_on_alloc = false);

struct OcgDeviceGuard {
  int prev;
  bool changed;
  explicit OcgDeviceGuard(int dev) : prev(-1), changed(false) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      changed = true;
    }
  }
  ~OcgDeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

// ---- stream-K work decomposition of the target-paired kernels (streamk.cuh) ----------------------
struct OcgRow {
  long long tgt_begin;   // first target of the row (global index)
  long long tile_begin;  // first source tile of the run this row streams
  int tgt_count;         // targets in the row (<= CTA tile)
  int pad;
};

struct StreamKParams {
  const OcgRow* rows;           // list mode (K4/K6); NULL = uniform mode (K1)
  const long long* row_prefix;  // list mode: [n_rows + 1] prefix sums of the rows' source-tile counts
  int n_rows;
  int n_slots;                  // slots per row in the partial buffer (>= the most CTAs any row is shared by)
  long long n_tgt;              // uniform mode: row r = targets [r*ct, min((r+1)*ct, n_tgt))
  int ct;
  const int* nst_uniform;       // uniform mode: device count of source tiles every row streams ...
  int nst_value;                // ... or, when that pointer is NULL, the count itself
  unsigned int* tickets;        // [n_rows]; zero on entry, zero again on exit
  int tile_cap = 0;             // uniform mode: source tiles per pass (0 = one pass), see streamk.cuh
};

// ---- parameter block of the direct-sum kernels (direct_sum.cu) -------------------------------
struct DirectParams {
  const float* tiles;        // packed source tiles
  const float4* tgt;         // targets (x,y,z,-)
  double* partial;           // [slot][NC][out_stride]
  long long out_stride;      // rows of the partial buffer
  const OcgWorkItem* items;  // list mode when non-null (K4), else arithmetic decode (K1)
  int n_items;
  long long n_tgt;           // arithmetic mode: item -> (chunk, target tile)
  int n_ttiles;
  int tiles_per_chunk;
  const int* n_fast_tiles;   // device: number of fast tiles actually present
  const float* scale_ptr;    // device: power-of-two length scale applied to targets (K1), or NULL
  float scale_val;           // host-chosen scale when scale_ptr is NULL (K4)
  // target-paired kernels (direct_sum_tp_kernel): stream-K rows and the fused finish
  StreamKParams sk;
  double* out_acc;           // final field [3][out_n] (and [out_n] potential), written by the row's last CTA:
  double* out_pot;           //   out = G * s^2 * sum (potential G * s), times M0 for mass-folded tiles
  long long out_n;
  double G;
  int accumulate;            // add to the outputs instead of overwriting (K1 source chunks)
  float self_e2s;            // K4: scaled eps^2 whose self term -m/eps is removed from the potential; < 0: none
  const float* m0_ptr;       // mass-folded tiles: device M0 (partials are in units of M0); NULL otherwise
};

// ---- work plan of the cluster kernels (self_gravity.cu; shared by K4 and the Hermite force loop) ----
struct OcgClusterPlan {
  long long total_tiles;  // source tiles over all segments
  long long n_chunks;     // source chunks per target tile = partial-sum slots
  long long n_items;
  const OcgWorkItem* d_items;   // device
  const long long* d_seg_tile;  // device [n_seg+1]: first source tile of each segment
  const long long* d_seg_off;   // device [n_seg+1]: first particle of each segment
};
int ocg_plan_cluster_items(ocg_ctx* ctx, int64_t n, const int64_t* seg_offsets_host, int32_t n_seg, int64_t tgt_begin,
                           int64_t tgt_end, int ct, int ts, long long slots, cudaStream_t st, OcgClusterPlan* out,
                           int which = 0);
// Stream-K form of the plan (streamk.cuh): one row per target tile of the shard, against the source tiles of its segment.
struct OcgClusterRows {
  long long total_tiles;        // source tiles over all segments
  int n_rows;
  int n_slots;                  // most CTAs any row is shared by, for a grid of `grid` CTAs
  const OcgRow* d_rows;         // device [n_rows]
  const long long* d_prefix;    // device [n_rows + 1]
  const long long* d_seg_tile;  // device [n_seg + 1]
  const long long* d_seg_off;   // device [n_seg + 1]
};
// ---- K6 (hermite.cu): tile format shared with the exchange layer (comm.cu) ----
#define HM_TS 512
#define HM_NARR 7 /* x | y | z | m | vx | vy | vz */
#define HM_TILE_FLOATS (HM_NARR * HM_TS)
#define HM_TILE_BYTES (HM_TILE_FLOATS * 4)
// Fills the source tiles ([tile][7][HM_TS], positions times `scale`) and the float4 targets of ONE cluster of n stars,
// recentred on star 0: pack_hermite_kernel for resident arrays, the peer-memory gather of comm.cu for a sharded cluster.
typedef int (*ocg_hermite_pack_fn)(ocg_ctx* ctx, void* user, float scale, long long total_tiles, float* tiles,
                                   float4* tgt_pos, float4* tgt_vel, cudaStream_t st);
// K6 for the targets [tgt_begin, tgt_end) of one cluster of n stars whose tiles `pack` lays out; outputs [3][n] / [n].
int ocg_hermite_force_packed(ocg_ctx* ctx, int64_t n, int64_t tgt_begin, int64_t tgt_end, double eps2, double G,
                             double vel_to_len, double* acc_dev, double* jerk_dev, double* pot_dev, cudaStream_t st,
                             ocg_hermite_pack_fn pack, void* user);

int ocg_plan_cluster_rows(ocg_ctx* ctx, int64_t n, const int64_t* seg_offsets_host, int32_t n_seg, int64_t tgt_begin,
                          int64_t tgt_end, int ct, int ts, long long grid, cudaStream_t st, OcgClusterRows* out, int which = 0);

// ---- entry points implemented in other translation units ------------------------------------
int ocg_pick_variant(ocg_ctx* ctx, int64_t n_tgt, int64_t seg_len, bool guard, bool allow_mf = false,
                     int64_t src_tiles = 0, bool fine_tiles = false);
int ocg_direct_n_variants();
const char* ocg_direct_variant_name(int id);
bool ocg_direct_variant_built(int id);
int ocg_hermite_n_variants();
const char* ocg_hermite_variant_name(int id);
bool ocg_variant_is_tp(int variant);
int ocg_variant_cluster_tp();  // the target-paired shape K4 uses for clusters (512-target rows)
int ocg_variant_tpt(int variant);
int ocg_variant_threads(int variant);
int ocg_variant_slots(ocg_ctx* ctx, int variant);
int ocg_launch_direct(ocg_ctx* ctx, DirectParams& p, int variant, bool pot, bool guard, cudaStream_t st);
int ocg_direct_sum_impl(ocg_ctx* ctx, const float* src_xyzm, const float* src_soft, int64_t n_src,
                        const float* tgt_xyzw, int64_t n_tgt, int kernel, double G, double* acc,
                        double* pot, int accumulate, cudaStream_t st);

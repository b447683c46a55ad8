// K4: cluster self-gravity — Plummer-softened direct sum inside each cluster of a batch.
// Replaces the force loop of the AMUSE ph4 worker the reference instantiates at
// oc_code.py:218-229 (epsilon_squared = (softening pc)^2, oc_code.py:225).
//
// Same streaming kernel as the field build (direct_sum.cu); here sources == targets, the softening
// is one scalar, segments (clusters) are independent, and each segment is recentred on its first
// particle in FP64 before rounding to FP32 (cluster stars sit ~8 kpc from the origin of the
// galaxy frame but ~pc from each other).
#include "ocg_internal.cuh"

#include <math.h>
#include <stdlib.h>

// Re-lay FP64 particles into (a) padded per-segment source tiles and (b) float4 targets.
// seg_off  : device int64 [n_seg+1] particle offsets; seg_tile : device int64 [n_seg+1] tile offsets
__global__ void pack_cluster_kernel(const double* __restrict__ pos, const double* __restrict__ mass,
                                    long long n, const long long* __restrict__ seg_off,
                                    const long long* __restrict__ seg_tile, int n_seg, float e2,
                                    float scale, float* __restrict__ tiles, float4* __restrict__ tgt) {
  const long long total_tiles = seg_tile[n_seg];
  const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (slot >= total_tiles * OCG_TS) return;
  const long long tile = slot / OCG_TS;
  const int j = (int)(slot - tile * OCG_TS);
  // segment owning this tile: last s with seg_tile[s] <= tile
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_tile[mid] <= tile) lo = mid;
    else hi = mid;
  }
  const long long first = seg_off[lo];
  const long long idx = first + (tile - seg_tile[lo]) * OCG_TS + j;
  float* T = tiles + tile * (long long)OCG_TILE_FLOATS;
  if (idx < seg_off[lo + 1]) {
    const double cx = pos[first], cy = pos[n + first], cz = pos[2 * n + first];
    const float x = (float)(pos[idx] - cx), y = (float)(pos[n + idx] - cy), z = (float)(pos[2 * n + idx] - cz);
    const float m = (float)mass[idx];
    // tiles hold scaled coordinates (scale = 2^k, exact); the kernel scales the targets itself
    T[j] = x * scale, T[OCG_TS + j] = y * scale, T[2 * OCG_TS + j] = z * scale, T[3 * OCG_TS + j] = m;
    T[4 * OCG_TS + j] = e2 * scale * scale;
    tgt[idx] = make_float4(x, y, z, m);
  } else {
    T[j] = 0.f, T[OCG_TS + j] = 0.f, T[2 * OCG_TS + j] = 0.f, T[3 * OCG_TS + j] = 0.f, T[4 * OCG_TS + j] = 1.f;
  }
}

// out[c][t] = G * s^2 * sum_slots partial (potential: G * s; partials are in scaled units).  The
// potential gets the self term (-m/eps, included by the kernel because targets == sources) removed
// with the same FP32 expression the kernel used for it.
__global__ void finish_self_kernel(const double* __restrict__ partial, long long stride, int n_slots, int nc,
                                   double G, long long t0, long long t1, const float4* __restrict__ tgt, float e2s,
                                   float scale, double* __restrict__ acc, double* __restrict__ pot) {
  long long t = t0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= t1) return;
  const double sc = (double)scale;
  for (int c = 0; c < nc; ++c) {
    double s = 0.0;
    for (int k = 0; k < n_slots; ++k) s += partial[((long long)k * nc + c) * stride + t];
    if (c == 3 && e2s > 0.f) {
      float y3;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y3) : "f"((e2s * e2s) * e2s));
      s += (double)((tgt[t].w * y3) * e2s);
    }
    s *= c < 3 ? G * sc * sc : G * sc;
    if (c < 3) acc[(long long)c * stride + t] = s;
    else pot[t] = s;
  }
}

// ------------------------------------------------------------------------ small clusters ----
// One launch for a cluster of a few thousand stars (the reference's own configuration is N = 1 024, test_options:57),
// where the streaming kernel above is latency-bound (pack + TMA ring start-up + finish = 3 launches for ~1e6 pairs).
// A CTA stages ALL sources once in shared memory (recentred on the first particle in FP64, rounded to FP32 — the same
// inputs the streaming path and the oracle use), each of its 8 warps owns one target, the 32 lanes split the sources,
// FP32 pair arithmetic (guarded rsqrt(r^2)^3 form: valid for any eps2 >= 0), four independent FP32 partial sums per
// lane, FP64 from the warp-shuffle reduction on.  The potential excludes the self term by index.
#define SG_SMALL_WARPS 8
#define SG_SMALL_MAX_N 4096
__global__ void __launch_bounds__(32 * SG_SMALL_WARPS) self_gravity_small_kernel(
    const double* __restrict__ pos, const double* __restrict__ mass, long long n, float e2, double G,
    long long tgt_begin, long long tgt_end, double* __restrict__ acc, double* __restrict__ pot) {
  extern __shared__ float4 s_src[];
  const double cx = pos[0], cy = pos[n], cz = pos[2 * n];
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    s_src[i] = make_float4((float)(pos[i] - cx), (float)(pos[n + i] - cy), (float)(pos[2 * n + i] - cz), (float)mass[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t = tgt_begin + (long long)blockIdx.x * SG_SMALL_WARPS + warp;
  if (t >= tgt_end) return;
  const float4 T = s_src[t];
  float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f}, ap[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long j0 = lane; j0 < n; j0 += 128) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = j0 + 32 * u;
      if (j < n) {
        const float4 S = s_src[j];
        const float dx = S.x - T.x, dy = S.y - T.y, dz = S.z - T.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, e2)));
        float ri;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"(r2));
        if (!(r2 > 0.f)) ri = 0.f;
        const float mri = S.w * ri;
        const float sc = mri * (ri * ri);
        ax[u] = fmaf(dx, sc, ax[u]), ay[u] = fmaf(dy, sc, ay[u]), az[u] = fmaf(dz, sc, az[u]);
        if (j != t) ap[u] += mri;
      }
    }
  }
  double v[4] = {((double)ax[0] + (double)ax[1]) + ((double)ax[2] + (double)ax[3]),
                 ((double)ay[0] + (double)ay[1]) + ((double)ay[2] + (double)ay[3]),
                 ((double)az[0] + (double)az[1]) + ((double)az[2] + (double)az[3]),
                 ((double)ap[0] + (double)ap[1]) + ((double)ap[2] + (double)ap[3])};
#pragma unroll
  for (int c = 0; c < 4; ++c)
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
  if (lane == 0) {
    acc[t] = G * v[0], acc[n + t] = G * v[1], acc[2 * n + t] = G * v[2];
    if (pot) pot[t] = -G * v[3];
  }
}

static unsigned long long fnv1a(const void* data, size_t bytes, unsigned long long h) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < bytes; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

// Host-side plan shared by K4 and the Hermite force loop (hermite.cu): source-tile offsets per segment, the
// number of source chunks, and the (target tile x source chunk) item list, uploaded to the device (skipped when
// the same plan is already there).  `ct` targets per CTA tile, `ts` sources per source tile, `slots` resident CTAs.
int ocg_plan_cluster_items(ocg_ctx* ctx, int64_t n, const int64_t* seg_offsets_host, int32_t n_seg, int64_t tgt_begin,
                           int64_t tgt_end, int ct, int ts, long long slots, cudaStream_t st, OcgClusterPlan* out, int which) {
  const int CT = ct;
  ocg_ctx::PlanCache& pc = ctx->plan[which ? 1 : 0];
  const int scr = which ? OCG_SCR_ITEMS_HM : OCG_SCR_ITEMS;
  long long* seg_tile = (long long*)malloc(sizeof(long long) * (2 * (size_t)n_seg + 2));
  if (!seg_tile) return ocg_fail(ctx, OCG_ERR_NOMEM, "malloc failed");
  long long* seg_off_ll = seg_tile + n_seg + 1;
  long long total_tiles = 0, n_tt_total = 0, max_tiles = 1;
  for (int s = 0; s <= n_seg; ++s) seg_off_ll[s] = seg_offsets_host[s];
  for (int s = 0; s < n_seg; ++s) {
    seg_tile[s] = total_tiles;
    long long len = seg_off_ll[s + 1] - seg_off_ll[s];
    long long nt = (len + ts - 1) / ts;
    total_tiles += nt;
    if (nt > max_tiles) max_tiles = nt;
    long long a = seg_off_ll[s] > tgt_begin ? seg_off_ll[s] : tgt_begin;
    long long b = seg_off_ll[s + 1] < tgt_end ? seg_off_ll[s + 1] : tgt_end;
    if (b > a) n_tt_total += (b - a + CT - 1) / CT;
  }
  seg_tile[n_seg] = total_tiles;
  // Source chunking: an item streams `tpc` tiles of its segment; items are strided statically over the resident
  // CTAs, so the makespan is rounds * tpc tile-times (+ ~1/16 tile of start-up per item).  Pick the tiles-per-chunk
  // that minimises it for the largest segment (every chunk non-empty there), keeping <= 256 partial-sum slots.
  long long n_chunks = 1;
  if (n_tt_total > 0) {
    double best = 1e300;
    for (long long tpc = 1; tpc <= max_tiles; ++tpc) {
      const long long ch = (max_tiles + tpc - 1) / tpc;
      if (ch > 256) continue;
      const long long rounds = (n_tt_total * ch + slots - 1) / slots;
      const double cost = (double)rounds * ((double)tpc + 0.0625);
      if (cost < best) best = cost, n_chunks = ch;
    }
  }
  const long long n_items = n_tt_total * n_chunks;
  if (n_items > 0x7fffffffll) {
    free(seg_tile);
    return ocg_fail(ctx, OCG_ERR_INVALID, "too many work items");
  }
  if ((size_t)n_items > pc.items_host_cap) {
    free(pc.items_host);
    pc.items_host = (OcgWorkItem*)malloc(sizeof(OcgWorkItem) * (size_t)n_items);
    pc.items_host_cap = pc.items_host ? (size_t)n_items : 0;
    pc.items_uploaded = 0;
    if (!pc.items_host) {
      free(seg_tile);
      return ocg_fail(ctx, OCG_ERR_NOMEM, "malloc of %lld work items failed", n_items);
    }
  }
  // chunk-major order: CTAs resident together stream the same source tiles (L2 reuse)
  long long w = 0;
  for (long long c = 0; c < n_chunks; ++c) {
    for (int s = 0; s < n_seg; ++s) {
      long long a = seg_off_ll[s] > tgt_begin ? seg_off_ll[s] : tgt_begin;
      long long b = seg_off_ll[s + 1] < tgt_end ? seg_off_ll[s + 1] : tgt_end;
      if (b <= a) continue;
      long long nt = seg_tile[s + 1] - seg_tile[s];
      long long tpc = (nt + n_chunks - 1) / n_chunks;
      long long tb = c * tpc, te = tb + tpc < nt ? tb + tpc : nt;
      for (long long t0 = a; t0 < b; t0 += CT) {
        OcgWorkItem& it = pc.items_host[w++];
        it.tgt_begin = t0;
        it.tgt_count = (int)(b - t0 < CT ? b - t0 : CT);
        it.tile_begin = seg_tile[s] + (tb < nt ? tb : nt);
        it.tile_count = (int)(te > tb ? te - tb : 0);
        it.out_slot = c;
      }
    }
  }
  unsigned long long h = fnv1a(pc.items_host, sizeof(OcgWorkItem) * (size_t)n_items, 1469598103934665603ull);
  h = fnv1a(seg_tile, sizeof(long long) * (2 * (size_t)n_seg + 2), h);

  int rc = 0;
  OcgWorkItem* d_items;
  long long* d_seg;
  const size_t seg_bytes = sizeof(long long) * (2 * (size_t)n_seg + 2);
  const size_t items_bytes = sizeof(OcgWorkItem) * (size_t)n_items;
  void* items_raw = nullptr;
  const size_t before = ctx->scratch_bytes[scr];
  if (!rc) rc = ocg_scratch(ctx, scr, items_bytes + seg_bytes + 64, &items_raw);
  if (rc) {
    free(seg_tile);
    return rc;
  }
  d_items = (OcgWorkItem*)items_raw;
  d_seg = (long long*)((char*)items_raw + ((items_bytes + 63) / 64) * 64);
  const bool reuse = before == ctx->scratch_bytes[scr] && pc.items_uploaded == (size_t)n_items &&
                     pc.items_hash == h;
  if (!reuse) {
    // synchronous small copies: the plan changes only when the segment layout or shard changes
    cudaError_t e = cudaMemcpyAsync(d_items, pc.items_host, items_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seg, seg_tile, seg_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // seg_tile is freed below
    if (e != cudaSuccess) {
      free(seg_tile);
      return ocg_fail(ctx, OCG_ERR_CUDA, "upload of the work plan failed: %s", cudaGetErrorString(e));
    }
    pc.items_uploaded = (size_t)n_items;
    pc.items_hash = h;
  }
  free(seg_tile);
  out->total_tiles = total_tiles, out->n_chunks = n_chunks, out->n_items = n_items;
  out->d_items = d_items, out->d_seg_tile = d_seg, out->d_seg_off = d_seg + n_seg + 1;
  return OCG_OK;
}


// Stream-K plan (streamk.cuh) shared by K4 and the Hermite force loop: rows = target tiles of the shard, each against the
// source tiles of its own segment; uploaded once and kept while the segment layout / shard / tile shape stay the same.
// The cache slot and device buffer are those of the item plan (`which`: 0 K4, 1 K6), so the capture epoch covers it.
int ocg_plan_cluster_rows(ocg_ctx* ctx, int64_t n, const int64_t* seg_offsets_host, int32_t n_seg, int64_t tgt_begin,
                          int64_t tgt_end, int ct, int ts, long long grid, cudaStream_t st, OcgClusterRows* out, int which) {
  ocg_ctx::PlanCache& pc = ctx->plan[which ? 1 : 0];
  const int scr = which ? OCG_SCR_ITEMS_HM : OCG_SCR_ITEMS;
  long long n_rows = 0, total_tiles = 0;
  for (int s = 0; s < n_seg; ++s) {
    const long long a = seg_offsets_host[s] > tgt_begin ? seg_offsets_host[s] : tgt_begin;
    const long long b = seg_offsets_host[s + 1] < tgt_end ? seg_offsets_host[s + 1] : tgt_end;
    if (b > a) n_rows += (b - a + ct - 1) / ct;
    total_tiles += (seg_offsets_host[s + 1] - seg_offsets_host[s] + ts - 1) / ts;
  }
  if (n_rows > 0x7fffffffll) return ocg_fail(ctx, OCG_ERR_INVALID, "too many target tiles");
  // one host block: rows | prefix | seg_tile | seg_off   (OcgRow is 24 bytes: keeps every part 8-byte aligned)
  const size_t rows_bytes = sizeof(OcgRow) * (size_t)n_rows, pre_bytes = sizeof(long long) * ((size_t)n_rows + 1);
  const size_t seg_bytes = sizeof(long long) * (2 * (size_t)n_seg + 2), bytes = rows_bytes + pre_bytes + seg_bytes;
  const size_t cap_items = (bytes + sizeof(OcgWorkItem) - 1) / sizeof(OcgWorkItem);
  if (cap_items > pc.items_host_cap) {
    free(pc.items_host);
    pc.items_host = (OcgWorkItem*)malloc(sizeof(OcgWorkItem) * cap_items);
    pc.items_host_cap = pc.items_host ? cap_items : 0;
    pc.items_uploaded = 0;
    if (!pc.items_host) return ocg_fail(ctx, OCG_ERR_NOMEM, "malloc of the work plan failed");
  }
  char* base = (char*)pc.items_host;
  OcgRow* rows = (OcgRow*)base;
  long long* prefix = (long long*)(base + rows_bytes);
  long long* seg_tile = (long long*)(base + rows_bytes + pre_bytes);
  long long* seg_off = seg_tile + n_seg + 1;
  long long tile0 = 0, r = 0, units = 0;
  for (int s = 0; s < n_seg; ++s) {
    const long long len = seg_offsets_host[s + 1] - seg_offsets_host[s], nt = (len + ts - 1) / ts;
    seg_tile[s] = tile0, seg_off[s] = seg_offsets_host[s];
    const long long a = seg_offsets_host[s] > tgt_begin ? seg_offsets_host[s] : tgt_begin;
    const long long b = seg_offsets_host[s + 1] < tgt_end ? seg_offsets_host[s + 1] : tgt_end;
    for (long long t0 = a; t0 < b; t0 += ct, ++r) {
      rows[r].tgt_begin = t0, rows[r].tile_begin = tile0;
      rows[r].tgt_count = (int)(b - t0 < ct ? b - t0 : ct), rows[r].pad = 0;
      prefix[r] = units;
      units += nt;
    }
    tile0 += nt;
  }
  prefix[n_rows] = units;
  seg_tile[n_seg] = tile0, seg_off[n_seg] = seg_offsets_host[n_seg];
  // the most CTAs a row is shared by, with exactly the kernel's arithmetic (streamk.cuh: sk_cta_of)
  long long n_slots = 1;
  if (units < grid) grid = units;  // sk_ctas
  for (long long i = 0; i < n_rows && units > 0; ++i) {
    const long long c0 = ((prefix[i] + 1) * grid + units - 1) / units - 1, c1 = (prefix[i + 1] * grid + units - 1) / units - 1;
    if (c1 - c0 + 1 > n_slots) n_slots = c1 - c0 + 1;
  }
  unsigned long long h = fnv1a(base, bytes, 1469598103934665603ull ^ (unsigned long long)grid);
  void* raw = nullptr;
  const size_t before = ctx->scratch_bytes[scr];
  int rc = ocg_scratch(ctx, scr, bytes + 64, &raw);
  if (rc) return rc;
  const bool reuse = before == ctx->scratch_bytes[scr] && pc.items_uploaded == bytes && pc.items_hash == h;
  if (!reuse) {
    // synchronous small copy: the plan changes only when the segment layout, the shard or the tile shape change
    cudaError_t e = cudaMemcpyAsync(raw, base, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return ocg_fail(ctx, OCG_ERR_CUDA, "upload of the work plan failed: %s", cudaGetErrorString(e));
    pc.items_uploaded = bytes;
    pc.items_hash = h;
  }
  out->total_tiles = tile0, out->n_rows = (int)n_rows, out->n_slots = (int)n_slots;
  out->d_rows = (const OcgRow*)raw;
  out->d_prefix = (const long long*)((char*)raw + rows_bytes);
  out->d_seg_tile = (const long long*)((char*)raw + rows_bytes + pre_bytes);
  out->d_seg_off = out->d_seg_tile + n_seg + 1;
  return OCG_OK;
}

extern "C" int ocg_self_gravity(ocg_ctx* ctx, const double* pos_dev, const double* mass_dev, int64_t n,
                                const int64_t* seg_offsets_host, int32_t n_seg, double eps2, double G,
                                int64_t tgt_begin, int64_t tgt_end, double* acc_dev, double* pot_dev,
                                void* stream) {
  if (!ctx) return OCG_ERR_INVALID;
  if (n < 0 || n_seg < 1) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: n < 0 or n_seg < 1");
  if (n == 0) return OCG_OK;
  if (!pos_dev || !mass_dev || !acc_dev) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: NULL argument");
  if (!(eps2 >= 0.0)) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: eps2 = %g must be >= 0", eps2);
  if (tgt_begin < 0 || tgt_end > n || tgt_begin > tgt_end)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: target range [%lld,%lld) outside [0,%lld)",
                    (long long)tgt_begin, (long long)tgt_end, (long long)n);
  int64_t one_seg[2] = {0, n};
  if (!seg_offsets_host) {
    if (n_seg != 1) return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: seg_offsets is NULL but n_seg = %d", n_seg);
    seg_offsets_host = one_seg;
  }
  if (seg_offsets_host[0] != 0 || seg_offsets_host[n_seg] != n)
    return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: seg_offsets must run from 0 to n");
  for (int s = 0; s < n_seg; ++s)
    if (seg_offsets_host[s + 1] < seg_offsets_host[s])
      return ocg_fail(ctx, OCG_ERR_INVALID, "ocg_self_gravity: seg_offsets not monotone at %d", s);
  if (tgt_begin == tgt_end) return OCG_OK;

  OcgDeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_pot = pot_dev != nullptr;
  const int NC = want_pot ? 4 : 3;
  const float e2f = (float)eps2;
  if (ctx->knobs.small_cluster_path && n_seg == 1 && n <= SG_SMALL_MAX_N) {
    // a single small cluster: one fused launch (see self_gravity_small_kernel)
    const size_t smem = sizeof(float4) * (size_t)n;
    OCG_CUDA(ctx, cudaFuncSetAttribute((const void*)self_gravity_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(float4) * SG_SMALL_MAX_N)));
    const long long nb = (tgt_end - tgt_begin + SG_SMALL_WARPS - 1) / SG_SMALL_WARPS;
    self_gravity_small_kernel<<<(int)nb, 32 * SG_SMALL_WARPS, smem, st>>>(pos_dev, mass_dev, n, e2f, G, tgt_begin, tgt_end,
                                                                         acc_dev, pot_dev);
    OCG_CHECK_LAUNCH(ctx, "self_gravity_small_kernel");
    ctx->ev_valid = 0;
    return OCG_OK;
  }
  const bool guard = !(e2f > 0.f);
  // power-of-two length scale that puts eps at ~2^-8, so that r^6 >= eps^6 ~ 3.5e-15 stays far from the
  // FP32 underflow threshold and separations up to ~1e8 eps stay below overflow (the guarded eps2 == 0
  // form uses rsqrt(r2)^3 and needs no scale)
  float scale = 1.0f;
  if (!guard) {
    int e;
    frexpf(sqrtf(e2f), &e);
    scale = ldexpf(1.0f, -8 - e);
    if (!(scale > 0.f) || !isfinite(scale) || !(e2f * scale * scale > 0.f)) scale = 1.0f;
  }
  const int64_t n_shard = tgt_end - tgt_begin;
  // the tile shape follows the typical cluster size, not the batch size: a 4096-star cluster would fill only
  // 1 1/3 of the 3072-target tiles of the big-grid kernel
  const int64_t seg_typ = (n + n_seg - 1) / n_seg;
  const int variant = ocg_pick_variant(ctx, n_shard, seg_typ, guard, /*allow_mf=*/false, (seg_typ + OCG_TS - 1) / OCG_TS,
                                       /*fine_tiles=*/true);
  const int tpt = ocg_variant_tpt(variant);
  const int CT = ocg_variant_threads(variant) * tpt;

  int rc;
  float* tiles;
  float4* tgt;
  double* partial;
  DirectParams p{};
  p.out_stride = n;
  p.n_tgt = n;
  p.scale_val = scale;
  const float e2s = guard ? 0.f : e2f * scale * scale;
  if (ocg_variant_is_tp(variant)) {
    // ---- target-paired kernel: stream-K rows, the row's last CTA writes the field (two launches: pack + kernel) ----
    const long long grid = ocg_variant_slots(ctx, variant);
    OcgClusterRows plan;
    if ((rc = ocg_plan_cluster_rows(ctx, n, seg_offsets_host, n_seg, tgt_begin, tgt_end, CT, OCG_TS, grid, st, &plan))) return rc;
    unsigned int* tickets;
    rc = ocg_scratch(ctx, OCG_SCR_TILES, (size_t)plan.total_tiles * OCG_TILE_BYTES, (void**)&tiles);
    if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TGT, sizeof(float4) * (size_t)n, (void**)&tgt);
    // always sized for the 4-component form and the largest slot count: a later call that also wants the potential
    // (bound_center_of_mass after a captured step) must not move a buffer a CUDA graph has frozen
    if (!rc) rc = ocg_scratch(ctx, OCG_SCR_PARTIAL, sizeof(double) * (size_t)plan.n_slots * 4 * (size_t)n, (void**)&partial);
    if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TICKETS, sizeof(unsigned int) * (size_t)(plan.n_rows > 0 ? plan.n_rows : 1), (void**)&tickets, true);
    if (rc) return rc;
    const long long nslots = plan.total_tiles * OCG_TS;
    pack_cluster_kernel<<<(int)((nslots + 255) / 256), 256, 0, st>>>(pos_dev, mass_dev, n, plan.d_seg_off, plan.d_seg_tile, n_seg,
                                                                    guard ? 0.f : e2f, scale, tiles, tgt);
    OCG_CHECK_LAUNCH(ctx, "pack_cluster_kernel");
    p.tiles = tiles, p.tgt = tgt, p.partial = partial;
    p.sk.rows = plan.d_rows, p.sk.row_prefix = plan.d_prefix, p.sk.n_rows = plan.n_rows, p.sk.n_slots = plan.n_slots;
    p.sk.n_tgt = n, p.sk.ct = CT, p.sk.nst_uniform = nullptr, p.sk.tickets = tickets;
    p.out_acc = acc_dev, p.out_pot = pot_dev, p.out_n = n;
    p.G = G, p.accumulate = 0, p.self_e2s = want_pot ? e2s : -1.f, p.m0_ptr = nullptr;
    if (plan.n_rows == 0) return OCG_OK;
    return ocg_launch_direct(ctx, p, variant, want_pot, guard, st);
  }
  // ---- source-paired kernels (few targets; the guarded eps2 == 0 form): (target tile x source chunk) items + finish ----
  OcgClusterPlan plan;
  if ((rc = ocg_plan_cluster_items(ctx, n, seg_offsets_host, n_seg, tgt_begin, tgt_end, CT, OCG_TS,
                                   ocg_variant_slots(ctx, variant), st, &plan)))
    return rc;
  const long long total_tiles = plan.total_tiles, n_chunks = plan.n_chunks, n_items = plan.n_items;
  rc = ocg_scratch(ctx, OCG_SCR_TILES, (size_t)total_tiles * OCG_TILE_BYTES, (void**)&tiles);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_TGT, sizeof(float4) * (size_t)n, (void**)&tgt);
  if (!rc) rc = ocg_scratch(ctx, OCG_SCR_PARTIAL, sizeof(double) * (size_t)n_chunks * 4 * (size_t)n, (void**)&partial);
  if (rc) return rc;
  {
    long long nslots = total_tiles * OCG_TS;
    pack_cluster_kernel<<<(int)((nslots + 255) / 256), 256, 0, st>>>(pos_dev, mass_dev, n, plan.d_seg_off, plan.d_seg_tile,
                                                                    n_seg, guard ? 0.f : e2f, scale, tiles, tgt);
    OCG_CHECK_LAUNCH(ctx, "pack_cluster_kernel");
  }
  p.tiles = tiles, p.tgt = tgt, p.partial = partial;
  p.items = plan.d_items;
  p.n_items = (int)n_items;
  if ((rc = ocg_launch_direct(ctx, p, variant, want_pot, guard, st))) return rc;
  finish_self_kernel<<<(int)((n_shard + 255) / 256), 256, 0, st>>>(partial, n, (int)n_chunks, NC, G, tgt_begin, tgt_end, tgt, e2s,
                                                                  scale, acc_dev, pot_dev);
  OCG_CHECK_LAUNCH(ctx, "finish_self_kernel");
  return OCG_OK;
}

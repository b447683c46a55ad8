/*
 * ocg.h — C ABI of liboc_nbody_b200 ("oceanic gravity"), the B200 (sm_100a) drop-in for
 * the gravity hot path of gusbeane/oc_nbody ("oceanic").
 *
 * Every entry point is plain C: pointers, sizes, scalars; int status return
 * (0 = OCG_OK, <0 = error, text via ocg_last_error()).  Nothing throws, nothing exits
 * (contrast the reference's sys.exit(0) at gizmo_interface.py:461).
 *
 * Pointer convention
 *   *_dev  : device pointer owned by the caller (e.g. torch tensor .data_ptr()); written in place.
 *   *_host : host pointer owned by the caller (e.g. numpy array .ctypes.data).
 * The library only allocates scratch tied to the ocg_ctx; ocg_destroy() frees it.
 * One ctx per GPU per process; a ctx is not thread-safe (the reference's only caller is the
 * single-threaded AMUSE Bridge, oc_nbody.py:49 `use_threading=False`).
 * All device-pointer calls are asynchronous on `stream` (a cudaStream_t passed as void*;
 * NULL = the legacy default stream).  Host-pointer calls return after the result is on the host.
 *
 * The library is unit-agnostic: lengths and masses in, G as an argument.
 *   field path  : kpc, Msun and the reference's G in kpc^2 km s^-1 Myr^-1 Msun^-1 (gizmo_interface.py:70)
 *                 = 4.3986004e-09, giving accelerations in km/s/Myr (gizmo_interface.py:706-708);
 *   cluster path: the same system (kpc, km/s, Myr, Msun) so the BRIDGE kick needs no conversion,
 *                 or e.g. pc, Msun, G = 4.30091727e-03 pc (km/s)^2 / Msun.
 *
 * Reference interface replaced by each entry point is cited as (file:line) into gusbeane/oc_nbody.
 */
#ifndef OCG_H
#define OCG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The library is built with -fvisibility=hidden; only the entry points declared here (and in ocg_debug.h) are exported. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define OCG_VERSION 200 /* 0.2.0 */

/* status codes */
#define OCG_OK 0
#define OCG_ERR_INVALID (-1) /* bad argument */
#define OCG_ERR_CUDA (-2)    /* CUDA runtime error; see ocg_last_error */
#define OCG_ERR_NOMEM (-3)   /* scratch allocation failed */
#define OCG_ERR_NODEVICE (-4)

/* softening kernels (gizmo_interface.py:530-550,561 pass a per-source length to pykdgrav) */
#define OCG_KERNEL_PLUMMER 0 /* K = (r^2 + s^2)^(-3/2); s = soft                           */
#define OCG_KERNEL_SPLINE 1  /* cubic-spline, compact support h = soft (pykdgrav ForceKernel) */

typedef struct ocg_ctx ocg_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int ocg_version(void);
int ocg_create(int device, ocg_ctx** out);
int ocg_destroy(ocg_ctx* ctx);
/* Last error text of this ctx ("" if none). ctx may be NULL: returns the text of the last failed ocg_create. */
const char* ocg_last_error(const ocg_ctx* ctx);
/* Device facts used by the roofline report. */
int ocg_device_info(ocg_ctx* ctx, int* sm_count, int* sm_clock_khz, int64_t* global_mem_bytes);
/* Counts kernels this library launched on this ctx since creation (bench.py "gpu_launches"). */
int64_t ocg_launch_count(const ocg_ctx* ctx);
/* CUDA-graph safety.  A captured graph of calls on this ctx freezes the addresses of the ctx's scratch buffers and
 * relies on the work plan (K4 / K6 item list) resident in them.  The value returned here changes whenever either does
 * (a later call needed a larger buffer; another particle set uploaded its plan): a caller that replays a captured
 * graph compares it with the value taken right after capture and re-captures on a mismatch (bridge.Bridge does).
 * The simplest way never to see a mismatch is one ctx per captured step: a ctx is only scratch + settings.        */
int64_t ocg_capture_epoch(const ocg_ctx* ctx);
/* Milliseconds (cudaEvent, on `stream`) spent inside the dominant kernel of the most recent
 * ocg_field_direct / ocg_self_gravity call; the call synchronises `stream`. <0 on error. */
double ocg_last_direct_kernel_ms(ocg_ctx* ctx);
/* Enable(1)/disable(0) event timing around the direct-sum kernel (default off). */
int ocg_set_kernel_timing(ocg_ctx* ctx, int enabled);
/* Source-sharded field build (every rank sums its own shard of the snapshot particles over all targets, the
 * partial fields are then added across ranks): tell this ctx that its ocg_field_direct / ocg_field_build_host calls see
 * 1/n_shards of the sources.  The size limit of the FP64 near set (the sources inside the precision radius) is a property
 * of the whole build; each rank takes 1/n_shards of it.  Default 1. */
int ocg_set_source_shards(ocg_ctx* ctx, int32_t n_shards);
/* Bytes the streaming kernel of the most recent ocg_field_direct call has to move through HBM (source tiles and
 * targets read once, FP64 chunk partials written once): the model bench.py prints beside the ncu-measured traffic. */
int64_t ocg_last_direct_traffic_bytes(const ocg_ctx* ctx);

/* ---- K0: recentre fp64 -> fp32 ------------------------------------------------
 * out_xyzw[i] = (float)(pos[i] - center) , w = (float)mass[i] (or 0 when mass_dev == NULL).
 * pos_dev is [n][3] fp64 row-major, exactly the `np.float64(r)` array the reference hands to
 * pykdgrav (gizmo_interface.py:561) or `grid.evolved_grid` (gizmo_interface.py:564).            */
int ocg_recentre_f64(ocg_ctx* ctx, const double* pos_dev, const double* mass_dev, int64_t n,
                     const double center[3], float* out_xyzw_dev, void* stream);
/* out[i] = (float)in[i] */
int ocg_cast_f64_f32(ocg_ctx* ctx, const double* in_dev, int64_t n, float* out_dev, void* stream);

/* ---- K0b: source assembly on the device (gizmo_interface.py:297-304, 515-558) ----
 * One species per call.  From the species' raw snapshot arrays — pos_dev fp64 [n][3] ('host.distance.principal'),
 * mass_dev fp64 [n], id_dev int64 [n] or NULL, hsml_dev fp64 [n] ('smooth.length', pc; gas only) — keep the particles with
 * id != exclude_id (the tracked star, :515) and |pos| < rmax (_clean_Rmag_, :297-304; rmax <= 0: no cut), give each its
 * softening length in kpc by the reference's rule times soft_scale (1, or the Plummer-equivalent factor), recentre on
 * `center` in fp64, round to fp32, and write the records at out_xyzm_dev[out_offset + j], out_soft_dev[out_offset + j] in
 * the particles' own order: three calls with running offsets (star, dark, gas) produce the concatenation of :518-549.
 * n_kept_host (HOST) receives the number kept; the call synchronises `stream`.                                      */
#define OCG_SOFT_CONSTANT 0        /* soft = soft_param [kpc]                    (star/dark_softening_in_pc / 1000, :536-545) */
#define OCG_SOFT_MASS_CUBE_ROOT 1  /* soft = (m / soft_param)^(1/3) / 1000       (star/dark_char_mass, :531-534, :540-543)     */
#define OCG_SOFT_GAS_SMOOTHING 2   /* soft = 2.8 * hsml / 1000                   (gas, :547)                                   */
int ocg_assemble_sources(ocg_ctx* ctx, const double* pos_dev, const double* mass_dev, const int64_t* id_dev,
                         const double* hsml_dev, int64_t n, int64_t exclude_id, double rmax, int32_t soft_rule,
                         double soft_param, double soft_scale, const double center[3], float* out_xyzm_dev,
                         float* out_soft_dev, int64_t out_offset, int64_t* n_kept_host, void* stream);

/* ---- K1: field build, softened direct sum ---------------------------------------------------
 * Replaces ConstructKDTree + GetAccelParallel (gizmo_interface.py:561,564,566) in the theta->0
 * limit:  acc[c][t] (+)= G * sum_s m_s K(|x_s-x_t|, soft_s) (x_s-x_t)[c]
 *         pot[t]    (+)= G * sum_s m_s P(|x_s-x_t|, soft_s)            (P<0)
 * src_xyzm_dev : [n_src][4] fp32 (x,y,z,m)  recentred (ocg_recentre_f64)
 * src_soft_dev : [n_src] fp32 softening length per source; NULL = all zero (Newtonian)
 * tgt_xyzw_dev : [n_tgt][4] fp32 (x,y,z,ignored)
 * acc_dev      : [3][n_tgt] fp64 (component-major: the three arrays the reference returns at
 *                gizmo_interface.py:573); pot_dev: [n_tgt] fp64 or NULL.
 * accumulate   : 0 overwrite, 1 add to what is there (source chunks streamed through HBM).
 * Pairs with r == 0 contribute nothing unless Plummer with soft > 0 (then acc 0, pot -m/soft). */
int ocg_field_direct(ocg_ctx* ctx, const float* src_xyzm_dev, const float* src_soft_dev,
                     int64_t n_src, const float* tgt_xyzw_dev, int64_t n_tgt, int kernel, double G,
                     double* acc_dev, double* pot_dev, int accumulate, void* stream);

/* ---- K1b: frame subtraction (gizmo_interface.py:566,569-571) --------------------------------
 * acc[c][t] -= acc[c][center_row] for all t (the centre row becomes exactly 0).               */
int ocg_frame_subtract(ocg_ctx* ctx, double* acc_dev, int64_t n_tgt, int64_t center_row,
                       void* stream);

/* ---- K1 host form: the whole of _populate_grid_acceleration_ (gizmo_interface.py:512-573) ---
 * Host fp64 in, host fp64 out; does H2D, recentre on `center`, K1, optional K1b, D2H.
 * src_pos_host [n_src][3], src_mass_host [n_src], src_soft_host [n_src] (NULL = 0),
 * tgt_pos_host [n_tgt][3] (= grid.evolved_grid), center[3] (= grid.ss_evolved_position),
 * center_row: row of tgt that sits at `center` (the appended origin, grid_cartesian.py:66-67)
 * or -1 for no subtraction. acc_host [3][n_tgt]; pot_host [n_tgt] or NULL.
 * The sources are streamed through HBM in chunks of 2^26 particles, so n_src is not limited by device
 * memory or by the 2^31 particles-per-call limit of ocg_field_direct.                              */
int ocg_field_build_host(ocg_ctx* ctx, const double* src_pos_host, const double* src_mass_host,
                         const double* src_soft_host, int64_t n_src, const double* tgt_pos_host,
                         int64_t n_tgt, const double center[3], int64_t center_row, int kernel,
                         double G, double* acc_host, double* pot_host);

/* ---- K2: pack / blend grid planes ------------------------------------------------------------
 * Pack one snapshot's field (fp64 [3][n_node] acc, optional [n_node] pot) into the node-record
 * layout K3 gathers from: rec[node] = float4(ax, ay, az, phi).  n_node = nx*ny*nz + 1
 * (grid_cartesian.py:59-69: C order, x outer, z inner, origin appended).                        */
int ocg_pack_planes(ocg_ctx* ctx, const double* acc_dev, const double* pot_dev, int64_t n_node,
                    float* rec_dev, void* stream);
/* Materialised time blend (gizmo_interface.py:607-620 in its linear form):
 * out[c][i] = (1-w_b)*rec_a[i][c] + w_b*rec_b[i][c]  for c in 0..2 (acc) ; pot_out nullable.   */
int ocg_grid_time_blend(ocg_ctx* ctx, const float* rec_a_dev, const float* rec_b_dev, double w_b,
                        int64_t n_node, double* acc_out_dev, double* pot_out_dev, void* stream);

/* ---- K3: get_gravity_at_point (gizmo_interface.py:677-717) as trilinear + linear-in-time -----
 * Grid lattice: nodes along axis d are node_d[i] + origin[cluster][d], i in [0,n[d]), with
 * node_d the fp64 np.linspace arrays of grid_cartesian.py:29-31 (device copies).
 * Cell along d = searchsorted(node_d + origin_d, x, side='right') - 1 clamped to [0, n[d]-2]
 * (bit-exact with the oracle); weights linear, unclamped (linear extrapolation outside).
 * rec_a_dev/rec_b_dev: [n_cluster][n_node] float4 records of the two bracketing snapshots,
 * w_b in [0,1] the weight of b.  star_{x,y,z}_dev fp64 [n_star] (the three vectors BRIDGE passes,
 * unwrapped to kpc).  star_cluster_dev: int32 [n_star] cluster of each star, NULL = all 0.
 * origin_dev: fp64 [n_cluster][3] (evolve_grid, gizmo_interface.py:640-642).
 * Outputs fp64: acc_out_dev [3][n_star]; pot_out_dev [n_star] nullable.                         */
typedef struct ocg_grid_desc {
  int32_t n[3];           /* nodes per axis (grid_cartesian.py:25-27) */
  int32_t n_cluster;      /* number of independent grids in the batch */
  const double* node_dev[3]; /* fp64 node coordinates per axis, device */
  const double* origin_dev;  /* fp64 [n_cluster][3], device */
} ocg_grid_desc;

int ocg_grid_interp(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* rec_a_dev,
                    const float* rec_b_dev, double w_b, const double* star_x_dev,
                    const double* star_y_dev, const double* star_z_dev,
                    const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                    double* pot_out_dev, int32_t* cell_out_dev /* [3][n_star] nullable */,
                    void* stream);

/* The same evaluation with 1..4 record planes blended in time: v = ((r0*w0 + r1*w1) + r2*w2) + r3*w3 in FP32.
 * With the four non-zero cubic B-spline basis weights and the matching coefficient planes this is the
 * reference's own time interpolation (splrep/splev per grid point, gizmo_interface.py:587-620): all grid
 * points share one knot vector, so splev collapses to this blend.
 * rec_dev: HOST array of n_rec device pointers, each [n_cluster][n_node] float4; weights: HOST array [n_rec]. */
int ocg_grid_interp_multi(ocg_ctx* ctx, const ocg_grid_desc* grid, const float* const* rec_dev,
                          const double* weights, int32_t n_rec, const double* star_x_dev,
                          const double* star_y_dev, const double* star_z_dev,
                          const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                          double* pot_out_dev, int32_t* cell_out_dev, void* stream);

/* Two-level form: the reference's nested fine grid (grid.add_fine_grid, grid_cartesian.py:34-53,71-91; its default
 * configuration, test_options:93-100).  A star inside the closed fine box [node2[0]+o, node2[n2-1]+o]^3 is
 * interpolated on the fine lattice, any other star on the coarse one (whose points inside the fine box, dropped by
 * grid_cartesian.py:71-81, are filled in from the fine lattice when the records are laid out).  `fine` shares
 * n_cluster and origin_dev with `coarse`; fine == NULL (and rec_fine_dev == NULL) is the single-level call.
 * rec_*_dev: HOST arrays of n_rec device pointers; weights: HOST array [n_rec].
 * tensor_out_dev: fp64 [9][n_star] or NULL — tensor[3*i + j][s] = d a_j / d x_i of the interpolant at star s, the
 *   T[i][j] of get_tidal_tensor_at_point (gizmo_interface.py:719-756), in acceleration units per length unit.
 * level_out_dev: int32 [n_star] or NULL — 0 coarse, 1 fine.                                              */
int ocg_grid_interp_nested(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                           const float* const* rec_coarse_dev, const float* const* rec_fine_dev,
                           const double* weights, int32_t n_rec, const double* star_x_dev,
                           const double* star_y_dev, const double* star_z_dev,
                           const int32_t* star_cluster_dev, int64_t n_star, double* acc_out_dev,
                           double* pot_out_dev, double* tensor_out_dev, int32_t* level_out_dev,
                           int32_t* cell_out_dev, void* stream);

/* Graph-replayable K3.  A captured CUDA graph freezes kernel parameters, but the time-blend weights change every
 * step; these two calls move them to constant memory: ocg_set_interp_weight_slots() writes `n_slots` weight
 * quadruples (HOST fp64 [n_slots][4], rounded to fp32) into slots first_slot.. (4 slots exist) in stream order —
 * call it before each replay — and ocg_grid_interp_slot() is ocg_grid_interp_nested() (without tensor/level/cell
 * outputs) reading its weights from slot `w_slot`.  Used by bridge.Bridge to replay a whole K(dt/2) D(dt) K(dt/2)
 * step (oc_nbody.py:56) as one graph launch.                                                                  */
int ocg_set_interp_weight_slots(ocg_ctx* ctx, const double* weights_host, int32_t first_slot, int32_t n_slots,
                                void* stream);
int ocg_grid_interp_slot(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                         const float* const* rec_coarse_dev, const float* const* rec_fine_dev, int32_t w_slot,
                         int32_t n_rec, const double* star_x_dev, const double* star_y_dev,
                         const double* star_z_dev, const int32_t* star_cluster_dev, int64_t n_star,
                         double* acc_out_dev, double* pot_out_dev, void* stream);

/* get_gravity_at_point / get_tidal_tensor_at_point with the reference's OWN spatial interpolation
 * (gizmo_interface.py:651-756; options nclose / basis / order, options.py:43-45): per star the
 * `nclose` nearest points of the evolved grid (cKDTree.query over lattice + appended origin row) and the
 * polyharmonic-spline RBF interpolant with polynomial terms of degree <= `order`
 * (rbf.interpolate.RBFInterpolant(..., basis=phs<phs>, order=order)), evaluated at the star.
 * field_dev: fp64 [n_comp][n_cluster][n_node] — the time-evaluated grid arrays the reference interpolates
 *   (grid.evolved_acceleration_x/y/z, evolved_potential), n_comp in 1..4; out_dev fp64 [n_comp][n_star].
 * nclose + C(order+3,3) <= 206 (the reference's 150 + 56); phs in {1,3,5,7}; include_origin: the appended origin row
 *   (grid_cartesian.py:66-67) takes part in the neighbour search (0 when it duplicates a lattice node).
 * tensor_out_dev: fp64 [3][n_comp][n_star] or NULL — d out_c / d x_i (T[i][j] of gizmo_interface.py:719-756 for c = j).
 * status_out_dev: int32 [n_star] or NULL — 0 ok; bit 0 neighbour window truncated (star far outside the grid),
 *   bit 1 refinement not converged (ill-conditioned stencil, e.g. clipped by the grid edge), bit 2 zero pivot;
 *   bit 3 see `embedded`; bits 8-15: refinement iterations used (status & 0xff == 0 means a good result).
 * neighbors_out_dev: int64 [nclose][n_star] or NULL — point-list indices of the neighbours, nearest first, ties by
 *   index.
 * embedded: 0 for the single-level grid.  1: `grid` is the FINE lattice of the reference's nested grid
 *   (grid_cartesian.py:34-53,71-91; field_dev then holds the fine rows + origin row of the point list).  Every kept
 *   coarse point lies on or outside the fine box, so the neighbour set is exact whenever the nclose-th neighbour is
 *   closer than the box surface; a star for which it is not gets status bit 3 (mixed-level stencil: not evaluated
 *   faithfully).                                                                                                 */
int ocg_grid_interp_rbf(ocg_ctx* ctx, const ocg_grid_desc* grid, const double* field_dev, int32_t n_comp,
                        int32_t nclose, int32_t order, int32_t phs, int32_t include_origin, int32_t embedded,
                        const double* star_x_dev, const double* star_y_dev, const double* star_z_dev,
                        const int32_t* star_cluster_dev, int64_t n_star, double* out_dev,
                        double* tensor_out_dev, int32_t* status_out_dev, int64_t* neighbors_out_dev,
                        void* stream);

/* The same interpolant on the reference's NESTED grid with BOTH levels searched (its default configuration,
 * test_options:93-100): the point list is kept coarse points | fine lattice | origin row (grid_cartesian.py:71-91), and a
 * star within a few fine cells of the fine-box surface has neighbours on both levels.  `coarse` / `fine`: the two lattices
 * (same n_cluster and origin_dev); coarse_row_dev int32 [ncx*ncy*ncz]: the row of each coarse lattice node in the point
 * list, -1 for the nodes the reference dropped; fine_row0: first fine row; field_dev fp64 [n_comp][n_cluster][n_point]
 * over the WHOLE point list; neighbors_out_dev receives point-list rows.  Other arguments as ocg_grid_interp_rbf.
 * Stars too far outside the grid for nclose points to be found get NaN and status bit 0.                          */
int ocg_grid_interp_rbf_nested(ocg_ctx* ctx, const ocg_grid_desc* coarse, const ocg_grid_desc* fine,
                               const int32_t* coarse_row_dev, int64_t fine_row0, int64_t n_point, const double* field_dev,
                               int32_t n_comp, int32_t nclose, int32_t order, int32_t phs, int32_t include_origin,
                               const double* star_x_dev, const double* star_y_dev, const double* star_z_dev,
                               const int32_t* star_cluster_dev, int64_t n_star, double* out_dev, double* tensor_out_dev,
                               int32_t* status_out_dev, int64_t* neighbors_out_dev, void* stream);

/* K2 pack with a scatter: rec[index[i]] = float4(acc[0][i], acc[1][i], acc[2][i], pot[i]) for i < n.
 * Lays rows of the reference's point list (kept coarse points | fine lattice | origin row,
 * grid_cartesian.py:71-91) out as full-lattice node records.  acc_dev fp64 [3][n]; pot_dev [n] or NULL;
 * index_dev int64 [n].                                                                                    */
int ocg_pack_planes_indexed(ocg_ctx* ctx, const double* acc_dev, const double* pot_dev, int64_t n,
                            const int64_t* index_dev, float* rec_dev, void* stream);

/* ---- K4: cluster self-gravity (ph4 force loop behind oc_code.py:218-229) ---------------------
 * Plummer direct sum inside each segment (cluster) of a batch.
 * pos_dev fp64 [3][n] component-major, mass_dev fp64 [n]; seg_offsets_host int64 [n_seg+1]
 * (host array; NULL with n_seg == 1 means one segment [0,n)).  eps2 = epsilon_squared
 * (oc_code.py:225).  Each segment is recentred on its first particle before rounding to fp32.
 * tgt_begin/tgt_end: the rank's shard of targets (global particle indices); outputs are written
 * for that range only, at the same global indices.  Self term excluded.
 * acc_dev fp64 [3][n]; pot_dev fp64 [n] nullable.                                               */
int ocg_self_gravity(ocg_ctx* ctx, const double* pos_dev, const double* mass_dev, int64_t n,
                     const int64_t* seg_offsets_host, int32_t n_seg, double eps2, double G,
                     int64_t tgt_begin, int64_t tgt_end, double* acc_dev, double* pot_dev,
                     void* stream);

/* ---- multi-GPU: one process per GPU on one NVLink / NVSwitch node ----------------
 * The reference has no multi-GPU path (its only parallel construct is a multiprocessing.Pool, gizmo_interface.py:600-605).
 * Partition (BASELINE.json north_star): K1 source-sharded — every rank computes the full-grid partial field of its share
 * of the snapshot, then an all-reduce; K4 star-sharded — every rank owns a block of stars and needs all positions at
 * every self-gravity evaluation.  The exchange runs over peer memory: each rank owns a WINDOW of device memory that every
 * other rank maps (CUDA IPC), and the kernels load / store the peers' windows directly, synchronising through flags in
 * them.  Set-up, once per process:
 *   1. ocg_comm_create  allocates this rank's window and fills an opaque OCG_COMM_HANDLE_BYTES handle;
 *   2. the caller all-gathers the handles by any means it has (MPI_Allgather, torch.distributed, a shared file);
 *   3. ocg_comm_connect maps the peers' windows (handles in rank order).
 * Every rank must then issue the same sequence of ocg_comm_* / *_sharded calls.  The force exchanges work in the first
 * half of the window, the all-reduce in the second (they may follow one another without a barrier).  window_bytes: at
 * least 96 * ceil(n_stars / nranks) for ocg_self_gravity_sharded, 288 * ceil(n_stars / nranks) with the Hermite form;
 * the all-reduce works through whatever size it is given (32 bytes per element for a single pass).  One sharded
 * cluster per communicator.  nranks <= 16.                                                                          */
#define OCG_COMM_HANDLE_BYTES 128
int ocg_comm_create(ocg_ctx* ctx, int32_t rank, int32_t nranks, int64_t window_bytes, void* handle_out);
int ocg_comm_connect(ocg_ctx* ctx, const void* all_handles /* [nranks][OCG_COMM_HANDLE_BYTES] */);
int ocg_comm_destroy(ocg_ctx* ctx); /* also done by ocg_destroy */
int ocg_comm_info(const ocg_ctx* ctx, int32_t* rank, int32_t* nranks, int64_t* window_bytes);
/* Synchronises `stream` and reports whether an exchange kernel gave up waiting for a peer (~2 s poll limit). */
int ocg_comm_status(ocg_ctx* ctx, void* stream);
/* buf[i] <- sum over the ranks of their buf[i], in place, fp64, summed in rank order: deterministic and bit-identical on
 * every rank (the all-reduce of the K1 partial fields, 3*(Ngrid+1) elements).  One kernel per window-full. */
int ocg_comm_allreduce_f64(ocg_ctx* ctx, double* buf_dev, int64_t n, void* stream);
/* K4 for this rank's block of ONE cluster of n stars, block = [rank*n/nranks ...) with sizes differing by at most one
 * (the first n % nranks ranks hold one more).  pos_local_dev fp64 [3][n_local] (this rank's stars only), mass_all_dev fp64
 * [n] (replicated); acc_local_dev fp64 [3][n_local], pot_local_dev [n_local] or NULL.  Two launches: one kernel publishes
 * the block, waits for the peers', reads all blocks over NVLink and writes the FP32 source tiles (recentred on star 0,
 * exactly the single-GPU inputs, so the sharded trajectory equals the single-GPU one); then the force kernel for the
 * rank's target rows.  eps2 > 0.                                                                                    */
int ocg_self_gravity_sharded(ocg_ctx* ctx, const double* pos_local_dev, const double* mass_all_dev, int64_t n, double eps2,
                             double G, double* acc_local_dev, double* pot_local_dev, void* stream);
/* The same for the Hermite force (K6, oc_code.py:218-229 with ph4): the rank's block of positions AND velocities
 * ([3][n_local] each) is published, all blocks are read through the mapped windows and packed into the 7-array tiles.
 * acc_dev / jerk_dev ([3][n]) and pot_dev ([n] or NULL) are FULL-size arrays of which the rank's rows [a, b) are written
 * (the layout ocg_self_gravity_hermite uses for a target range). */
int ocg_self_gravity_hermite_sharded(ocg_ctx* ctx, const double* pos_local_dev, const double* vel_local_dev,
                                     const double* mass_all_dev, int64_t n, double eps2, double G, double vel_to_len,
                                     double* acc_dev, double* jerk_dev, double* pot_dev, void* stream);

/* ---- K6: Hermite force loop + 4th-order Hermite predictor / corrector ------
 * The arithmetic of the ph4 worker itself (4th-order Hermite, oc_code.py:218-229; options.py:248-253):
 *   acc[c][i]  = G sum_j m_j d_c / r^3                                  d = x_j - x_i, r^2 = |d|^2 + eps2
 *   jerk[c][i] = G vel_to_len sum_j m_j [ w_c / r^3 - 3 (d.w) d_c / r^5 ]   w = v_j - v_i
 * vel_dev fp64 [3][n] in the caller's velocity unit; vel_to_len converts it to length/time of the
 * acceleration's time unit (kpc, km/s, Myr: 1.0227e-3), so jerk is in acceleration units per time.
 * Positions AND velocities of each segment are recentred on its first particle before rounding to fp32.
 * Other arguments as ocg_self_gravity; jerk_dev fp64 [3][n]; pot_dev nullable.  ph4's block time steps
 * are not reproduced: all stars share one step (see ocg_hermite_correct for the step criterion).      */
int ocg_self_gravity_hermite(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev,
                             const double* mass_dev, int64_t n, const int64_t* seg_offsets_host,
                             int32_t n_seg, double eps2, double G, double vel_to_len, int64_t tgt_begin,
                             int64_t tgt_end, double* acc_dev, double* jerk_dev, double* pot_dev,
                             void* stream);
/* Predictor: pos_pred = pos + ((vel*dt + acc*(dt^2/2)) + jerk*(dt^3/6)) * vel_to_len,
 *            vel_pred = vel + (acc*dt + jerk*(dt^2/2)); fp64 [3][n], every op rounded separately.     */
int ocg_hermite_predict(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* acc_dev,
                        const double* jerk_dev, int64_t n, double dt, double vel_to_len,
                        double* pos_pred_dev, double* vel_pred_dev, void* stream);
/* Corrector (Makino & Aarseth 1992) from the force at the predicted state (acc1, jerk1):
 *   a2 = (-6 (a0 - a1) - dt (4 j0 + 2 j1)) / dt^2      a3 = (12 (a0 - a1) + 6 dt (j0 + j1)) / dt^3
 *   vel = vel_pred + a2 dt^3/6 + a3 dt^4/24            pos = pos_pred + (a2 dt^4/24 + a3 dt^5/120) vel_to_len
 * then acc0 <- acc1, jerk0 <- jerk1.  dt_min_dev (DEVICE fp64 scalar, nullable) receives the minimum over the
 * stars of the Aarseth step sqrt(eta (|a1||a2'| + |j1|^2) / (|j1||a3| + |a2'|^2)), a2' = a2 + dt a3.          */
int ocg_hermite_correct(ocg_ctx* ctx, double* pos_dev, double* vel_dev, double* acc0_dev, double* jerk0_dev,
                        const double* pos_pred_dev, const double* vel_pred_dev, const double* acc1_dev,
                        const double* jerk1_dev, int64_t n, double dt, double vel_to_len, double eta,
                        double* dt_min_dev, void* stream);

/* ph4's individual BLOCK TIME STEPS (AMUSE ph4 behind oc_code.py:218-229; options.py:248-253): advance one cluster by
 * `span` with every star on its own power-of-two step span / 2^k, k <= max_level (<= 20), chosen from the Aarseth criterion
 * with accuracy parameter eta (start-up: eta |a| / |j|).  Per block step: t_next = min (t_i + dt_i); every star is predicted
 * to t_next; acceleration and jerk of the ACTIVE stars (those that reach t_next) from ALL predicted stars (K6 over active
 * targets x all sources); the active stars are corrected over their own step and pick their next one (halved as often as the
 * criterion demands, doubled at most once, and only onto a commensurate time).  All stars arrive at t + span together.
 * pos_dev, vel_dev fp64 [3][n] are advanced in place; acc_dev, jerk_dev fp64 [3][n] receive each star's end-of-step force
 * (evaluated at the predicted end state, as the corrector uses it).
 * The call synchronises `stream` once per block step (16 bytes come back to size the next launches).  Host counters
 * (nullable): block steps taken, star-steps taken (sum of the active counts; a shared step would take n * 2^k_max).     */
int ocg_hermite_block_evolve(ocg_ctx* ctx, double* pos_dev, double* vel_dev, const double* mass_dev, double* acc_dev,
                             double* jerk_dev, int64_t n, double eps2, double G, double vel_to_len, double span, double eta,
                             int32_t max_level, int64_t* n_block_steps_host, int64_t* n_star_steps_host, void* stream);

/* ---- K5: BRIDGE kick / drift (amuse.couple.bridge kick + leapfrog drift, oc_nbody.py:56) ----
 * vel[c][i] += dt * acc[c][i]   ;   pos[c][i] += dt * vel[c][i] * vel_to_len
 * All fp64 [3][n] component-major; mul and add rounded separately (no FMA) so that a numpy
 * `v + a*dt` reproduces the bits.                                                               */
int ocg_kick(ocg_ctx* ctx, double* vel_dev, const double* acc_dev, int64_t n, double dt,
             void* stream);
int ocg_drift(ocg_ctx* ctx, double* pos_dev, const double* vel_dev, int64_t n, double dt,
              double vel_to_len, void* stream);
/* acc_sum = a + scale_b * b  (combine self-gravity with the tidal kick in other units) */
int ocg_axpy(ocg_ctx* ctx, double* y_dev, const double* x_dev, double a, int64_t n, void* stream);

/* ---- per-step cluster bookkeeping of the driver loop, on the device ----------------------------
 * Bound subset + its centre of mass (oc_nbody.py:60-61: particles.bound_subset().center_of_mass(), fed to
 * evolve_grid at oc_nbody.py:64).  A star is bound iff 0.5*|v - v_com|^2 + pot_to_v2*pot < 0, v_com the velocity of
 * the cluster's centre of mass; when no star is bound every star counts.  One result row per segment (cluster):
 *   out_dev[seg][0..2] centre of mass of the bound stars, [3] their mass, [4] their number, [5..7] v_com.
 * pos_dev/vel_dev fp64 [3][n], mass_dev/pot_dev fp64 [n] (pot from ocg_self_gravity); seg_offsets_dev: DEVICE
 * int64 [n_seg+1] or NULL for one segment [0,n); bound_mask_dev: uint8 [n] or NULL.                            */
int ocg_bound_com(ocg_ctx* ctx, const double* pos_dev, const double* vel_dev, const double* mass_dev,
                  const double* pot_dev, int64_t n, const int64_t* seg_offsets_dev, int32_t n_seg,
                  double pot_to_v2, double* out_dev, uint8_t* bound_mask_dev, void* stream);
/* Ejection cut of clean_ejections (oc_code.py:231-246): keep[i] = 0 iff |len_scale*(x_i - median(x))| > cut,
 * the median taken per axis exactly as numpy.median does (mean of the two middle order statistics).
 * median_out_dev: fp64 [3] or NULL.                                                                           */
int ocg_eject_mask(ocg_ctx* ctx, const double* pos_dev, int64_t n, double len_scale, double cut,
                   uint8_t* keep_mask_dev, double* median_out_dev, void* stream);
/* Stable compaction (particles.remove_particles, oc_code.py:241-242): out[r][j] = in[r][i_j] for the j-th kept
 * index i_j; in_dev fp64 [rows][n], out_dev fp64 with row stride out_stride (NULL: count only);
 * n_keep_dev: DEVICE int64 receiving the number kept (nullable).                                              */
int ocg_compact_rows(ocg_ctx* ctx, const double* in_dev, int32_t rows, int64_t n, const uint8_t* keep_mask_dev,
                     double* out_dev, int64_t out_stride, int64_t* n_keep_dev, void* stream);

/* ---- probes (roofline denominators measured live by bench.py) ------------------------------- */
/* Runs an FFMA-only kernel (packed=0: FFMA, packed=1: FFMA2) over the whole GPU and returns
 * achieved TFLOP/s (2 flop per lane-FMA). which: 0 FFMA, 1 FFMA2, 2 MUFU.RSQ (G ops/s).     */
double ocg_probe_throughput(ocg_ctx* ctx, int which);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OCG_H */

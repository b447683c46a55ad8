/*
 * ocg_debug.h — tuning and test hooks of liboc_nbody_b200.  NOT part of the drop-in ABI (include/ocg.h): nothing the
 * reference-side integration calls lives here.  Used by tests/ (to force both branches of a heuristic) and tools/
 * (shape sweeps, accuracy surveys).
 *
 * Every knob is state of ONE ocg_ctx (no process-global state), and none of them can make a shipped kernel return
 * wrong numbers: the timing-only experiments (kernels that skip part of the arithmetic to bound a pipe's cost) exist
 * only in the separate -DOCG_TUNING build of the library (oc_nbody_b200/build.py: build_library(tuning=True) ->
 * lib/liboc_nbody_b200_tuning.so); in the default build their table entries are empty and ocg_debug_set refuses them.
 */
#ifndef OCG_DEBUG_H
#define OCG_DEBUG_H

#include "ocg.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

enum {
  OCG_KNOB_DIRECT_VARIANT = 0,     /* K1/K4 kernel shape id (direct_sum.cu table), -1 = heuristic (default)            */
  OCG_KNOB_PRECISE_NEAR = 1,       /* 1 (default): sources inside the precision radius take the FP64 pair path          */
  OCG_KNOB_MASS_FOLD = 2,          /* 1 (default): K1 without potential uses mass-folded tiles                          */
  OCG_KNOB_SMALL_CLUSTER_PATH = 3, /* 1 (default): a single cluster <= 4096 stars takes the fused one-launch K4 kernel   */
  OCG_KNOB_HOST_CHUNK = 4,         /* particles staged at a time by ocg_field_build_host (default 2^26; <= 0 restores it) */
  OCG_KNOB_HERMITE_VARIANT = 5,    /* K6 kernel shape id, -1 = production (default)                                     */
  OCG_KNOB_HERMITE_SMALL_PATH = 6, /* 1 (default): fused one-launch K6 kernel for a single small cluster                */
  OCG_KNOB_INTERP_VARIANT = 7,     /* K3 register bound: 0 <= 128, 1 <= 80, 2 <= 64 registers (default 2)               */
  /* 8: retired (a factorisation shared between stars was measured and dropped) */
  OCG_KNOB_NEAR_CAP = 9,           /* size limit of K1's FP64 precision-radius set; 0 (default) = max(n_src / 512, 2^36 / n_src)  */
  OCG_KNOB_PASS_BYTES = 10         /* K1: bytes of source tiles per stream-K pass (L2 residency); 0 = one pass, default 32 MiB      */
};

/* Set one knob of this ctx.  OCG_ERR_INVALID (text in ocg_last_error) for an unknown knob, an out-of-range value, or a
 * kernel shape that is not compiled into this build. */
int ocg_debug_set(ocg_ctx* ctx, int knob, int64_t value);
/* Shape tables: family 0 = K1/K4 direct sum, 1 = K6 Hermite.  Ids are stable across builds (profiles/ quotes them);
 * ocg_debug_variant_built tells whether this build carries the kernel. */
int ocg_debug_variant_count(int family);
const char* ocg_debug_variant_name(int family, int id);
int ocg_debug_variant_built(int family, int id);
/* K7 phase counters (cycles summed over CTAs since the last call: select, assemble, factorise, residual, solve, output).
 * Only the OCG_TUNING build counts; the default build returns zeros. */
int ocg_debug_rbf_phase_cycles(ocg_ctx* ctx, double* out6);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OCG_DEBUG_H */

/*
 * lerf_b200_testing.h -- test and tuning switches of liblerf_b200.so.  NOT part of the drop-in interface
 * (include/lerf_b200.h): integrators never call these.  The GPU parity tests use them to run a second
 * implementation of a kernel against the production one, scripts/kbench.py to time tuning variants.
 *
 * Every switch is PER CALLING THREAD (one thread_local struct in the library): it changes the kernels that the
 * calling thread launches afterwards and nothing else.  Variants marked "(x)" exist only in a library built with
 * -DLERF_EXPERIMENTS (lerf_pytorch_b200/liblerf_b200_exp.so, `python -m lerf_pytorch_b200.build --experiments`);
 * the product library silently keeps its production kernel for them.
 */
#ifndef LERF_B200_TESTING_H_
#define LERF_B200_TESTING_H_

#include "lerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 1 when the library was built with -DLERF_EXPERIMENTS. */
int lerf_build_has_experiments(void);

/* Kernel behind lerf_lut_stage1 (stage = 1) / lerf_lut_stage2 (stage = 2).  0 = production (stage 1: cell-packed tables;
 * stage 2: paired-window tables for oC = 3, cell-packed for oC = 1).  20..39 = cell-packed-table kernel (20 + n: (x) tuning
 * variant n; 27 = one sort per lookup instead of production's two lookups per 16x2 sorting network, in every build), 80.. = paired-window kernel (stage 2; stage 1 and the cell-pair format 90.. are (x)), 1..19 = (x) row-major
 * table kernel, 40..59 = (x) table-format mix, 60..79 = (x) max-tap block kernel (production of round 1).
 * All variants produce identical bytes. */
void lerf_debug_lut_variant(int stage, int variant);
/* Weights of the cell-block swizzle baked into the tables by the NEXT lerf_luts_create on this thread
 * (0,0,0 = plain layout; default 9,5,3).  Results never depend on it. */
void lerf_debug_cell_hash(int ha, int hb, int hc);

/* 0 = production dispatch (cell-owner kernel for integer scales, any-scale cell kernel for other scales from x3 per axis
 * up, tile kernel for the remaining scales >= 1, fast warp kernel); 1 = only the float64 operation-order kernels (the
 * parity path); 2 = like 0 without the cell-owner kernels, so every scale takes the tile kernel; 3 = the any-scale cell
 * kernel wherever it applies (scales in [1, 4]), integer scales included.  All must agree within the fp32 tolerance. */
void lerf_debug_force_generic(int on);

/* lerf_warp's fast Gaussian kernel: 1 (default) = every input sample is first decoded into a 32-byte record and a tap
 * is one 256-bit gather; 0 = taps are gathered from feat/codes and decoded through tables.  Identical results. */
void lerf_debug_warp_records(int on);

/* Integer-scale resampler: 0 = production (weights relative to the phase's NEAREST tap where the plan allows it, whose weight is
 * exactly 1 -- no minimum over the taps, one exponential less per sample); 13 = weights relative to the smallest exponent
 * (production until r2g; still what the tile / cell kernels do); 7 = the nearest-tap form by name;
 * 10 = production arithmetic with the byte-store uint8 epilogue of round 1;
 * 11 = geometry factors from kernel parameters instead of immediates (what odd scales use); 12 = 11 + the interleaved
 * uint8 tile copied out by lanes (funnel-shifted 128-bit stores) instead of bulk stores; (x) 1 = hoisted-FP64 form,
 * 2 / 5 = plain form at 4 / 5 blocks per SM, 4 = production form at 5 blocks per SM.  All stay within the 1e-4 bar. */
void lerf_debug_resize_variant(int variant);

/* (x) lerf_sr_fused through the role-interleaved pipeline kernel: enabled = 0 (default) issues the three plain launches;
 * min_blocks (2..4) and group_planes (0 = auto) tune it.  Results never depend on it. */
void lerf_debug_pipeline(int enabled, int min_blocks, int group_planes);

/* Which table block lerf_luts_pin_l2 puts under the access-policy window: 0 = cell-packed block (production),
 * 1 = paired-window block.  For A/B timing; results never depend on it. */
void lerf_debug_l2_window(int which);

/* Shared-memory carve-out of the two stage kernels in percent of the SM's unified L1 / shared memory
 * (cudaFuncAttributePreferredSharedMemoryCarveout); -1 = production choice.  For A/B timing. */
void lerf_debug_carveout(int percent);

#ifdef __cplusplus
}
#endif
#endif /* LERF_B200_TESTING_H_ */

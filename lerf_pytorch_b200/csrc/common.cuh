// Shared device/host helpers for liblerf_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "lerf_b200.h"
#include "lerf_b200_testing.h"

namespace lerf {

extern thread_local std::string g_last_error;
extern thread_local long long g_launches;

int fail(int code, const char* fmt, ...);

// Test / tuning switches behind include/lerf_b200_testing.h.  They live in ONE thread-local struct: a hook only changes
// the kernels that the calling thread launches afterwards, never another thread's.  Variants marked (x) exist only in a
// library built with -DLERF_EXPERIMENTS (build.py: liblerf_b200_exp.so); the product library ignores them.
struct DebugState {
  int lut_variant[2] = {0, 0};    // per stage, see lerf_debug_lut_variant
  int cell_hash[3] = {9, 5, 3};   // swizzle baked into the NEXT LutSet's cell tables
  int resize_variant = 0;         // integer-scale resampler flavour
  int u8_staged = 1;              // 0 = byte-store uint8 epilogue of r1
  int force_generic = 0;          // 1 = float64 parity kernels only; 2 = no cell-owner kernel; 3 = the any-scale cell kernel wherever it applies
  int warp_records = 1;           // 0 = table form of the fast warp kernel
  int pipe_enabled = 0, pipe_minb = 3, pipe_group = 0;  // (x) role-interleaved pipeline kernel
  int l2_window = 0;              // lerf_luts_pin_l2: 0 = cell-packed block, 1 = paired-window block
  int carveout = -1;              // shared-memory carve-out (percent) of the stage kernels, -1 = production
};
extern thread_local DebugState g_dbg;

#define LERF_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ::lerf::fail(LERF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                               \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define LERF_LAUNCHED()                                                                      \
  do {                                                                                       \
    ++::lerf::g_launches;                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return ::lerf::fail(LERF_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                               \
  } while (0)

constexpr int kL = 17;                 // LUT grid points per axis (interval 4)
constexpr int kEntries = kL * kL * kL * kL;  // 83521
constexpr int kStrideA = kL * kL * kL;       // 4913
constexpr int kStrideB = kL * kL;            // 289
constexpr int kStrideC = kL;                 // 17
constexpr int kStrideAll = kStrideA + kStrideB + kStrideC + 1;  // 5220: base -> p1111

struct InAddr {  // strided uint8 source (see lerf_lut_stage1 in lerf_b200.h)
  int channels;
  long long batch_stride, chan_stride, row_stride, pix_stride;
};

struct lerf_luts_impl {
  int device;
  int oC2;
  int num_sms;
  void* block;          // one allocation holding every table (L2 window target)
  size_t block_bytes;
  const int8_t* s1[3];  // s, c, t            int8 [83521]
  const void* s2[6];    // s r0, s r1, c r0, c r1, t r0, t r1
                        // oC2 == 3: uint32 per entry = bytes (c0, c1, c2, 0);  oC2 == 1: int8
  // cell-packed copies (lut_cell.cuh): 65536 cells; stage 1: 16 B per cell; stage 2: 16 B (oC 1) or 64 B (oC 3,
  // channel k at [16k, 16k+16))
  void* cell_block;
  size_t cell_block_bytes;
  int cell_hash[3];  // block swizzle weights (lut_cell.cuh), fixed when the set is created
  const uint8_t* c1[3];
  const uint8_t* c2[6];
  // max-tap block copies of the oC = 3 stage-2 tables (lut_mt.cuh): 65536 cells x 4 blocks x 32 B each
  void* mt_block;
  const uint8_t* mt2[6];
  // paired-window copies (lut_pw.cuh): 6 window families per stage, 64 planes x 65536 blocks of 16 B (oC 1) / 32 B (oC 3)
  void* pw_block;
  size_t pw_block_bytes;
  const uint8_t* pw1[6];
  const uint8_t* pw2[6];
  // cell-pair copies (lut_pw.cu FmtCP, oC = 1): 6 families x 65536 blocks of 32 B; cp2 only for oC2 == 1
  void* cp_block;
  const uint8_t* cp1[6];
  const uint8_t* cp2[6];
};

struct lerf_sr_plan_impl {
  int device;
  int H, W, oH, oW;
  int* left_y;     // device [oH]   first tap row (unpadded, may be -1)
  int* left_x;     // device [oW]
  double* dist_y;  // device [2*oH] distance to tap 0 / tap 1 (row axis)
  double* dist_x;  // device [2*oW]
  int* h_left_y;   // host copy (row-band planning)
  int int_scale;   // S if out = S*in on both axes with the periodic phase pattern, else 0
  int ph_y, ph_x;  // periodic geometry: the outputs whose first tap is l are S*l + ph + m, m = 0..S-1
  double ph_dist_y[8][2], ph_dist_x[8][2];  // their distances to tap 0 / tap 1
  int general;     // lerf_sr_plan_create_ex: `support` taps per axis (dist tables hold `support` entries per output), np.pad mode
  int support;     //   of the image and the antialias distance scale of the Gaussian kind; served by the float64 support kernel only
  int pad_mode;
  double aa_scale;
  int tile_ok;     // every 32 x 32 output group has its taps in a 33 x 33 input window (any scale >= 1), |dist| <= 1: resample_tile.cu
  int tile_rows;   // output rows per block of the tile kernel: the largest of 128, 96, 64, 32 whose taps fit 33 input rows
  int* cell_y;     // device [H + 2]: cell_y[l + 1] = first output row whose first tap is >= l (l = -1 .. H), so the outputs of
  int* cell_x;     // device [W + 2]  cell l are [cell[l + 1], cell[l + 2]); built when tile_ok (resample_tile.cu cell kernel)
  int cell_max_y, cell_max_x;  // longest run
  static constexpr int kCoefSlots = 8;
  void* coef_dev[kCoefSlots];    // rsi::CoefTabs per max_sigma (resample_int.cu plan_coef_tabs), immutable once uploaded
  float coef_sigma[kCoefSlots];
  int coef_n;
};

// lut_cell.cu
int build_cell_tables(lerf_luts_impl* L, const int8_t* const host_tables[9]);
int launch_stage_cell(const lerf_luts_impl* L, int stage, const uint8_t* in, const InAddr& ia, int planes, int H, int W,
                      int y0, int y1, uint8_t* out, int variant, cudaStream_t st);

// lut_pw.cu
int build_pw_tables(lerf_luts_impl* L);
int launch_stage_pw(const lerf_luts_impl* L, int stage, const uint8_t* in, const InAddr& ia, int planes, int H, int W,
                    int y0, int y1, uint8_t* out, int variant, cudaStream_t st);

int launch_stage2_mix(const lerf_luts_impl* L, const uint8_t* feat, int planes, int H, int W, int y0, int y1, uint8_t* out,
                      int variant, cudaStream_t st);

int launch_stage2_mt(const lerf_luts_impl* L, const uint8_t* feat, int planes, int H, int W, int y0, int y1, uint8_t* out,
                     int variant, cudaStream_t st);

// pipeline.cu
int sr_pipeline(const lerf_luts_impl* L, int kind, const lerf_sr_plan_impl* P, const uint8_t* in, int planes, const InAddr& ia,
                float max_sigma, int oy0, int oy1, uint8_t* feat, uint8_t* codes, void* out, int fmt, cudaStream_t st);

// resample_int.cu
int resize_sr_int_gauss(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                        float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st);

int resize_sr_int_linear(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                         float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st);


// resample_tile.cu: fast paths for uint8 code inputs; return -1 when they do not apply
int resize_sr_tile(int kind, const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                   float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st);
int resize_sr_cell(int kind, const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                   float max_sigma, int oy0, int oy1, void* out, int fmt, bool any_scale, cudaStream_t st);
int warp_fast(int kind, const uint8_t* feat, const uint8_t* codes, int planes, int channels, int H, int W, int oH, int oW,
              const double minv[9], int pad0_y, int pad0_x, int mpad0_y, int mpad0_x, int border, float max_sigma, void* out,
              int fmt, uint8_t* mask, cudaStream_t st);

}  // namespace lerf

// Stage 2 for oC = 3 with a MIX of the two table formats: device body.
// Reference being replaced: the stage-2 ensembling loop resample/eval_lut_sr.py:579-628 over
// FourSimplexInterpFaster (:24-470).
//
// Why a mix (ncu, profiles/r1c_*): with row-major tables a lookup is five scattered 32-bit gathers and the kernel is
// bound by the L1 TAG stage (~8 tag wavefronts per gather, 89 % busy); with cell-packed tables a lookup is three
// 128-bit loads and the kernel is bound by the L1 DATA stage (~15 data wavefronts per load, 97 % busy).  The two
// stages are pipelined, so giving a compile-time subset of the 12 passes (CELLMASK, bit mode*4+rot) to the cell
// path and the rest to the row-major path loads both stages instead of one.  Every pass is exact integer
// arithmetic in either format, so the bytes do not depend on the split.
#pragma once
#include "lut_cell_body.cuh"
#include "lut_rm.cuh"

namespace lerf {
namespace mix {

struct MixTables {
  const void* r[6];     // row-major uint32 (c0,c1,c2,0) tables: s r0, s r1, c r0, c r1, t r0, t r1
  const uint8_t* c[6];  // cell-packed 48-byte-cell tables, same order
  cell::Hash h;
};

constexpr int kTX = rm::kTX, kTY = rm::kTY, kHalo = rm::kHalo;
static_assert(rm::kPitch == cellk::kPitch && rm::kTX == cellk::kTX && rm::kTY == cellk::kTY, "tile geometry");
constexpr int kSmemBytes = rm::kTileBytes + 16 + cellk::kTileWords * 4;

template <unsigned CELLMASK, int M, int R>
__device__ __forceinline__ void pass(const MixTables& t, const uint8_t* c8, const uint32_t* c32, int& n0, int& n1, int& n2) {
  if ((CELLMASK >> (M * 4 + R)) & 1u)
    cellk::lookup3(t.c[2 * M + (R & 1)], cellk::simplex_at<M, R>(c32, t.h), n0, n1, n2);
  else
    rm::blend3((const uint32_t*)t.r[2 * M + (R & 1)], rm::simplex_at<M, R>(c8), n0, n1, n2);
}

// smem = kSmemBytes, 16-byte aligned.  (bxi, byi, p) = tile column, tile row, plane.
template <unsigned CELLMASK>
__device__ __forceinline__ void lut_stage2_mix_body(const MixTables& t, const uint8_t* __restrict__ feat, int H, int W, int y0,
                                                    int y1, uint8_t* __restrict__ out, int bxi, int byi, int p,
                                                    unsigned char* smem) {
  uint8_t* tile8 = smem;
  uint32_t* tile32 = reinterpret_cast<uint32_t*>(smem + (rm::kTileBytes + 15) / 16 * 16);
  const int bx = bxi * kTX, by = y0 + byi * kTY;
  const uint8_t* src = feat + (long long)p * H * W;
  const int tid = threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = min(max(by + r - kHalo, 0), H - 1), gx = min(max(bx + c - kHalo, 0), W - 1);
    const uint32_t v = __ldcg(src + (long long)gy * W + gx);
    tile8[r * rm::kPitch + c] = (uint8_t)v;
    tile32[r * cellk::kPitch + c] = cell::split_px(v);
  }
  __syncthreads();
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);  // a warp = an 8x4 pixel patch
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint8_t* c8 = tile8 + (ty + kHalo) * rm::kPitch + tx + kHalo;
  const uint32_t* c32 = tile32 + (ty + kHalo) * cellk::kPitch + tx + kHalo;
  int n0 = 0, n1 = 0, n2 = 0;
#define LERF_P(M, R) pass<CELLMASK, M, R>(t, c8, c32, n0, n1, n2);
  LERF_P(0, 0) LERF_P(0, 1) LERF_P(0, 2) LERF_P(0, 3)
  LERF_P(1, 0) LERF_P(1, 1) LERF_P(1, 2) LERF_P(1, 3)
  LERF_P(2, 0) LERF_P(2, 1) LERF_P(2, 2) LERF_P(2, 3)
#undef LERF_P
  const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
  const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
  __stcg(out + o, (uint8_t)(t0 <= 0 ? 0 : min(rm::rhe_div(t0, 192), 255)));
  __stcg(out + o + ps, (uint8_t)(t1 <= 0 ? 0 : min(rm::rhe_div(t1, 192), 255)));
  __stcg(out + o + 2 * ps, (uint8_t)(t2 <= 0 ? 0 : min(rm::rhe_div(t2, 192), 255)));
}

}  // namespace mix
}  // namespace lerf

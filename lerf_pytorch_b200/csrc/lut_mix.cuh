// Stage 2 for oC = 3 with a MIX of two table formats: device body.
// Reference being replaced: the stage-2 ensembling loop resample/eval_lut_sr.py:579-628 over
// FourSimplexInterpFaster (:24-470).
//
// Why a mix (ncu, profiles/): on row-major tables (lut_rm.cuh) a pass is five scattered 32-bit gathers that hit L1
// (89 %) and the kernel is bound by the L1 tag stage; on max-tap blocks (lut_mt.cuh) a pass is ONE 32-byte load that
// bypasses L1 and the kernel is bound by L2 bandwidth (~9 TB/s).  Those are different resources, so giving a
// compile-time subset of the 12 passes (MTMASK, bit mode*4+rot) to the max-tap path and the rest to the row-major
// path loads both.  Max-tap loads use L1::no_allocate, so they do not evict the row-major tables from L1.
// Every pass is exact integer arithmetic in either format: the bytes do not depend on the split.
#pragma once
#include "lut_mt.cuh"
#include "lut_rm.cuh"

namespace lerf {
namespace mix {

struct MixTables {
  const void* r[6];     // row-major uint32 (c0,c1,c2,0) tables: s r0, s r1, c r0, c r1, t r0, t r1
  const uint8_t* m[6];  // max-tap block tables, same order
};

constexpr int kTX = rm::kTX, kTY = rm::kTY, kHalo = rm::kHalo;
static_assert(rm::kPitch == mt::kPitch && rm::kTX == mt::kTX, "tile geometry");
constexpr int kSmemBytes = (rm::kTileBytes + 15) / 16 * 16 + (kTY + 2 * kHalo) * mt::kPitch * 8;

template <unsigned MTMASK, int M, int R>
__device__ __forceinline__ void pass(const MixTables& t, const uint8_t* c8, const uint2* c64, int& n0, int& n1, int& n2) {
  if ((MTMASK >> (M * 4 + R)) & 1u)
    mt::pass<M, R, 1>(t.m[2 * M + (R & 1)], c64, n0, n1, n2);
  else
    rm::blend3((const uint32_t*)t.r[2 * M + (R & 1)], rm::simplex_at<M, R>(c8), n0, n1, n2);
}

// smem = kSmemBytes, 16-byte aligned.  (bxi, byi, p) = tile column, tile row, plane.  256 threads, 32x8 tile.
template <unsigned MTMASK>
__device__ __forceinline__ void lut_stage2_mix_body(const MixTables& t, const uint8_t* __restrict__ feat, int H, int W, int y0,
                                                    int y1, uint8_t* __restrict__ out, int bxi, int byi, int p,
                                                    unsigned char* smem) {
  uint8_t* tile8 = smem;
  uint2* tile64 = reinterpret_cast<uint2*>(smem + (rm::kTileBytes + 15) / 16 * 16);
  const int bx = bxi * kTX, by = y0 + byi * kTY;
  const uint8_t* src = feat + (long long)p * H * W;
  const int tid = threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = min(max(by + r - kHalo, 0), H - 1), gx = min(max(bx + c - kHalo, 0), W - 1);
    const uint32_t v = __ldcg(src + (long long)gy * W + gx);
    tile8[r * rm::kPitch + c] = (uint8_t)v;
    uint2 w;
    mt::split_px2(v, w.x, w.y);
    tile64[r * mt::kPitch + c] = w;
  }
  __syncthreads();
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);  // a warp = an 8x4 pixel patch
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint8_t* c8 = tile8 + (ty + kHalo) * rm::kPitch + tx + kHalo;
  const uint2* c64 = tile64 + (ty + kHalo) * mt::kPitch + tx + kHalo;
  int n0 = 0, n1 = 0, n2 = 0;
#define LERF_P(M, R) pass<MTMASK, M, R>(t, c8, c64, n0, n1, n2);
  LERF_P(0, 0) LERF_P(0, 2) LERF_P(0, 1) LERF_P(0, 3)
  LERF_P(1, 0) LERF_P(1, 2) LERF_P(1, 1) LERF_P(1, 3)
  LERF_P(2, 0) LERF_P(2, 2) LERF_P(2, 1) LERF_P(2, 3)
#undef LERF_P
  const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
  const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
  __stcg(out + o, (uint8_t)(t0 <= 0 ? 0 : min(rm::rhe_div(t0, 192), 255)));
  __stcg(out + o + ps, (uint8_t)(t1 <= 0 ? 0 : min(rm::rhe_div(t1, 192), 255)));
  __stcg(out + o + 2 * ps, (uint8_t)(t2 <= 0 ? 0 : min(rm::rhe_div(t2, 192), 255)));
}

}  // namespace mix
}  // namespace lerf

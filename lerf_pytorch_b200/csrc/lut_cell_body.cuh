// Cell-packed-table LUT stage: device body shared by the plain kernel in lut_cell.cu and the pipeline kernel in
// pipeline.cu.  See lut_cell.cuh for the lookup primitive.
//
// Reference being replaced (ddlee-cn/LeRF-PyTorch): FourSimplexInterpFaster resample/eval_lut_sr.py:24-470 and the
// stage-1 / stage-2 ensembling loops :541-628 (= eval_lut_warp.py:104-191).
//
// One thread = one sample; the tile (+3 halo) is kept in shared memory as pre-split words (cell::split_px); each of
// the 12 passes is one 5-compare-exchange sort, one 128-bit table load per channel, two PRMT + two DP4A per channel.
#pragma once
#include "common.cuh"
#include "lut_cell.cuh"

namespace lerf {
namespace cellk {

using cell::Simplex;


constexpr int kHalo = 3;
constexpr int kTX = 32, kTY = 8;
constexpr int kPitch = 40;  // words; 40 mod 32 = 8: the four 8-word rows of a warp's 8x4 patch hit disjoint banks

template <int MODE, int R, int K>
struct Tap {  // mode pattern (eval_lut_sr.py:30-81) composed with the rotation
  static constexpr int di = MODE == 0 ? (K >> 1) : (MODE == 1 ? 0 : K);
  static constexpr int dj = MODE == 0 ? (K & 1) : K;
  static constexpr int dy = R == 0 ? di : (R == 1 ? dj : (R == 2 ? -di : -dj));
  static constexpr int dx = R == 0 ? dj : (R == 1 ? -di : (R == 2 ? -dj : di));
};

template <int MODE, int R>
__device__ __forceinline__ Simplex simplex_at(const uint32_t* c, const cell::Hash& h) {
  return cell::simplex_of(c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx],
                          c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx],
                          c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx],
                          c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx], h);
}

// Lookups (MODE, RA) and (MODE, RB) through one sorting network (cell::simplex_pair_of).
template <int MODE, int RA, int RB>
__device__ __forceinline__ void simplex_pair_at(const uint32_t* c, const cell::Hash& h, Simplex& sa, Simplex& sb) {
  const uint32_t xa[4] = {c[Tap<MODE, RA, 0>::dy * kPitch + Tap<MODE, RA, 0>::dx], c[Tap<MODE, RA, 1>::dy * kPitch + Tap<MODE, RA, 1>::dx],
                          c[Tap<MODE, RA, 2>::dy * kPitch + Tap<MODE, RA, 2>::dx], c[Tap<MODE, RA, 3>::dy * kPitch + Tap<MODE, RA, 3>::dx]};
  const uint32_t xb[4] = {c[Tap<MODE, RB, 0>::dy * kPitch + Tap<MODE, RB, 0>::dx], c[Tap<MODE, RB, 1>::dy * kPitch + Tap<MODE, RB, 1>::dx],
                          c[Tap<MODE, RB, 2>::dy * kPitch + Tap<MODE, RB, 2>::dx], c[Tap<MODE, RB, 3>::dy * kPitch + Tap<MODE, RB, 3>::dx]};
  cell::simplex_pair_of(xa, xb, h, sa, sb);
}

__device__ __forceinline__ int lookup1(const uint8_t* __restrict__ tab, const Simplex& s) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(tab) + s.cell);
  return cell::blend(q.x, q.y, q.z, q.w, s);
}

// oC = 3: 48-byte cells, channel k at bytes [16k, 16k+16): three 128-bit loads from two 32-byte sectors
__device__ __forceinline__ void lookup3(const uint8_t* __restrict__ tab, const Simplex& s, int& n0, int& n1, int& n2) {
  const uint4* p = reinterpret_cast<const uint4*>(tab) + s.cell * 3u;
  const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  n0 += cell::blend(a.x, a.y, a.z, a.w, s);
  n1 += cell::blend(b.x, b.y, b.z, b.w, s);
  n2 += cell::blend(c.x, c.y, c.z, c.w, s);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ int rhe_div(int num, int den) {  // round_half_even(num / den), num > 0, den even
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;
  return q;
}

struct CellTables {
  const uint8_t* t[6];
  cell::Hash h;
};

constexpr int kTileWords = (kTY + 2 * kHalo) * kPitch;

// `tile` = kTileWords words of shared memory; (bxi, byi, p) = the block's tile column, tile row and plane.
template <int STAGE, int OC, bool PAIRED = false>
__device__ __forceinline__ void lut_stage_cell_body(const CellTables& tabs, const uint8_t* __restrict__ in, const InAddr& ia,
                                                    int H, int W, int y0, int y1, uint8_t* __restrict__ out, int bxi,
                                                    int byi, int p, uint32_t* tile) {
  const int bx = bxi * kTX, by = y0 + byi * kTY;
  const uint8_t* src = in + (long long)(p / ia.channels) * ia.batch_stride + (long long)(p % ia.channels) * ia.chan_stride;
  const int tid = threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = clampi(by + r - kHalo, 0, H - 1), gx = clampi(bx + c - kHalo, 0, W - 1);
    tile[r * kPitch + c] = cell::split_px(__ldcg(src + (long long)gy * ia.row_stride + (long long)gx * ia.pix_stride));
  }
  __syncthreads();
  // a warp covers an 8x4 pixel patch: 2-D neighbours have closer values than the ends of a 32-pixel row, so the 32
  // cells of one load fall into fewer cache lines
  const int lane = tid & 31, wrp = tid >> 5;
  // inside the patch a QUARTER-warp (the unit a 128-bit load is served in) covers 2 x 4 pixels: measured 186.8 us per frame
  // against 187.4 for 4 x 2 and 189.5 for 8 x 1 quarters (fewer slot collisions between the eight cells of a quarter)
  const int tx = (wrp & 3) * 8 + (lane & 1) + 2 * (lane >> 3), ty = (wrp >> 2) * 4 + ((lane >> 1) & 3);
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint32_t* c = tile + (ty + kHalo) * kPitch + tx + kHalo;

  if (OC == 1) {
    int n = 0;
    if (PAIRED) {  // two lookups per sorting network (production)
#define LERF_P1(M, RA, RB)                                                \
  {                                                                       \
    Simplex sa, sb;                                                       \
    simplex_pair_at<M, RA, RB>(c, tabs.h, sa, sb);                        \
    n += lookup1(tabs.t[STAGE == 1 ? M : 2 * M + (RA & 1)], sa);          \
    n += lookup1(tabs.t[STAGE == 1 ? M : 2 * M + (RB & 1)], sb);          \
  }
      LERF_P1(0, 0, 1) LERF_P1(0, 2, 3) LERF_P1(1, 0, 1) LERF_P1(1, 2, 3) LERF_P1(2, 0, 1) LERF_P1(2, 2, 3)
#undef LERF_P1
    } else {
#define LERF_L1(M, R) n += lookup1(tabs.t[STAGE == 1 ? M : 2 * M + (R & 1)], simplex_at<M, R>(c, tabs.h));
      LERF_L1(0, 0) LERF_L1(0, 1) LERF_L1(0, 2) LERF_L1(0, 3)
      LERF_L1(1, 0) LERF_L1(1, 1) LERF_L1(1, 2) LERF_L1(1, 3)
      LERF_L1(2, 0) LERF_L1(2, 1) LERF_L1(2, 2) LERF_L1(2, 3)
#undef LERF_L1
    }
    int v;
    if (STAGE == 1) {
      v = n <= 0 ? 0 : min(rhe_div(n, 48), 255);
    } else {
      const int t = n + 127 * 192;
      v = t <= 0 ? 0 : min(rhe_div(t, 192), 255);
    }
    __stcg(out + ((long long)p * H + y) * W + x, (uint8_t)v);
  } else {
    int n0 = 0, n1 = 0, n2 = 0;
#define LERF_L3(M, R) lookup3(tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c, tabs.h), n0, n1, n2);
    LERF_L3(0, 0) LERF_L3(0, 1) LERF_L3(0, 2) LERF_L3(0, 3)
    LERF_L3(1, 0) LERF_L3(1, 1) LERF_L3(1, 2) LERF_L3(1, 3)
    LERF_L3(2, 0) LERF_L3(2, 1) LERF_L3(2, 2) LERF_L3(2, 3)
#undef LERF_L3
    const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
    const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
    __stcg(out + o, (uint8_t)(t0 <= 0 ? 0 : min(rhe_div(t0, 192), 255)));
    __stcg(out + o + ps, (uint8_t)(t1 <= 0 ? 0 : min(rhe_div(t1, 192), 255)));
    __stcg(out + o + 2 * ps, (uint8_t)(t2 <= 0 ? 0 : min(rhe_div(t2, 192), 255)));
  }
}

}  // namespace cellk
}  // namespace lerf

// Row-major-table LUT stage (first implementation; production path of stage 2 for oC = 3): device body shared by
// the plain kernel in lut.cu and the pipeline kernel in pipeline.cu.
//
// Reference being replaced (ddlee-cn/LeRF-PyTorch): FourSimplexInterpFaster resample/eval_lut_sr.py:24-470 and the
// ensembling loops :541-628.  One thread = one sample; rotations are clamped constant offsets into a shared-memory
// tile; the 24-way branch is a 5-compare-exchange sort of (lsb<<13 | stride) keys.
#pragma once
#include "common.cuh"

namespace lerf {
namespace rm {

// ---------------------------------------------------------------------------------------------
// simplex walk
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap_desc(int& a, int& b) {
  const int hi = max(a, b);
  b = min(a, b);
  a = hi;
}

// Sorted-simplex form of the 24 cases of eval_lut_sr.py:218-462: sort taps by LSB,
// largest first; vertex k adds the stride of the k-th sorted tap; weights are the LSB gaps.  Ties
// have zero weight, so their order is irrelevant.
struct Simplex {
  int i0, i1, i2, i3, i4;  // table row of p0000 .. p1111
  int w0, w1, w2, w3, w4;  // 16-f1, f1-f2, f2-f3, f3-f4, f4   (sum = 16)
};

__device__ __forceinline__ Simplex simplex_of(int va, int vb, int vc, int vd) {
  Simplex s;
  s.i0 = (((va >> 4) * kL + (vb >> 4)) * kL + (vc >> 4)) * kL + (vd >> 4);
  int ka = ((va & 15) << 13) | kStrideA;
  int kb = ((vb & 15) << 13) | kStrideB;
  int kc = ((vc & 15) << 13) | kStrideC;
  int kd = ((vd & 15) << 13) | 1;
  cswap_desc(ka, kb);
  cswap_desc(kc, kd);
  cswap_desc(ka, kc);
  cswap_desc(kb, kd);
  cswap_desc(kb, kc);
  const int f1 = ka >> 13, f2 = kb >> 13, f3 = kc >> 13, f4 = kd >> 13;
  s.i1 = s.i0 + (ka & 8191);
  s.i2 = s.i1 + (kb & 8191);
  s.i3 = s.i2 + (kc & 8191);
  s.i4 = s.i0 + kStrideAll;
  s.w0 = 16 - f1;
  s.w1 = f1 - f2;
  s.w2 = f2 - f3;
  s.w3 = f3 - f4;
  s.w4 = f4;
  return s;
}

__device__ __forceinline__ int blend1(const int8_t* __restrict__ t, const Simplex& s) {
  return s.w0 * (int)__ldg(t + s.i0) + s.w1 * (int)__ldg(t + s.i1) + s.w2 * (int)__ldg(t + s.i2) +
         s.w3 * (int)__ldg(t + s.i3) + s.w4 * (int)__ldg(t + s.i4);
}

// Roofline experiments only (wrong results by construction): EXP 1 = every lane reads row 0 (loads issued, no
// address divergence); EXP 2 = no table load at all (the index stands in for the value).
template <int EXP>
__device__ __forceinline__ int blend1x(const int8_t* __restrict__ t, const Simplex& s) {
  if (EXP == 1)
    return s.w0 * (int)__ldg(t + (s.i0 & 1)) + s.w1 * (int)__ldg(t + (s.i1 & 1)) + s.w2 * (int)__ldg(t + (s.i2 & 1)) +
           s.w3 * (int)__ldg(t + (s.i3 & 1)) + s.w4 * (int)__ldg(t + (s.i4 & 1));
  return s.w0 * s.i0 + s.w1 * s.i1 + s.w2 * s.i2 + s.w3 * s.i3 + s.w4 * s.i4;
}

// oC = 3 tables are repacked to one uint32 per row, bytes (c0, c1, c2, 0): one 32-bit load per
// vertex and three dp4a per vertex with the weight placed in the byte lane of the wanted channel.
__device__ __forceinline__ void blend3(const uint32_t* __restrict__ t, const Simplex& s, int& n0,
                                       int& n1, int& n2) {
  const int e0 = (int)__ldg(t + s.i0), e1 = (int)__ldg(t + s.i1), e2 = (int)__ldg(t + s.i2),
            e3 = (int)__ldg(t + s.i3), e4 = (int)__ldg(t + s.i4);
#define LERF_ACC3(e, w)              \
  n0 = __dp4a(e, (w), n0);           \
  n1 = __dp4a(e, (w) << 8, n1);      \
  n2 = __dp4a(e, (w) << 16, n2);
  LERF_ACC3(e0, s.w0)
  LERF_ACC3(e1, s.w1)
  LERF_ACC3(e2, s.w2)
  LERF_ACC3(e3, s.w3)
  LERF_ACC3(e4, s.w4)
#undef LERF_ACC3
}

// ---------------------------------------------------------------------------------------------
// tap geometry: mode pattern (eval_lut_sr.py:30-81) composed with the rotation
// ---------------------------------------------------------------------------------------------
// MODE 0 = 's', 1 = 'c', 2 = 't'.  (di, dj) is the tap offset in the rotated frame; rotating the
// image by r quarter turns, edge-padding bottom/right and un-rotating the result is the same as
// reading the un-rotated image at the offsets below, clamped to the image.
template <int MODE, int R, int K>
struct Tap {
  static constexpr int di = MODE == 0 ? (K >> 1) : (MODE == 1 ? 0 : K);
  static constexpr int dj = MODE == 0 ? (K & 1) : K;
  static constexpr int dy = R == 0 ? di : (R == 1 ? dj : (R == 2 ? -di : -dj));
  static constexpr int dx = R == 0 ? dj : (R == 1 ? -di : (R == 2 ? -dj : di));
};

constexpr int kHalo = 3;  // reach of modes c and t
constexpr int kTX = 32, kTY = 8;
constexpr int kPitch = kTX + 2 * kHalo + 2;  // 40: rows of the tile, bytes

template <int MODE, int R>
__device__ __forceinline__ Simplex simplex_at(const uint8_t* c) {
  const int va = c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx];
  const int vb = c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx];
  const int vc = c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx];
  const int vd = c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx];
  return simplex_of(va, vb, vc, vd);
}

struct StageTables {
  const void* t[6];
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// round_half_even(num / den) for num > 0, den even
__device__ __forceinline__ int rhe_div(int num, int den) {
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;  // exact .5 -> even
  return q;
}

// One rotation-ensembled stage.  STAGE 1: 3 tables (s,c,t) used for all four rotations, output
// feat = clip(rhe(sum/48)).  STAGE 2: 6 tables ([mode][r&1]), oC outputs, code = clip(rhe(sum/192 + 127)).
// EXP: 0 = real; 1, 2 = roofline experiments (see blend1x).  `tile` = kTileBytes of shared memory; (bxi, byi, p) = the
// block's tile column, tile row and plane (blockIdx of the plain launch, or a flattened role block of the pipeline
// kernel in pipeline.cu).
constexpr int kTileBytes = (kTY + 2 * kHalo) * kPitch;

template <int STAGE, int OC, int EXP>
__device__ __forceinline__ void lut_stage_body(const StageTables& tabs, const uint8_t* __restrict__ in, const InAddr& ia,
                                               int H, int W, int y0, int y1, uint8_t* __restrict__ out, int bxi, int byi,
                                               int p, uint8_t* tile) {
  const int bx = bxi * kTX, by = y0 + byi * kTY;
  const uint8_t* src = in + (long long)(p / ia.channels) * ia.batch_stride +
                       (long long)(p % ia.channels) * ia.chan_stride;
  const int tid = threadIdx.x;
  for (int i = tid; i < (kTY + 2 * kHalo) * (kTX + 2 * kHalo); i += kTX * kTY) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = clampi(by + r - kHalo, 0, H - 1), gx = clampi(bx + c - kHalo, 0, W - 1);
    tile[r * kPitch + c] = __ldcg(src + (long long)gy * ia.row_stride + (long long)gx * ia.pix_stride);
  }
  __syncthreads();
  // A warp covers an 8x4 pixel patch, not a 32x1 row: neighbours in 2-D have closer values than the ends of a
  // 32-pixel row, so the 32 table addresses of a gather fall into fewer cache lines (measured).
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);
  const int x = bx + tx, y = by + ty;
  if (x >= W || y >= y1) return;
  const uint8_t* c = tile + (ty + kHalo) * kPitch + tx + kHalo;

  if (STAGE == 1) {
    int n = 0;
#define LERF_S1(M, R)                                                                  \
  n += (EXP ? blend1x<EXP>((const int8_t*)tabs.t[M], simplex_at<M, R>(c)) \
            : blend1((const int8_t*)tabs.t[M], simplex_at<M, R>(c)));
    LERF_S1(0, 0) LERF_S1(0, 1) LERF_S1(0, 2) LERF_S1(0, 3)
    LERF_S1(1, 0) LERF_S1(1, 1) LERF_S1(1, 2) LERF_S1(1, 3)
    LERF_S1(2, 0) LERF_S1(2, 1) LERF_S1(2, 2) LERF_S1(2, 3)
#undef LERF_S1
    const int v = n <= 0 ? 0 : min(rhe_div(n, 48), 255);
    __stcg(out + ((long long)p * H + y) * W + x, (uint8_t)v);
  } else if (OC == 1) {
    int n = 0;
#define LERF_S2(M, R) n += blend1((const int8_t*)tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c));
    LERF_S2(0, 0) LERF_S2(0, 1) LERF_S2(0, 2) LERF_S2(0, 3)
    LERF_S2(1, 0) LERF_S2(1, 1) LERF_S2(1, 2) LERF_S2(1, 3)
    LERF_S2(2, 0) LERF_S2(2, 1) LERF_S2(2, 2) LERF_S2(2, 3)
#undef LERF_S2
    const int t = n + 127 * 192;
    __stcg(out + ((long long)p * H + y) * W + x, (uint8_t)(t <= 0 ? 0 : min(rhe_div(t, 192), 255)));
  } else {
    int n0 = 0, n1 = 0, n2 = 0;
#define LERF_S2(M, R) blend3((const uint32_t*)tabs.t[2 * M + (R & 1)], simplex_at<M, R>(c), n0, n1, n2);
    LERF_S2(0, 0) LERF_S2(0, 1) LERF_S2(0, 2) LERF_S2(0, 3)
    LERF_S2(1, 0) LERF_S2(1, 1) LERF_S2(1, 2) LERF_S2(1, 3)
    LERF_S2(2, 0) LERF_S2(2, 1) LERF_S2(2, 2) LERF_S2(2, 3)
#undef LERF_S2
    const long long o = ((long long)p * 3 * H + y) * W + x, ps = (long long)H * W;
    const int t0 = n0 + 127 * 192, t1 = n1 + 127 * 192, t2 = n2 + 127 * 192;
    __stcg(out + o, (uint8_t)(t0 <= 0 ? 0 : min(rhe_div(t0, 192), 255)));
    __stcg(out + o + ps, (uint8_t)(t1 <= 0 ? 0 : min(rhe_div(t1, 192), 255)));
    __stcg(out + o + 2 * ps, (uint8_t)(t2 <= 0 ? 0 : min(rhe_div(t2, 192), 255)));
  }
}

}  // namespace rm
}  // namespace lerf

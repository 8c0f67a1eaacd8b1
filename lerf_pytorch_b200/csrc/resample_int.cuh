// Integer-scale specialisation of the LeRF-G SR resampler (S = 2, 3, 4, 8 on both axes, out = S * in): device body
// shared by the plain kernel in resample_int.cu and the pipeline kernel in pipeline.cu.
// Replaces SteeringGaussianResize2dNumpy.resize (resize_right/resize_right2d_numpy.py:162-223 of the reference)
// for the configurations where the geometry is periodic (BASELINE.json cfg-1, -3, -5).
//
// Design.  For an integer scale the S x S output pixels whose first tap is input pixel (ly, lx) share the same
// 2x2 taps, so one thread owns one such CELL: it reads the four taps' coefficients once and produces S*S outputs.
// Per tap pixel and plane the exponent is a quadratic form
//     log2(w) = a' dr^2 + b' dr dc + c' dc^2,   a' = -L/2 sx^2,  b' = L rho sx sy,  c' = -L/2 sy^2,  L = log2(e)
// whose coefficients are computed once per input sample in float64 from the reference's float32 hyper values (a
// block stages them in shared memory), and whose geometry factors are float64 kernel-parameter constants.
// HOIST variant: the column term c' dc^2 (+ the rounding constant) is hoisted per cell and the row terms a' dr^2,
// b' dr per output row, so an exponent costs 2 FP64 ops instead of 4 -- but 80+ registers; measured no faster
// than the plain form at 48 registers, which is the default.  Adding 1.5*2^(52-FB) leaves round(log2(w) * 2^FB) in the low word
// of the double: the max over the four taps and the subtraction are then exact INTEGER ops, and only the difference
// (<= 0) is converted to fp32 for ex2.approx -- the fp32 error stays in the low bits of weights that matter.  The
// output is v00 + sum w_t (v_t - v00) / sum w_t with exact integer differences, so fp32 rounding scales with the
// local contrast, not with 255.
#pragma once
#include "common.cuh"

namespace lerf {
namespace rsi {

constexpr double kLog2e = 1.4426950408889634;
// The fixed-point exponent of the fast resamplers is rounded (up to three times) at 2^-FB: with FB = 23 that is 1.2e-7 of a
// weight, 3e-5 of an output sample in the worst case, on top of ex2.approx's 6e-5 -- inside the 1e-4 bar.  A plan whose
// exponent range (max_sigma, distances) leaves fewer bits goes to the next kernel down, in the end the float64 one.
constexpr int kMinFracBits = 23;

template <int S>
struct IntGeom {
  double xr[S][2];        // dr^2       [row phase][tap b]
  double xc[S][2];        // dc^2       [col phase][tap a]
  double dr[S][2];        // dr
  double dc[S][2];        // dc
  double pp[S][S][2][2];  // dr * dc    [row phase][col phase][b][a]   (S = 8 form)
  double magic;           // 1.5 * 2^(52 - FB)
  float inv_scale;        // 2^-FB
  int ph_y, ph_x;         // first output of cell l is S*l + ph
  int fb;                 // fraction bits of the fixed-point exponent (host only: the launchers require >= kMinFracBits)
};

// Compile-time geometry of out = S * in for even S (r2): the distances of output S*l + S/2 + m to its two taps are
// (2m + 1) / (2S) and that minus 1 -- dyadic numbers whose squares and products have an all-zero low word, so every DFMA of
// the ROWQ form takes its geometry factor as an IMMEDIATE instead of a uniform register that first has to be loaded from
// the constant bank (71 LDCU per cell of 16 outputs in the r1 kernel, 7 % of its instructions).  Values are NEGATED like
// make_geom's unsigned form.  The launcher checks the plan's float64 tables against d() before it picks this flavour.
template <int S>
struct CGeom {
  __host__ __device__ static constexpr double d(int m, int k) { return (2.0 * m + 1.0) / (2.0 * S) - k; }
  __host__ __device__ static constexpr double xq(int m, int k) { return -(d(m, k) * d(m, k)); }
  __host__ __device__ static constexpr double pp(int mr, int mc, int b, int a) { return -(d(mr, b) * d(mc, a)); }
};
template <int S, bool C>
__device__ __forceinline__ double geom_xr(const IntGeom<S>& g, int m, int k) {
  if constexpr (C) return CGeom<S>::xq(m, k);
  else return g.xr[m][k];
}
template <int S, bool C>
__device__ __forceinline__ double geom_xc(const IntGeom<S>& g, int m, int k) {
  if constexpr (C) return CGeom<S>::xq(m, k);
  else return g.xc[m][k];
}
template <int S, bool C>
__device__ __forceinline__ double geom_pp(const IntGeom<S>& g, int mr, int mc, int b, int a) {
  if constexpr (C) return CGeom<S>::pp(mr, mc, b, a);
  else return g.pp[mr][mc][b][a];
}

// Per-code float64 tables, exact promotions of the reference's float32 hyper values; built once per launch on the
// host (IEEE float32 arithmetic, same operation order as numpy) per max_sigma and kept in device memory by the plan.
struct CoefTabs {
  double s2[256];  // -L/2 * sigma^2
  double sg[256];  // sigma
  double rl[256];  // L * rho
};

constexpr int kCX = 32, kCY = 8;  // cells per block

struct Smem {
  CoefTabs tab;
  double sA[kCY + 1][kCX + 1], sB[kCY + 1][kCX + 1], sC[kCY + 1][kCX + 1];
  float sV[kCY + 1][kCX + 1];
};

template <int FMT>
__device__ __forceinline__ void store1(void* out, long long ip, long long ih, float val) {
  if (FMT == LERF_OUT_F32) {
    __stcg((float*)out + ip, val);
  } else {
    int q = __float2int_rn(val);  // round half to even
    q = min(max(q, 0), 255);
    __stcg((uint8_t*)out + (FMT == LERF_OUT_U8 ? ip : ih), (uint8_t)q);
  }
}

// 4 fixed-point exponents (round(log2 w * 2^FB)) and the 4 tap differences -> one output sample.
// (Building 2^x without the int->float conversion -- mantissa bits + exponent-field add -- was measured SLOWER: the
// kernel is issue-bound and that form needs one more instruction per tap)
__device__ __forceinline__ float combine_q(const int q[4], const float dv[4], float v0, float inv_scale) {
  const int qm = max(max(q[0], q[1]), max(q[2], q[3]));
  float w[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float x = (float)(q[t] - qm) * inv_scale;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(x));
  }
  const float den = (w[0] + w[1]) + (w[2] + w[3]);  // in [1, 4]: the max tap has weight 1
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);  // residual correction with the approximate reciprocal: ~0.5 ulp quotient
  return v0 + qn;
}

// Unsigned form (ROWQ variant): q[t] = round(-log2 w * 2^FB) + 16 >= 0 as uint32 (the form is positive semi-definite, so
// -log2 w >= 0 and the whole 32-bit range carries magnitude: one more fraction bit than the signed form).  The tap
// with the SMALLEST q has weight 1; neg_scale = -2^-FB.
__device__ __forceinline__ float combine_uq(const unsigned q[4], const float dv[4], float v0, float neg_scale) {
  const unsigned qm = min(min(q[0], q[1]), min(q[2], q[3]));
  float w[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float x = __uint2float_rn(q[t] - qm) * neg_scale;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(x));
  }
  const float den = (w[0] + w[1]) + (w[2] + w[3]);
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);
  return v0 + qn;
}

// combine_uq with a FIXED reference tap R instead of the smallest exponent (r2): for a compile-time phase the nearest tap is
// known, its distances are <= 1/2 per axis, so its -log2 w is at most L/2 (sigma (1/2 + 1/2))^2 = 72 for sigma = 10 and the
// other weights, taken relative to it, stay below 2^72 -- far inside float32 -- while the reference weight is exactly 1: no
// minimum (two VIMNMX3) and one tap's subtract / convert / scale / ex2 less per output sample.  The differences are SIGNED
// now, so the plan's fixed point keeps one bit less (make_geom signed_diff: bound * 2^FB < 2^31).
constexpr bool kRefTapDefault = true;  // production of the integer-scale kernels since r2h: 258 -> 234 us per 2K frame (float32)

template <int R>
__device__ __forceinline__ void weights_ref(const unsigned q[4], float neg_scale, float w[4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t == R) {
      w[t] = 1.0f;
    } else {
      const float x = __int2float_rn((int)(q[t] - q[R])) * neg_scale;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(x));
    }
  }
}
template <int R>
__device__ __forceinline__ float combine_ref(const unsigned q[4], const float dv[4], float v0, float neg_scale) {
  float w[4];
  weights_ref<R>(q, neg_scale, w);
  const float den = (w[0] + w[1]) + (w[2] + w[3]);
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);
  return v0 + qn;
}

template <int R>  // uint8 flavour, see combine_uq_u8
__device__ __forceinline__ uint32_t combine_ref_u8(const unsigned q[4], const float dv[4], float v0m, float neg_scale) {
  float w[4];
  weights_ref<R>(q, neg_scale, w);
  const float den = (w[0] + w[1]) + (w[2] + w[3]);
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);
  return __float_as_uint(v0m + qn);
}

// 1.5 * 2^23: adding it to a value in [0, 2^22) leaves round_half_even(value) in the low mantissa bits.
constexpr float kRoundMagic = 12582912.0f;

// combine_uq for the uint8 formats: returns the BITS of (v0 + 1.5*2^23) + qn.  v0 is an integer <= 255, so the first sum is
// exact and the second rounds the exact v0 + qn to the integer grid, half to even -- the reference's np.round
// (eval_lut_sr.py:663-665) -- in ONE rounding, for free (it replaces the final add).  The weights are non-negative and
// normalised, so v0 + qn lies in [0, 255] up to rounding error far below 0.5 and the low BYTE is the clipped result.
__device__ __forceinline__ uint32_t combine_uq_u8(const unsigned q[4], const float dv[4], float v0m, float neg_scale) {
  const unsigned qm = min(min(q[0], q[1]), min(q[2], q[3]));
  float w[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float x = __uint2float_rn(q[t] - qm) * neg_scale;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(x));
  }
  const float den = (w[0] + w[1]) + (w[2] + w[3]);
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);
  return __float_as_uint(v0m + qn);
}

// Streaming traffic (codes, feat, outputs) uses .cg loads/stores: it must not evict the LUT lines that the stage
// roles of the pipeline kernel keep in L1.
// (bxi, byi, p): the block's cell-tile column, cell-tile row and plane; 256 threads.
// MODE: 0 = plain (4 FP64 per exponent), 1 = fully hoisted signed form, 2 = ROWQ: unsigned fixed point with the row
// term and the rounding constant hoisted per output row (2 FP64 per exponent, 8 more registers than MODE 0).
template <int S, int FMT, int MODE>
__device__ __forceinline__ void resize_int_body(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H,
                                                int W, int oH, int oW, const IntGeom<S>& g, const CoefTabs* __restrict__ ct, int channels,
                                                int ly0, int oy0, int oy1, void* __restrict__ out, int bxi, int byi, int p,
                                                Smem& sm) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  sm.tab.s2[tid] = __ldg(ct->s2 + tid);  // global (L1/L2-resident, coalesced) -> shared memory; the tile fill indexes by code
  sm.tab.sg[tid] = __ldg(ct->sg + tid);
  sm.tab.rl[tid] = __ldg(ct->rl + tid);
  __syncthreads();
  const int lx0 = bxi * kCX - 1;    // first cell column of the block (cells start at -1)
  const int lyb = ly0 + byi * kCY;  // first cell row of the block
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  const uint8_t* cp = codes + (long long)p * 3 * plane_sz;
  for (int i = tid; i < (kCY + 1) * (kCX + 1); i += kCX * kCY) {
    const int r = i / (kCX + 1), c = i - r * (kCX + 1);
    const int sy = lyb + r, sx = lx0 + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'edge' (:172-174)
    const long long off = (long long)cy * W + cx;
    const int kr = __ldcg(cp + off), kx = __ldcg(cp + plane_sz + off), ky = __ldcg(cp + 2 * plane_sz + off);
    sm.sA[r][c] = sm.tab.s2[kx];
    sm.sC[r][c] = sm.tab.s2[ky];
    sm.sB[r][c] = sm.tab.rl[kr] * sm.tab.sg[kx] * sm.tab.sg[ky];
    sm.sV[r][c] = (sy == cy && sx == cx) ? (float)__ldcg(fp + off) : 0.0f;  // image: 'constant' 0 (:208)
  }
  __syncthreads();
  const int lx = lx0 + tx, ly = lyb + ty;
  if (lx > W - 1 || ly > H - 1) return;
  // taps t = a*2+b: row ly+b, column lx+a (same patch order as the reference, :95-98)
  double ca[4], cb[4], cc[4];
  float dv[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      ca[a * 2 + b] = sm.sA[ty + b][tx + a];
      cb[a * 2 + b] = sm.sB[ty + b][tx + a];
      cc[a * 2 + b] = sm.sC[ty + b][tx + a];
      dv[a * 2 + b] = sm.sV[ty + b][tx + a];
    }
  const float v0 = dv[0];
  dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;  // exact: integers in [-255, 255]
  const int oyb = S * ly + g.ph_y, oxb = S * lx + g.ph_x;
  const int pc_ = p % channels;
  long long rowp = ((long long)p * oH + oyb) * oW;                  // planar index of (p, oyb, 0); advanced by oW per row
  long long rowh = ((long long)(p / channels) * oH + oyb) * oW;     // same for the interleaved layout
  const bool full = oxb >= 0 && oxb + S <= oW;
  constexpr bool kHoist = MODE == 1 && S <= 4;
  constexpr bool kRowQ = MODE == 2 || MODE == 3 || MODE == 5 || MODE == 6;  // S = 8 would need 64 registers for the column terms
  constexpr bool kCG = MODE == 3 || MODE == 5;                              // geometry factors as immediates (CGeom)
  constexpr bool kRef = MODE == 5 || MODE == 6;                             // weights relative to the phase's nearest tap (combine_ref)
  double colq[kHoist ? 4 : 1][kHoist ? S : 1];
  if (kHoist) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int mc = 0; mc < S; ++mc) colq[t][mc] = fma(cc[t], g.xc[mc][t >> 1], g.magic);
  }
#pragma unroll
  for (int mr = 0; mr < S; ++mr, rowp += oW, rowh += oW) {
    const int oy = oyb + mr;
    if (oy < oy0 || oy >= oy1) continue;
    double rowa[4], rowb[4];
    if (kHoist) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        rowa[t] = ca[t] * g.xr[mr][t & 1];
        rowb[t] = cb[t] * g.dr[mr][t & 1];
      }
    }
    if (kRowQ) {  // geometry constants are NEGATED in this mode (make_geom): rowa = magic - a' dr^2 >= magic
#pragma unroll
      for (int t = 0; t < 4; ++t) rowa[t] = fma(ca[t], geom_xr<S, kCG>(g, mr, t & 1), g.magic);
    }
    float res[S];
#pragma unroll
    for (int mc = 0; mc < S; ++mc) {
      int q[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int a = t >> 1, b = t & 1;
        if (kRowQ) {
          // order matters: the two non-negative terms first, the signed cross term last, so no partial sum drops
          // below the binade of `magic` (which carries a +16-unit guard for the two extra roundings)
          double e = fma(cc[t], geom_xc<S, kCG>(g, mc, a), rowa[t]);
          e = fma(cb[t], geom_pp<S, kCG>(g, mr, mc, b, a), e);
          q[t] = __double2loint(e);
        } else if (kHoist) {
          const double e = fma(rowb[t], g.dc[mc][a], rowa[t]);
          q[t] = __double2loint(e + colq[t][mc]);  // round(log2 w * 2^FB), two's complement
        } else {
          double e = cc[t] * g.xc[mc][a];
          e = fma(cb[t], g.pp[mr][mc][b][a], e);
          e = fma(ca[t], g.xr[mr][b], e);
          q[t] = __double2loint(e + g.magic);
        }
      }
      if (kRowQ) {
        const unsigned uq[4] = {(unsigned)q[0], (unsigned)q[1], (unsigned)q[2], (unsigned)q[3]};
        if (FMT == LERF_OUT_F32 && kRef) {  // centred phases: tap (a, b) = (2 mc + 1 >= S, 2 mr + 1 >= S) is the nearest one (ref_tap_ok)
          if (mc * 2 + 1 >= S) res[mc] = mr * 2 + 1 >= S ? combine_ref<3>(uq, dv, v0, g.inv_scale) : combine_ref<2>(uq, dv, v0, g.inv_scale);
          else res[mc] = mr * 2 + 1 >= S ? combine_ref<1>(uq, dv, v0, g.inv_scale) : combine_ref<0>(uq, dv, v0, g.inv_scale);
        } else if (kRef) {  // uint8 formats that do not take a staged kernel (x8 interleaved): the same weights
          const float v0m = v0 + kRoundMagic;
          uint32_t bits;
          if (mc * 2 + 1 >= S) bits = mr * 2 + 1 >= S ? combine_ref_u8<3>(uq, dv, v0m, g.inv_scale) : combine_ref_u8<2>(uq, dv, v0m, g.inv_scale);
          else bits = mr * 2 + 1 >= S ? combine_ref_u8<1>(uq, dv, v0m, g.inv_scale) : combine_ref_u8<0>(uq, dv, v0m, g.inv_scale);
          res[mc] = (float)(bits & 255u);
        } else if (FMT == LERF_OUT_F32) res[mc] = combine_uq(uq, dv, v0, g.inv_scale);
        else res[mc] = (float)(combine_uq_u8(uq, dv, v0 + kRoundMagic, g.inv_scale) & 255u);  // one rounding, like every uint8 epilogue (r2)
      } else {
        res[mc] = combine_q(q, dv, v0, g.inv_scale);
      }
    }
    if (FMT == LERF_OUT_F32 && full && (S % 2 == 0)) {
      float* o = (float*)out + rowp + oxb;
      if (S == 8) {  // ph = 4: 16-byte aligned
        __stcg(reinterpret_cast<float4*>(o), make_float4(res[0], res[1], res[2], res[3]));
        __stcg(reinterpret_cast<float4*>(o + 4), make_float4(res[4 % S], res[5 % S], res[6 % S], res[7 % S]));
      } else if (S == 4) {  // ph = 2: 8-byte aligned
        __stcg(reinterpret_cast<float2*>(o), make_float2(res[0], res[1]));
        __stcg(reinterpret_cast<float2*>(o + 2), make_float2(res[2 % S], res[3 % S]));
      } else {
        __stcg(o, res[0]);
        __stcg(o + 1, res[1 % S]);
      }
    } else {
#pragma unroll
      for (int mc = 0; mc < S; ++mc) {
        const int ox = oxb + mc;
        if (ox >= 0 && ox < oW) store1<FMT>(out, rowp + ox, (rowh + ox) * channels + pc_, res[mc]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Staged uint8 output (r2).  The byte-per-sample stores of the first version made the uint8 formats SLOWER than float32
// (301 / 350 us against 268 per 2K frame) while writing a quarter of the bytes.  Here a block keeps its whole output
// tile in shared memory -- planar: packed words; interleaved HWC: the block walks the three colour planes itself and
// scatters bytes into pixel order -- and then copies it out row by row in 16-byte units aligned to GLOBAL memory
// (the tile starts at output column S*lx0 + ph, so the row segment is realigned on the way: five shared words, four
// byte-funnel PRMTs, one 128-bit store).  Only the ragged first / last unit of a row segment goes out byte by byte.
// ---------------------------------------------------------------------------------------------------------------
template <int S, int CH>
struct OutTile {
  static constexpr int kRows = kCY * S;
  static constexpr int kRowBytes = kCX * S * CH;
  static constexpr int kSlack = 16;                        // bytes before the row data (the realigning copy reads around it)
  static constexpr int kPitch = kRowBytes + 2 * kSlack;    // multiple of 16
  alignas(16) unsigned char b[kRows][kPitch];
};

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // low bytes of a, b, c, d
  uint32_t ab, cd, r;
  asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(ab) : "r"(a), "r"(b));
  asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(cd) : "r"(c), "r"(d));
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(ab), "r"(cd));
  return r;
}

__device__ __forceinline__ uint32_t to_u8_sat(float v) {  // clip(round_half_even(v), 0, 255) (eval_lut_sr.py:663-665); NaN -> 0
  uint32_t r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// The ROWQ arithmetic of resize_int_body for one cell: `emit(mr, res)` receives the S samples of output row oyb + mr.
// (uint8 flavour: res[mc] carries the rounded sample in its low byte, see combine_uq_u8.)
// CG: bit 0 = geometry factors as immediates (CGeom) instead of kernel parameters, bit 1 = weights relative to the phase's
// nearest tap (combine_ref_u8) instead of the smallest exponent.
template <int S, int CG, typename Emit>
__device__ __forceinline__ void gauss_cell_rowq(const Smem& sm, const IntGeom<S>& g, int tx, int ty, int oyb, int oy0, int oy1,
                                                Emit emit) {
  double ca[4], cb[4], cc[4];
  float dv[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      ca[a * 2 + b] = sm.sA[ty + b][tx + a];
      cb[a * 2 + b] = sm.sB[ty + b][tx + a];
      cc[a * 2 + b] = sm.sC[ty + b][tx + a];
      dv[a * 2 + b] = sm.sV[ty + b][tx + a];
    }
  const float v0 = dv[0];
  dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;
  const float v0m = v0 + kRoundMagic;
#pragma unroll
  for (int mr = 0; mr < S; ++mr) {
    const int oy = oyb + mr;
    if (oy < oy0 || oy >= oy1) continue;
    double rowa[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) rowa[t] = fma(ca[t], geom_xr<S, (CG & 1) != 0>(g, mr, t & 1), g.magic);
    uint32_t res[S];
#pragma unroll
    for (int mc = 0; mc < S; ++mc) {
      unsigned uq[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        double e = fma(cc[t], geom_xc<S, (CG & 1) != 0>(g, mc, t >> 1), rowa[t]);
        e = fma(cb[t], geom_pp<S, (CG & 1) != 0>(g, mr, mc, t & 1, t >> 1), e);
        uq[t] = (unsigned)__double2loint(e);
      }
      if (CG & 2) {  // centred phases: tap (a, b) = (2 mc + 1 >= S, 2 mr + 1 >= S) is the nearest one (ref_tap_ok)
        if (mc * 2 + 1 >= S) res[mc] = mr * 2 + 1 >= S ? combine_ref_u8<3>(uq, dv, v0m, g.inv_scale) : combine_ref_u8<2>(uq, dv, v0m, g.inv_scale);
        else res[mc] = mr * 2 + 1 >= S ? combine_ref_u8<1>(uq, dv, v0m, g.inv_scale) : combine_ref_u8<0>(uq, dv, v0m, g.inv_scale);
      } else {
        res[mc] = combine_uq_u8(uq, dv, v0m, g.inv_scale);
      }
    }
    emit(mr, res);
  }
}

// Fills the coefficient tiles of plane p (hypers 'edge'-replicated, image zero-padded); the caller synchronises.
__device__ __forceinline__ void stage_coefs(Smem& sm, const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H,
                                            int W, int p, int lx0, int lyb, int tid) {
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  const uint8_t* cp = codes + (long long)p * 3 * plane_sz;
  for (int i = tid; i < (kCY + 1) * (kCX + 1); i += kCX * kCY) {
    const int r = i / (kCX + 1), c = i - r * (kCX + 1);
    const int sy = lyb + r, sx = lx0 + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);
    const long long off = (long long)cy * W + cx;
    const int kr = __ldcg(cp + off), kx = __ldcg(cp + plane_sz + off), ky = __ldcg(cp + 2 * plane_sz + off);
    sm.sA[r][c] = sm.tab.s2[kx];
    sm.sC[r][c] = sm.tab.s2[ky];
    sm.sB[r][c] = sm.tab.rl[kr] * sm.tab.sg[kx] * sm.tab.sg[ky];
    sm.sV[r][c] = (sy == cy && sx == cx) ? (float)__ldcg(fp + off) : 0.0f;
  }
}

// Copies the valid part of a staged tile to global memory.  Row r of the tile is output row oyt + r; tile byte j of a
// row is global byte  g0(row) + j  with g0 = row_base + col0 (col0 may be negative: the tile starts left of the image);
// bytes [jlo, jhi) of every row are valid.
template <int S, int CH>
__device__ __forceinline__ void copy_tile_out(const OutTile<S, CH>& ot, unsigned char* __restrict__ out_base, long long row_pitch,
                                              long long col0, int oyt, int r_lo, int r_hi, int jlo, int jhi, int tid) {
  using OT = OutTile<S, CH>;
  constexpr int kUnits = OT::kRowBytes / 16 + 1;
  if (jlo >= jhi) return;
  // whole 16-byte units
  for (int idx = tid; idx < (r_hi - r_lo) * kUnits; idx += kCX * kCY) {
    const int r = r_lo + idx / kUnits, u = idx - (idx / kUnits) * kUnits;
    unsigned char* grow = out_base + (long long)(oyt + r) * row_pitch + col0;  // global address of tile byte 0
    const int mis = (int)((uintptr_t)(grow + jlo) & 15);
    const int j0 = jlo + ((16 - mis) & 15) + 16 * u;             // tile byte offset of this row's u-th whole unit
    if (j0 + 16 > jhi) continue;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&ot.b[r][0]) + ((OT::kSlack + j0) >> 2);
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)((OT::kSlack + j0) & 3);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    uint4 v;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.x) : "r"(w0), "r"(w1), "r"(sel));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.y) : "r"(w1), "r"(w2), "r"(sel));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.z) : "r"(w2), "r"(w3), "r"(sel));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v.w) : "r"(w3), "r"(w4), "r"(sel));
    __stcg(reinterpret_cast<uint4*>(grow + j0), v);
  }
  // ragged ends: the < 16 bytes before the first and after the last whole unit of each row, one lane per byte
  for (int idx = tid; idx < (r_hi - r_lo) * 32; idx += kCX * kCY) {
    const int r = r_lo + (idx >> 5), k = idx & 31;
    unsigned char* grow = out_base + (long long)(oyt + r) * row_pitch + col0;
    const int mis = (int)((uintptr_t)(grow + jlo) & 15);
    const int ja = min(jlo + ((16 - mis) & 15), jhi);            // head = [jlo, ja)
    const int jb = ja + ((jhi - ja) & ~15);                      // tail = [jb, jhi)
    const int j = k < 16 ? jlo + k : jb + (k - 16);
    if (k < 16 ? j < ja : j < jhi) __stcg(grow + j, ot.b[r][OT::kSlack + j]);
  }
}

// uint8 outputs of the integer-scale Gaussian resampler through a staged tile.  CH = 1: planar [P][oH][oW], one plane per
// block (blockIdx.z = plane).  CH = 3: interleaved [B][oH][oW][3], one image per block (blockIdx.z = batch index), the
// three colour planes one after the other.
template <int S, int CH, int CG, bool TMA = true>
__device__ __forceinline__ void resize_int_u8_body(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W,
                                                   int oH, int oW, const IntGeom<S>& g, const CoefTabs* __restrict__ ct, int ly0,
                                                   int oy0, int oy1, unsigned char* __restrict__ out, int bxi, int byi, int bz,
                                                   Smem& sm, OutTile<S, CH>& ot) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  sm.tab.s2[tid] = __ldg(ct->s2 + tid);
  sm.tab.sg[tid] = __ldg(ct->sg + tid);
  sm.tab.rl[tid] = __ldg(ct->rl + tid);
  const int lx0 = bxi * kCX - 1, lyb = ly0 + byi * kCY;
  const int lx = lx0 + tx, ly = lyb + ty;
  const bool active = lx <= W - 1 && ly <= H - 1;
  const int oyb = S * ly + g.ph_y;
  // valid part of the tile: rows inside the band and the image, columns inside the image
  const int oyt = S * lyb + g.ph_y;                       // output row of tile row 0
  const int ox0 = S * lx0 + g.ph_x;                       // output column of tile byte 0 (pixel units)
  const int r_lo = max(max(oy0, 0) - oyt, 0), r_hi = min(min(oy1, oH) - oyt, OutTile<S, CH>::kRows);
  const int jlo = max(0, -ox0) * CH, jhi = min(kCX * S, oW - ox0) * CH;
  unsigned char* base = out + (long long)bz * oH * oW * CH;  // planar: plane bz; HWC: image bz
  const long long row_pitch = (long long)oW * CH, col0 = (long long)ox0 * CH;
  // Bulk-store path (TMA engine, cp.async.bulk): when the output row pitch is a multiple of 16 bytes every row segment of
  // the tile has the SAME misalignment `mis` against 16 bytes.  The tile is then staged at byte offset `mis` of its shared
  // rows -- shared and global addresses congruent mod 16 -- so the aligned middle of every row goes out as ONE bulk copy
  // issued by one thread per row; only the ragged ends (< 16 bytes each) are stored by lanes.  Byte-granular staging only
  // (CH > 1): the packed-word staging of the planar form needs 4-byte aligned shared addresses.
  const bool bulk = TMA && CH > 1 && (row_pitch & 15) == 0;
  const int mis = bulk ? (int)((uintptr_t)(base + (long long)oyt * row_pitch + col0) & 15) : 0;
#pragma unroll 1
  for (int c = 0; c < CH; ++c) {
    __syncthreads();  // tables staged / the previous plane's coefficient tiles are no longer read
    stage_coefs(sm, feat, codes, H, W, bz * CH + c, lx0, lyb, tid);
    __syncthreads();
    if (active) {
      gauss_cell_rowq<S, CG>(sm, g, tx, ty, oyb, oy0, oy1, [&](int mr, const uint32_t* res) {
        unsigned char* row = &ot.b[ty * S + mr][OutTile<S, CH>::kSlack + mis];
        if (CH == 1 && S == 4) {
          *reinterpret_cast<uint32_t*>(row + 4 * tx) = pack4(res[0], res[1 % S], res[2 % S], res[3 % S]);
        } else {
#pragma unroll
          for (int mc = 0; mc < S; ++mc) row[(tx * S + mc) * CH + c] = (unsigned char)res[mc];
        }
      });
    }
  }
  if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (r_lo >= r_hi || jlo >= jhi) return;
  if (bulk) {
    const int ja = min(jlo + ((16 - ((mis + jlo) & 15)) & 15), jhi);  // head = [jlo, ja)
    const int nb = (jhi - ja) & ~15;                                  // bytes of the aligned middle
    if (tid < r_hi - r_lo && nb > 0) {
      const int r = r_lo + tid;
      unsigned char* gdst = base + (long long)(oyt + r) * row_pitch + col0 + ja;
      const uint32_t ssrc = (uint32_t)__cvta_generic_to_shared(&ot.b[r][OutTile<S, CH>::kSlack + mis + ja]);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(nb) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // ragged ends, one lane per byte
    for (int idx = tid; idx < (r_hi - r_lo) * 32; idx += kCX * kCY) {
      const int r = r_lo + (idx >> 5), k = idx & 31;
      unsigned char* grow = base + (long long)(oyt + r) * row_pitch + col0;
      const int jb = ja + nb;
      const int j = k < 16 ? jlo + k : jb + (k - 16);
      if (k < 16 ? j < ja : j < jhi) __stcg(grow + j, ot.b[r][OutTile<S, CH>::kSlack + mis + j]);
    }
    if (tid < r_hi - r_lo && nb > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile is read before the block retires
    return;
  }
  copy_tile_out<S, CH>(ot, base, row_pitch, col0, oyt, r_lo, r_hi, jlo, jhi, tid);
}

// Planar uint8 output for S = 4 and S = 8 without a staged tile.  A x4 cell owns output columns 4*lx + 2 .. 4*lx + 5: its four
// bytes straddle an aligned word, so every lane takes the two leading bytes of its right-hand neighbour (one shuffle,
// one PRMT) and stores the ALIGNED word 4*lx + 4 .. 4*lx + 7; only lane 0 (two leading bytes) and the last lane of a warp or
// of the image (two trailing bytes) store bytes.  A x8 cell starts at 8*lx + 4 and stores two aligned words as they are.
template <int S, int CG>
__device__ __forceinline__ void resize_int_u8_planar_body(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H,
                                                          int W, int oH, int oW, const IntGeom<S>& g, const CoefTabs* __restrict__ ct,
                                                          int ly0, int oy0, int oy1, unsigned char* __restrict__ out, int bxi, int byi,
                                                          int p, Smem& sm) {
  static_assert(S == 4 || S == 8, "shuffle epilogue: x4 and x8 only");
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  sm.tab.s2[tid] = __ldg(ct->s2 + tid);
  sm.tab.sg[tid] = __ldg(ct->sg + tid);
  sm.tab.rl[tid] = __ldg(ct->rl + tid);
  __syncthreads();
  const int lx0 = bxi * kCX - 1, lyb = ly0 + byi * kCY;
  stage_coefs(sm, feat, codes, H, W, p, lx0, lyb, tid);
  __syncthreads();
  const int lx = lx0 + tx, ly = lyb + ty;
  if (ly > H - 1) return;  // warp-uniform
  const bool valid = lx <= W - 1;
  const int col = S * lx + g.ph_x;
  unsigned char* plane = out + (long long)p * oH * oW;
  // store predicates and offsets are per thread; nothing below branches (every lane of the warp reaches the shuffle)
  const bool partner = valid && tx < 31 && lx + 1 <= W - 1;                   // x4: the aligned word 4*lx + 4 .. + 7 is mine
  const bool lead = valid && tx == 0 && col >= 0, trail = valid && !partner && col + 3 < oW;
  const bool half = lead || trail, both = lead && trail;                      // 16-bit stores: lane 0 / a lane without partner
  const int off16 = col + (lead ? 0 : 2), sh16 = lead ? 0 : 16;
  const bool whole = valid && col >= 0 && col + S <= oW;                       // x8
  // the cell's first row, once; row mr is a compile-time multiple of the pitch away (one 64-bit multiply-add per pointer and row
  // instead of the full 64-bit row-index product: the kernel is issue-bound and this was 9 instructions per row)
  unsigned char* const row0 = plane + (long long)(S * ly + g.ph_y) * oW;
  unsigned char* const p_al = row0 + col + 2;
  unsigned char* const p_16 = row0 + off16;
  gauss_cell_rowq<S, CG>(sm, g, tx, ty, S * ly + g.ph_y, oy0, oy1, [&](int mr, const uint32_t* res) {
    const long long ro = (long long)mr * oW;
    unsigned char* row = row0 + ro;
    if (S == 4) {
      const uint32_t w = pack4(res[0], res[1], res[2 % S], res[3 % S]);
      const uint32_t n = __shfl_down_sync(0xffffffffu, w, 1);
      uint32_t al;
      asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(al) : "r"(w), "r"(n));
      if (partner) __stcg(reinterpret_cast<uint32_t*>(p_al + ro), al);
      if (half) __stcg(reinterpret_cast<unsigned short*>(p_16 + ro), (unsigned short)(w >> sh16));
      if (both) __stcg(reinterpret_cast<unsigned short*>(p_al + ro), (unsigned short)(w >> 16));  // a one-cell-wide block column
    } else {
      const uint32_t w0 = pack4(res[0], res[1], res[2 % S], res[3 % S]), w1 = pack4(res[4 % S], res[5 % S], res[6 % S], res[7 % S]);
      if (whole) {
        __stcg(reinterpret_cast<uint32_t*>(row + col), w0);
        __stcg(reinterpret_cast<uint32_t*>(row + col + 4), w1);
      } else if (valid) {
#pragma unroll
        for (int mc = 0; mc < S; ++mc)
          if (col + mc >= 0 && col + mc < oW) __stcg(row + col + mc, (unsigned char)((mc < 4 ? w0 : w1) >> (8 * (mc & 3))));
      }
    }
  });
}

// Host: geometry constants of a periodic plan.
inline void make_coef_tabs(float max_sigma, CoefTabs& t) {
  for (int c = 0; c < 256; ++c) {  // hyper decode exactly like numpy in float32 (eval_lut_sr.py:623-628, resize_right2d_numpy.py:168-170)
    volatile float h = (float)c / 255.0f;
    volatile float h2 = h * 2.0f;
    volatile float rho = h2 - 1.0f;
    volatile float sig = h * max_sigma;
    t.s2[c] = -0.5 * kLog2e * ((double)sig * (double)sig);
    t.sg[c] = (double)sig;
    t.rl[c] = kLog2e * (double)rho;
  }
}

// unsigned_form: constants for resize_int_body MODE 2 (negated geometry, magic = 2^(52-FB) + 16 units, inv_scale < 0).
template <int S>
inline IntGeom<S> make_geom(const lerf_sr_plan_impl* P, float max_sigma, bool unsigned_form = false, bool signed_diff = false) {
  IntGeom<S> g;
  double dmax_y = 0.0, dmax_x = 0.0;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k) {
      const double dy = P->ph_dist_y[m][k], dx = P->ph_dist_x[m][k];
      g.dr[m][k] = dy;
      g.dc[m][k] = dx;
      g.xr[m][k] = dy * dy;
      g.xc[m][k] = dx * dx;
      dmax_y = fabs(dy) > dmax_y ? fabs(dy) : dmax_y;
      dmax_x = fabs(dx) > dmax_x ? fabs(dx) : dmax_x;
    }
  for (int mr = 0; mr < S; ++mr)
    for (int mc = 0; mc < S; ++mc)
      for (int b = 0; b < 2; ++b)
        for (int a = 0; a < 2; ++a) g.pp[mr][mc][b][a] = P->ph_dist_y[mr][b] * P->ph_dist_x[mc][a];
  // fixed point: |log2 w| <= L/2 (sigma |dr| + sigma |dc|)^2 (|rho| <= 1) must stay below 2^(31-FB)
  const double reach = (double)max_sigma * (dmax_y + dmax_x);
  const double bound = 0.5 * kLog2e * reach * reach + 2.0;
  if (unsigned_form) {
    int fb = 26;
    while (fb > 8 && bound * (double)(1u << fb) >= (signed_diff ? 2147483000.0 : 4294967000.0)) --fb;
    g.magic = (double)(1ull << (52 - fb)) + 16.0 / (double)(1u << fb);
    g.inv_scale = -1.0f / (float)(1u << fb);
    g.fb = fb;
    for (int m = 0; m < S; ++m)
      for (int k = 0; k < 2; ++k) {
        g.xr[m][k] = -g.xr[m][k];
        g.xc[m][k] = -g.xc[m][k];
      }
    for (int mr = 0; mr < S; ++mr)
      for (int mc = 0; mc < S; ++mc)
        for (int b = 0; b < 2; ++b)
          for (int a = 0; a < 2; ++a) g.pp[mr][mc][b][a] = -g.pp[mr][mc][b][a];
    g.ph_y = P->ph_y;
    g.ph_x = P->ph_x;
    return g;
  }
  int fb = 24;
  while (fb > 8 && bound * (double)(1u << fb) >= 2147483000.0) --fb;
  g.magic = 1.5 * (double)(1ull << (52 - fb));
  g.inv_scale = 1.0f / (float)(1u << fb);
  g.fb = fb;
  g.ph_y = P->ph_y;
  g.ph_x = P->ph_x;
  return g;
}

// true when, for every phase m, tap k = (2m + 1 >= S) is the nearest one (|distance| <= 1/2) on both axes: the condition of
// the nearest-tap weights (combine_ref).  Holds for the geometry of out = S * in: even S have phases (2m+1)/(2S), x3 has
// 1/3, 2/3 (-1/3 to the second tap) and 1 (0 to the second tap).
template <int S>
inline bool ref_tap_ok(const lerf_sr_plan_impl* P, float max_sigma) {
  // the reference tap's own -log2 w is at most L/2 (sigma (1/2 + 1/2))^2: the other weights, relative to it, reach 2^that and
  // the numerator 765 times more -- keep it below 2^100 (sigma <= 11.7; the models use 10), else the minimum form
  if (0.5 * kLog2e * (double)max_sigma * (double)max_sigma > 100.0) return false;
  for (int m = 0; m < S; ++m) {
    const int k = 2 * m + 1 >= S ? 1 : 0;
    if (fabs(P->ph_dist_y[m][k]) > 0.5 + 1e-9 || fabs(P->ph_dist_x[m][k]) > 0.5 + 1e-9) return false;
  }
  return true;
}

// true when the plan's float64 phase distances are exactly CGeom<S>'s (even S, out = S * in)
template <int S>
inline bool geom_is_constexpr(const lerf_sr_plan_impl* P) {
  if (S % 2) return false;
  if (P->ph_y != S / 2 || P->ph_x != S / 2) return false;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k)
      if (P->ph_dist_y[m][k] != CGeom<S>::d(m, k) || P->ph_dist_x[m][k] != CGeom<S>::d(m, k)) return false;
  return true;
}

const CoefTabs* plan_coef_tabs(const lerf_sr_plan_impl* P, float max_sigma, cudaStream_t st);  // resample_int.cu

}  // namespace rsi
}  // namespace lerf

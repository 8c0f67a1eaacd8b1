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
};

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
  constexpr bool kRowQ = MODE == 2;  // S = 8 would need 64 registers for the column terms
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
      for (int t = 0; t < 4; ++t) rowa[t] = fma(ca[t], g.xr[mr][t & 1], g.magic);
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
          double e = fma(cc[t], g.xc[mc][a], rowa[t]);
          e = fma(cb[t], g.pp[mr][mc][b][a], e);
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
        res[mc] = combine_uq(uq, dv, v0, g.inv_scale);
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
inline IntGeom<S> make_geom(const lerf_sr_plan_impl* P, float max_sigma, bool unsigned_form = false) {
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
    while (fb > 8 && bound * (double)(1u << fb) >= 4294967000.0) --fb;
    g.magic = (double)(1ull << (52 - fb)) + 16.0 / (double)(1u << fb);
    g.inv_scale = -1.0f / (float)(1u << fb);
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
  g.ph_y = P->ph_y;
  g.ph_x = P->ph_x;
  return g;
}

const CoefTabs* plan_coef_tabs(const lerf_sr_plan_impl* P, float max_sigma, cudaStream_t st);  // resample_int.cu

}  // namespace rsi
}  // namespace lerf

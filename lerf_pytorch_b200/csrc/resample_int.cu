// Integer-scale specialisation of the LeRF-G SR resampler (S = 2, 3, 4, 8 on both axes, out = S * in).
// Replaces SteeringGaussianResize2dNumpy.resize (resize_right/resize_right2d_numpy.py:162-223 of the
// reference) for the configurations where the geometry is periodic (BASELINE.json cfg-1, -3, -5).
//
// Design.  For an integer scale the S x S output pixels whose first tap is input pixel (ly, lx) share the
// same 2x2 taps, so one thread owns one such CELL: it reads the four taps' coefficients once and produces
// S*S outputs.  Per tap pixel and plane the exponent is a quadratic form
//     log2(w) = a' dr^2 + b' dr dc + c' dc^2,   a' = -L/2 sx^2,  b' = L rho sx sy,  c' = -L/2 sy^2,  L = log2(e)
// whose coefficients are computed once per input sample in float64 from the reference's float32 hyper values
// (a block stages them in shared memory), and whose geometry factors (dr^2, dr dc, dc^2 for the S phases) are
// float64 constants in the kernel parameters.  Each exponent therefore costs 4 FP64 ops.  Adding 1.5*2^(52-FB)
// leaves round(log2(w) * 2^FB) in the low word of the double: the max over the four taps and the subtraction
// are then exact INTEGER ops, and only the difference (<= 0) is converted to fp32 for ex2.approx -- the fp32
// error stays in the low bits of weights that matter.  The output is v00 + sum w_t (v_t - v00) / sum w_t with
// exact integer differences, so fp32 rounding scales with the local contrast, not with 255.
#include "common.cuh"

namespace lerf {

constexpr double kLog2e = 1.4426950408889634;

template <int S>
struct IntGeom {
  double xr[S][2];        // dr^2       [row phase][tap b]
  double xc[S][2];        // dc^2       [col phase][tap a]
  double pp[S][S][2][2];  // dr * dc    [row phase][col phase][b][a]
  double magic;           // 1.5 * 2^(52 - FB)
  float inv_scale;        // 2^-FB
  int ph_y, ph_x;         // first output of cell l is S*l + ph
};

constexpr int kCX = 32, kCY = 8;  // cells per block

struct CoefTabs {           // per-code float64 tables, exact promotions of the reference's float32 values
  double s2[256];           // -L/2 * sigma^2
  double sg[256];           // sigma
  double rl[256];           // L * rho
};

template <int FMT>
__device__ __forceinline__ void store1(void* out, long long ip, long long ih, float val) {
  if (FMT == LERF_OUT_F32) {
    ((float*)out)[ip] = val;
  } else {
    int q = __float2int_rn(val);  // round half to even
    q = min(max(q, 0), 255);
    ((uint8_t*)out)[FMT == LERF_OUT_U8 ? ip : ih] = (uint8_t)q;
  }
}

template <int S, int FMT>
__global__ void __launch_bounds__(kCX* kCY)
    resize_sr_int_gauss_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W,
                               int oH, int oW, IntGeom<S> g, float max_sigma, int channels, int ly0, int oy0,
                               int oy1, void* __restrict__ out) {
  __shared__ CoefTabs tab;
  __shared__ double sA[kCY + 1][kCX + 1], sB[kCY + 1][kCX + 1], sC[kCY + 1][kCX + 1];
  __shared__ float sV[kCY + 1][kCX + 1];
  const int tid = threadIdx.y * kCX + threadIdx.x;
  {  // hyper decode exactly like numpy in float32 (eval_lut_sr.py:623-628, resize_right2d_numpy.py:168-170)
    const float h = __fdiv_rn((float)tid, 255.0f);
    const float rho = __fsub_rn(__fmul_rn(h, 2.0f), 1.0f);
    const float sig = __fmul_rn(h, max_sigma);
    tab.s2[tid] = -0.5 * kLog2e * ((double)sig * (double)sig);
    tab.sg[tid] = (double)sig;
    tab.rl[tid] = kLog2e * (double)rho;
  }
  __syncthreads();
  const int p = blockIdx.z;
  const int lx0 = (int)blockIdx.x * kCX - 1;        // first cell column of the block (cells start at -1)
  const int lyb = ly0 + (int)blockIdx.y * kCY;      // first cell row of the block
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  const uint8_t* cp = codes + (long long)p * 3 * plane_sz;
  for (int i = tid; i < (kCY + 1) * (kCX + 1); i += kCX * kCY) {
    const int r = i / (kCX + 1), c = i - r * (kCX + 1);
    const int sy = lyb + r, sx = lx0 + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'edge' (:172-174)
    const long long off = (long long)cy * W + cx;
    const int kr = __ldg(cp + off), kx = __ldg(cp + plane_sz + off), ky = __ldg(cp + 2 * plane_sz + off);
    sA[r][c] = tab.s2[kx];
    sC[r][c] = tab.s2[ky];
    sB[r][c] = tab.rl[kr] * tab.sg[kx] * tab.sg[ky];
    sV[r][c] = (sy == cy && sx == cx) ? (float)__ldg(fp + off) : 0.0f;   // image: 'constant' 0 (:208)
  }
  __syncthreads();
  const int lx = lx0 + threadIdx.x, ly = lyb + threadIdx.y;
  if (lx > W - 1 || ly > H - 1) return;
  // taps t = a*2+b: row ly+b, column lx+a (same patch order as the reference, :95-98)
  double ca[4], cb[4], cc[4];
  float dv[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      ca[a * 2 + b] = sA[threadIdx.y + b][threadIdx.x + a];
      cb[a * 2 + b] = sB[threadIdx.y + b][threadIdx.x + a];
      cc[a * 2 + b] = sC[threadIdx.y + b][threadIdx.x + a];
      dv[a * 2 + b] = sV[threadIdx.y + b][threadIdx.x + a];
    }
  const float v0 = dv[0];
  dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;  // exact: integers in [-255, 255]
  const int oyb = S * ly + g.ph_y, oxb = S * lx + g.ph_x;
  const long long pbase = (long long)p * oH;
  const long long hbase = (long long)(p / channels) * oH;
  const int pc_ = p % channels;
#pragma unroll
  for (int mr = 0; mr < S; ++mr) {
    const int oy = oyb + mr;
    if (oy < oy0 || oy >= oy1) continue;
    float res[S];
#pragma unroll
    for (int mc = 0; mc < S; ++mc) {
      int q[4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int t = a * 2 + b;
          double e = cc[t] * g.xc[mc][a];
          e = fma(cb[t], g.pp[mr][mc][b][a], e);
          e = fma(ca[t], g.xr[mr][b], e);
          q[t] = __double2loint(e + g.magic);  // round(log2 w * 2^FB), two's complement
        }
      const int qm = max(max(q[0], q[1]), max(q[2], q[3]));
      float w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float x = (float)(q[t] - qm) * g.inv_scale;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(x));
      }
      const float den = (w[0] + w[1]) + (w[2] + w[3]);            // in [1, 4]: the max tap has weight 1
      const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
      r = fmaf(r, fmaf(-den, r, 1.0f), r);                        // one Newton step
      float qn = num * r;
      qn = fmaf(fmaf(-den, qn, num), r, qn);                      // residual correction: quotient to ~0.5 ulp
      res[mc] = v0 + qn;
    }
    const long long rowp = (pbase + oy) * oW, rowh = (hbase + oy) * oW;
    const bool full = oxb >= 0 && oxb + S <= oW;
    if (FMT == LERF_OUT_F32 && full && (S % 2 == 0)) {
      float* o = (float*)out + rowp + oxb;
      if (S == 8) {  // ph = 4: 16-byte aligned
        *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(res[4 % S], res[5 % S], res[6 % S], res[7 % S]);
      } else if (S == 4) {  // ph = 2: 8-byte aligned
        *reinterpret_cast<float2*>(o) = make_float2(res[0], res[1]);
        *reinterpret_cast<float2*>(o + 2) = make_float2(res[2 % S], res[3 % S]);
      } else {
        o[0] = res[0];
        o[1] = res[1 % S];
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

template <int S>
static int launch_int(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                      float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  IntGeom<S> g;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k) {
      g.xr[m][k] = P->ph_dist_y[m][k] * P->ph_dist_y[m][k];
      g.xc[m][k] = P->ph_dist_x[m][k] * P->ph_dist_x[m][k];
    }
  for (int mr = 0; mr < S; ++mr)
    for (int mc = 0; mc < S; ++mc)
      for (int b = 0; b < 2; ++b)
        for (int a = 0; a < 2; ++a) g.pp[mr][mc][b][a] = P->ph_dist_y[mr][b] * P->ph_dist_x[mc][a];
  // fixed point: |log2 w| <= 2 L (max_sigma * dmax)^2 must stay below 2^(31-FB); SR distances are <= 1
  const double bound = 2.0 * kLog2e * (double)max_sigma * (double)max_sigma + 1.0;
  int fb = 24;
  while (fb > 8 && bound * (double)(1u << fb) >= 2147483000.0) --fb;
  g.magic = 1.5 * (double)(1ull << (52 - fb));
  g.inv_scale = 1.0f / (float)(1u << fb);
  g.ph_y = P->ph_y;
  g.ph_x = P->ph_x;
  // cell rows touched by the output band
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kCX, kCY), grid((P->W + 1 + kCX - 1) / kCX, (ly1 - ly0 + 1 + kCY - 1) / kCY, planes);
#define LERF_GO(F)                                                                                               \
  resize_sr_int_gauss_kernel<S, F><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, max_sigma, \
                                                           channels, ly0, oy0, oy1, out)
  switch (fmt) {
    case LERF_OUT_F32: LERF_GO(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GO(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GO(LERF_OUT_U8_HWC); break;
    default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GO
  LERF_LAUNCHED();
  return LERF_OK;
}

// Called by lerf_resize_sr when the plan is periodic.  Returns -1 when this path does not apply.
int resize_sr_int_gauss(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                        float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  if (!(max_sigma >= 0.0f) || max_sigma > 64.0f) return -1;
  if (fmt == LERF_OUT_F32 && ((uintptr_t)out & 15)) return -1;
  switch (P->int_scale) {
    case 2: return launch_int<2>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 3: return launch_int<3>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 4: return launch_int<4>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 8: return launch_int<8>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    default: return -1;
  }
}

}  // namespace lerf

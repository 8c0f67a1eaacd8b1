// Integer-scale LeRF-G SR resampler: plain launches.  Kernel body and design notes: resample_int.cuh.
#include <math.h>

#include <mutex>

#include "resample_int.cuh"

namespace lerf {

using namespace rsi;

// Device copies of the per-code tables, one per max_sigma seen (at most kCoefSlots), owned by the plan.  A copy is
// built and uploaded with a SYNCHRONOUS cudaMemcpy under a mutex the first time its sigma is asked for and never
// written again, so any number of streams and threads can launch on one plan (the r1 version uploaded on the
// calling stream only and overwrote the table when sigma changed).
const CoefTabs* rsi::plan_coef_tabs(const lerf_sr_plan_impl* P, float max_sigma, cudaStream_t) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  lerf_sr_plan_impl* M = const_cast<lerf_sr_plan_impl*>(P);
  for (int i = 0; i < M->coef_n; ++i)
    if (M->coef_sigma[i] == max_sigma) return (const CoefTabs*)M->coef_dev[i];
  if (M->coef_n >= lerf_sr_plan_impl::kCoefSlots) return nullptr;
  CoefTabs* host = new CoefTabs;
  make_coef_tabs(max_sigma, *host);
  void* dev = nullptr;
  bool ok = cudaMalloc(&dev, sizeof(CoefTabs)) == cudaSuccess &&
            cudaMemcpy(dev, host, sizeof(CoefTabs), cudaMemcpyHostToDevice) == cudaSuccess;
  delete host;
  if (!ok) {
    cudaFree(dev);
    cudaGetLastError();
    return nullptr;
  }
  M->coef_dev[M->coef_n] = dev;
  M->coef_sigma[M->coef_n] = max_sigma;
  ++M->coef_n;
  return (const CoefTabs*)dev;
}

// testing hook: 0 = production = ROWQ form (unsigned fixed point, row term hoisted; resample_int.cuh), 4 blocks/SM;
// 4 = ROWQ, 5 blocks/SM; 5 = plain form (4 FP64 per exponent), 5 blocks/SM (production until r1d); 2 = plain, 4 blocks/SM;
// 1 = fully hoisted signed form, 3 blocks/SM

template <int S, int FMT, int MODE, int MINB>
__global__ void __launch_bounds__(kCX* kCY, MINB)
    resize_sr_int_gauss_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH,
                               int oW, const __grid_constant__ IntGeom<S> g, const CoefTabs* __restrict__ ct,
                               int channels, int ly0, int oy0, int oy1, void* __restrict__ out) {
  __shared__ Smem sm;
  resize_int_body<S, FMT, MODE>(feat, codes, H, W, oH, oW, g, ct, channels, ly0, oy0, oy1, out, blockIdx.x, blockIdx.y,
                                 blockIdx.z, sm);
}

// uint8 outputs through a staged tile (resample_int.cuh): CH = 1 planar, CH = 3 interleaved
template <int S, int CH, int CG, bool TMA = true>
__global__ void __launch_bounds__(kCX* kCY, 4)
    resize_sr_int_gauss_u8_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH,
                                  int oW, const __grid_constant__ IntGeom<S> g, const CoefTabs* __restrict__ ct, int ly0,
                                  int oy0, int oy1, unsigned char* __restrict__ out) {
  __shared__ Smem sm;
  __shared__ OutTile<S, CH> ot;
  resize_int_u8_body<S, CH, CG, TMA>(feat, codes, H, W, oH, oW, g, ct, ly0, oy0, oy1, out, blockIdx.x, blockIdx.y, blockIdx.z, sm, ot);
}

template <int S, int CG>
__global__ void __launch_bounds__(kCX* kCY, 4)
    resize_sr_int_gauss_u8p_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH,
                                   int oW, const __grid_constant__ IntGeom<S> g, const CoefTabs* __restrict__ ct, int ly0,
                                   int oy0, int oy1, unsigned char* __restrict__ out) {
  __shared__ Smem sm;
  resize_int_u8_planar_body<S, CG>(feat, codes, H, W, oH, oW, g, ct, ly0, oy0, oy1, out, blockIdx.x, blockIdx.y, blockIdx.z, sm);
}


template <int S>
static int launch_int(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                      float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
#ifdef LERF_EXPERIMENTS
  const int rv = g_dbg.resize_variant == 12 ? 11 : g_dbg.resize_variant;
#else
  // 11: geometry from kernel parameters (the flavour of odd scales); 12: 11 + HWC tile copied out by lanes instead of bulk
  // stores; 7 / 13: weights relative to the phase's nearest tap / to the smallest exponent (one of them is production)
  const int rv = g_dbg.resize_variant == 11 || g_dbg.resize_variant == 12 ? 11 : (g_dbg.resize_variant == 7 || g_dbg.resize_variant == 13 ? g_dbg.resize_variant : 0);
#endif
  const bool prod = rv == 0 || rv == 7 || rv == 13;
  const bool cg = prod && geom_is_constexpr<S>(P);  // geometry factors as immediates (resample_int.cuh CGeom)
  bool ref = prod && ref_tap_ok<S>(P, max_sigma) && (rv == 7 || (rv == 0 && kRefTapDefault));  // weights relative to the nearest tap (combine_ref)
  IntGeom<S> g = make_geom<S>(P, max_sigma, /*unsigned_form=*/prod || rv == 4 || rv == 11, /*signed_diff=*/ref);
  if (ref && g.fb < kMinFracBits) {  // the signed differences cost a bit: the minimum form keeps it
    ref = false;
    g = make_geom<S>(P, max_sigma, true, false);
  }
  if (prod && g.fb < kMinFracBits) return -1;  // too wide an exponent range for the fixed point (large max_sigma): next kernel
  const CoefTabs* ct = plan_coef_tabs(P, max_sigma, st);
  if (!ct) return fail(LERF_ECUDA, "uploading the hyper decode tables failed");
  // cell rows touched by the output band
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kCX * kCY), grid((P->W + 1 + kCX - 1) / kCX, (ly1 - ly0 + 1 + kCY - 1) / kCY, planes);
#define LERF_U8P(CGV) resize_sr_int_gauss_u8p_kernel<S, CGV><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, ct, ly0, oy0, oy1, (unsigned char*)out)
#define LERF_U8T(CHV, CGV) resize_sr_int_gauss_u8_kernel<S, CHV, CGV><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, ct, ly0, oy0, oy1, (unsigned char*)out)
  // staged uint8 epilogue: the tile must fit the 48 KiB of static shared memory next to the coefficient tiles
  if (g_dbg.u8_staged && (prod || rv == 11)) {
    if constexpr (S == 4 || S == 8) {  // planar: aligned words through a lane shuffle (x4) or as they are (x8)
      if (fmt == LERF_OUT_U8 && g.ph_x == S / 2 && P->oW % 4 == 0 && ((uintptr_t)out & 3) == 0) {
        if (ref && cg) LERF_U8P(3);
        else if (ref) LERF_U8P(2);
        else if (cg) LERF_U8P(1);
        else LERF_U8P(0);
        LERF_LAUNCHED();
        return LERF_OK;
      }
    }
    if constexpr (S != 4 && S != 8) {  // planar at x2 / x3: the staged tile
      if (fmt == LERF_OUT_U8) {
        if (ref && cg) LERF_U8T(1, 3);
        else if (ref) LERF_U8T(1, 2);
        else if (cg) LERF_U8T(1, 1);
        else LERF_U8T(1, 0);
        LERF_LAUNCHED();
        return LERF_OK;
      }
    }
    if constexpr (sizeof(OutTile<S, 3>) + sizeof(Smem) <= 48 * 1024) {
      if (fmt == LERF_OUT_U8_HWC && channels == 3 && planes % 3 == 0) {
        grid.z = planes / 3;
        if (g_dbg.resize_variant == 12) resize_sr_int_gauss_u8_kernel<S, 3, 0, false><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, ct, ly0, oy0, oy1, (unsigned char*)out);
        else if (ref && cg) LERF_U8T(3, 3);
        else if (ref) LERF_U8T(3, 2);
        else if (cg) LERF_U8T(3, 1);
        else LERF_U8T(3, 0);
        LERF_LAUNCHED();
        return LERF_OK;
      }
    }
  }
#undef LERF_U8P
#undef LERF_U8T
#define LERF_GK(F, HO, B)                                                                                              \
  resize_sr_int_gauss_kernel<S, F, HO, B><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, ct, \
                                                                  channels, ly0, oy0, oy1, out)
#ifdef LERF_EXPERIMENTS
#define LERF_GO(F)                                              \
  if (g_dbg.resize_variant == 1) LERF_GK(F, 1, 3);              \
  else if (g_dbg.resize_variant == 2) LERF_GK(F, 0, 4);         \
  else if (g_dbg.resize_variant == 5) LERF_GK(F, 0, 5);         \
  else if (g_dbg.resize_variant == 4) LERF_GK(F, 2, 5);         \
  else if (ref && cg) LERF_GK(F, 5, 4);                         \
  else if (ref) LERF_GK(F, 6, 4);                               \
  else if (cg) LERF_GK(F, 3, 4);                                \
  else LERF_GK(F, 2, 4)
#else
#define LERF_GO(F)               \
  if (ref && cg) LERF_GK(F, 5, 4); \
  else if (ref) LERF_GK(F, 6, 4);  \
  else if (cg) LERF_GK(F, 3, 4);   \
  else LERF_GK(F, 2, 4)
#endif
  switch (fmt) {
    case LERF_OUT_F32: LERF_GO(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GO(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GO(LERF_OUT_U8_HWC); break;
    default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GO
#undef LERF_GK
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

// ---------------------------------------------------------------------------------------------------------------
// Integer-scale cell-owner kernel for the amplified-linear kind (LeRF-L at x2, x3, x4, x8: the scales of the published
// LeRF-L table, scripts.sh:36-38).  Replaces AmplifiedLinearResize2dNumpy.resize (resize_right2d_numpy.py:243-282) where
// the geometry is periodic.  One thread = one cell (the S x S outputs sharing their 2x2 taps); the four taps' alpha
// values are read once; weights are (1 - alpha |dr|)(1 - alpha |dc|) in float64 like the reference (:233-241) with the
// phase distances as kernel-parameter constants, then fp32 normalisation around v0 with exact integer differences --
// the arithmetic of the tile kernel's linear branch (resample_tile.cu), which stays the path for other scales and for
// max_sigma > 1 (the max(., 0) clamp) or distances outside [-1, 1].
// ---------------------------------------------------------------------------------------------------------------
namespace lerf {
namespace {

template <int S>
struct LinGeom {
  double adr[S][2], adc[S][2];  // |dr| [row phase][tap b], |dc| [col phase][tap a]
  // A phase that sits ON the window edge (|d| = 1 up to rounding; x3 has one): the linear kernel is discontinuous there
  // (weight 1 - alpha inside, 0 outside) and the per-output distances wobble by an ulp around the phase value, so such a
  // tap is decided per output sample from the plan's float64 distance tables, exactly like the reference.
  unsigned char edge_r[S][2], edge_c[S][2];
  int ph_y, ph_x;
};

struct SmemLin {
  double al[256];
  double sA[kCY + 1][kCX + 1];
  float sV[kCY + 1][kCX + 1];
};

// EDGE: some phase sits on the window edge (x3); the lean instantiation (x2, x4, x8) carries none of that logic.
template <int S, int FMT, bool EDGE>
__global__ void __launch_bounds__(kCX* kCY)
    resize_sr_int_linear_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH, int oW,
                                const __grid_constant__ LinGeom<S> g, const double* __restrict__ dist_y,
                                const double* __restrict__ dist_x, float max_sigma, int channels, int ly0, int oy0, int oy1,
                                void* __restrict__ out) {
  __shared__ SmemLin sm;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int p = blockIdx.z;
  {  // alpha = fl(max_sigma * rho) in float32 like numpy (:249-250), promoted exactly
    const float h = __fdiv_rn((float)tid, 255.0f);
    sm.al[tid] = (double)__fmul_rn(max_sigma, __fsub_rn(__fmul_rn(h, 2.0f), 1.0f));
  }
  __syncthreads();
  const int lx0 = blockIdx.x * kCX - 1, lyb = ly0 + blockIdx.y * kCY;
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  const uint8_t* cp = codes + (long long)p * plane_sz;
  for (int i = tid; i < (kCY + 1) * (kCX + 1); i += kCX * kCY) {
    const int r = i / (kCX + 1), c = i - r * (kCX + 1);
    const int sy = lyb + r, sx = lx0 + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'edge' (:252-254)
    const long long off = (long long)cy * W + cx;
    sm.sA[r][c] = sm.al[__ldcg(cp + off)];
    sm.sV[r][c] = (sy == cy && sx == cx) ? (float)__ldcg(fp + off) : 0.0f;  // image: 'constant' 0 (:268)
  }
  __syncthreads();
  const int lx = lx0 + tx, ly = lyb + ty;
  if (lx > W - 1 || ly > H - 1) return;
  double al[4];
  float dv[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {  // taps t = a*2+b: row ly+b, column lx+a (patch order of the reference, :95-98)
      al[a * 2 + b] = sm.sA[ty + b][tx + a];
      dv[a * 2 + b] = sm.sV[ty + b][tx + a];
    }
  const float v0 = dv[0];
  dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;  // exact: integers in [-255, 255]
  const int oyb = S * ly + g.ph_y, oxb = S * lx + g.ph_x;
  const int pc_ = p % channels;
  long long rowp = ((long long)p * oH + oyb) * oW;
  long long rowh = ((long long)(p / channels) * oH + oyb) * oW;
  const bool full = oxb >= 0 && oxb + S <= oW;
  // column factors of edge phases, from the exact per-output distances (|d| and the [|d| <= 1] window)
  double edc[S][2];
  float evc[S][2];
#pragma unroll
  for (int mc = 0; mc < S; ++mc)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      edc[mc][a] = g.adc[mc][a];
      evc[mc][a] = 1.0f;
      if (EDGE && g.edge_c[mc][a]) {
        const int ox = min(max(oxb + mc, 0), oW - 1);
        edc[mc][a] = fabs(__ldg(dist_x + 2 * ox + a));
        evc[mc][a] = edc[mc][a] <= 1.0 ? 1.0f : 0.0f;
      }
    }
#pragma unroll
  for (int mr = 0; mr < S; ++mr, rowp += oW, rowh += oW) {
    const int oy = oyb + mr;
    if (oy < oy0 || oy >= oy1) continue;
    double lr[4];
    float vr[2] = {1.0f, 1.0f};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      double d = g.adr[mr][b];
      if (EDGE && g.edge_r[mr][b]) {
        d = fabs(__ldg(dist_y + 2 * oy + b));
        vr[b] = d <= 1.0 ? 1.0f : 0.0f;
      }
      lr[b] = fma(-al[b], d, 1.0);
      lr[2 + b] = fma(-al[2 + b], d, 1.0);
    }
    float res[S];
#pragma unroll
    for (int mc = 0; mc < S; ++mc) {
      float w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        w[t] = (float)(lr[t] * fma(-al[t], edc[mc][t >> 1], 1.0));
        if (EDGE && (g.edge_r[mr][t & 1] || g.edge_c[mc][t >> 1])) w[t] *= vr[t & 1] * evc[mc][t >> 1];
      }
      const float den = (w[0] + w[1]) + (w[2] + w[3]);
      const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
      float qn = num * r;
      qn = fmaf(fmaf(-den, qn, num), r, qn);
      res[mc] = v0 + qn;
    }
    if (FMT == LERF_OUT_F32 && full && (S % 2 == 0)) {  // same vector stores as the Gaussian kernel (resample_int.cuh)
      float* o = (float*)out + rowp + oxb;
      if (S == 8) {
        __stcg(reinterpret_cast<float4*>(o), make_float4(res[0], res[1], res[2], res[3]));
        __stcg(reinterpret_cast<float4*>(o + 4), make_float4(res[4 % S], res[5 % S], res[6 % S], res[7 % S]));
      } else if (S == 4) {
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

template <int S>
int launch_int_linear(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                      float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  LinGeom<S> g;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k) {
      g.adr[m][k] = fabs(P->ph_dist_y[m][k]);
      g.adc[m][k] = fabs(P->ph_dist_x[m][k]);
      if (g.adr[m][k] > 1.0 + 1e-6 || g.adc[m][k] > 1.0 + 1e-6) return -1;  // outside the window for good: tile kernel
      // on the window edge: decided per output sample in the kernel (taking the phase constant moved the Set5 x3 entry
      // of the published LeRF-L table from 30.72 to 30.76 dB)
      g.edge_r[m][k] = g.adr[m][k] > 1.0 - 1e-6;
      g.edge_c[m][k] = g.adc[m][k] > 1.0 - 1e-6;
    }
  g.ph_y = P->ph_y;
  g.ph_x = P->ph_x;
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kCX * kCY), grid((P->W + 1 + kCX - 1) / kCX, (ly1 - ly0 + 1 + kCY - 1) / kCY, planes);
  bool edge = false;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k) edge = edge || g.edge_r[m][k] || g.edge_c[m][k];
#define LERF_GL(F)                                                                                                                  \
  if (edge)                                                                                                                         \
    resize_sr_int_linear_kernel<S, F, true><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, P->dist_y, P->dist_x,  \
                                                                    max_sigma, channels, ly0, oy0, oy1, out);                        \
  else                                                                                                                              \
    resize_sr_int_linear_kernel<S, F, false><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, g, P->dist_y, P->dist_x, \
                                                                     max_sigma, channels, ly0, oy0, oy1, out)
  switch (fmt) {
    case LERF_OUT_F32: LERF_GL(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GL(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GL(LERF_OUT_U8_HWC); break;
    default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GL
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // namespace

// Called by lerf_resize_sr (kind = linear) when the plan is periodic.  Returns -1 when this path does not apply.
int resize_sr_int_linear(const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                         float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  if (!(max_sigma >= 0.0f) || max_sigma > 1.0f) return -1;  // |alpha| <= 1 keeps both factors >= 0: no clamp needed
  if (fmt == LERF_OUT_F32 && ((uintptr_t)out & 15)) return -1;
  switch (P->int_scale) {
    case 2: return launch_int_linear<2>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 3: return launch_int_linear<3>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 4: return launch_int_linear<4>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    case 8: return launch_int_linear<8>(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, fmt, st);
    default: return -1;
  }
}

}  // namespace lerf

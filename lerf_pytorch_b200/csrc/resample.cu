// Resampling half of the LeRF hot path for sm_100a: spatially-varying steerable resampling with
// per-TAP hyper-parameters (anisotropic Gaussian for LeRF-G, amplified linear for LeRF-L) for
// arbitrary-scale SR and homographic warping.
//
// Reference being replaced (ddlee-cn/LeRF-PyTorch, resize_right/resize_right2d_numpy.py):
//   Resize2dNumpy.set_shape/get_distance :18-140       SteeringGaussianResize2dNumpy.resize :162-223
//   AmplifiedLinearResize2dNumpy.resize  :243-282      Warp2dNumpy.set_shape/get_distance   :292-407
//   SteeringGaussianWarp2dNumpy.warp     :516-577      AmplifiedLinearWarp2dNumpy.warp      :597-635
//   NearestWarp2dNumpy                   :460-467      uint8 epilogue eval_lut_sr.py:663-665
//
// Design: the reference materialises [C, 2*oH, 2*oW] float64 index/distance/weight arrays (39 GB for
// one 2K x4 frame).  Here the SR geometry is two 1-D float64 tables (it is separable), each thread
// owns one output pixel of one plane, gathers its 2x2 taps' uint8 hyper codes and image bytes,
// evaluates the exponent in float64 exactly in the reference's operation order, subtracts the
// largest exponent and only then drops to fp32 for ex2 -- so the fp32 error is confined to the
// weights' low bits -- and accumulates/divides in float64.  Nothing but the output touches HBM.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace lerf {

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return min(max(v, lo), hi); }

// hyper decode tables, float32 exactly like numpy: h = fl(code/255); rho = fl(fl(2h)-1);
// sigma = fl(h*max_sigma)  (eval_lut_sr.py:623-628, resize_right2d_numpy.py:168-170, :249-250)
struct HyperTab {
  float rho[256];
  float sig[256];
};

__device__ __forceinline__ void build_hyper_tab(HyperTab& t, float max_sigma, int tid, int nthreads) {
  for (int c = tid; c < 256; c += nthreads) {
    const float h = __fdiv_rn((float)c, 255.0f);
    t.rho[c] = __fsub_rn(__fmul_rn(h, 2.0f), 1.0f);
    t.sig[c] = __fmul_rn(h, max_sigma);
  }
}

// hyper sources: uint8 codes through the table, or float32 planes decoded on the fly
struct CodeSrc {
  const uint8_t* codes;  // [P*oC][H][W]
  __device__ __forceinline__ void gauss(const HyperTab& t, long long plane_sz, int p, long long off, float max_sigma,
                                        float& rho, float& sx, float& sy) const {
    (void)max_sigma;
    const uint8_t* b = codes + (long long)p * 3 * plane_sz + off;
    rho = t.rho[__ldg(b)];
    sx = t.sig[__ldg(b + plane_sz)];
    sy = t.sig[__ldg(b + 2 * plane_sz)];
  }
  __device__ __forceinline__ float alpha(const HyperTab& t, long long plane_sz, int p, long long off, float max_sigma) const {
    return __fmul_rn(max_sigma, t.rho[__ldg(codes + (long long)p * plane_sz + off)]);
  }
};
struct FloatSrc {
  const float *h0, *h1, *h2;  // each [P][H][W]
  __device__ __forceinline__ void gauss(const HyperTab&, long long plane_sz, int p, long long off, float max_sigma,
                                        float& rho, float& sx, float& sy) const {
    const long long i = (long long)p * plane_sz + off;
    rho = __fsub_rn(__fmul_rn(__ldg(h0 + i), 2.0f), 1.0f);
    sx = __fmul_rn(__ldg(h1 + i), max_sigma);
    sy = __fmul_rn(__ldg(h2 + i), max_sigma);
  }
  __device__ __forceinline__ float alpha(const HyperTab&, long long plane_sz, int p, long long off, float max_sigma) const {
    return __fmul_rn(max_sigma, __fsub_rn(__fmul_rn(__ldg(h0 + (long long)p * plane_sz + off), 2.0f), 1.0f));
  }
};

// exponent of the steerable Gaussian, float64, operation order of sk_weight (:150-160)
__device__ __forceinline__ double gauss_exponent(float rho, float sx, float sy, double dr, double dc) {
  const double sxx = (double)sx * dr;
  const double syy = (double)sy * dc;
  const double xy = sxx * (double)sy * dc;
  return -0.5 * (sxx * sxx - (double)(2.0f * rho) * xy + syy * syy);
}

// linear_alpha / linear_weight (:233-241)
__device__ __forceinline__ double lin_alpha(double x, double a) {
  double r = 0.0;
  if (-1.0 <= x && x < 0.0) r = a * x + 1.0;
  if (0.0 <= x && x <= 1.0) r = 1.0 - a * x;
  return r > 0.0 ? r : 0.0;
}

// 4 taps -> one output sample.  e[] holds exponents (Gauss) or weights (linear).
template <int KIND>
__device__ __forceinline__ double combine4(const double e[4], const float v[4]) {
  if (KIND == LERF_KIND_GAUSS) {
    const double m = fmax(fmax(e[0], e[1]), fmax(e[2], e[3]));
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float t = (float)((e[k] - m) * 1.4426950408889634);  // log2(e)
      float w;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(t));
      num += (double)w * (double)v[k];
      den += (double)w;
    }
    return num / den;  // den >= 1: the max tap has weight exactly 1
  } else {
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      num += e[k] * (double)v[k];
      den += e[k];
    }
    return num / den;  // 0/0 -> NaN exactly where numpy gives NaN
  }
}

template <int FMT>
__device__ __forceinline__ void store_sample(void* out, int FMT_unused, long long idx_planar, long long idx_hwc, double val) {
  (void)FMT_unused;
  if (FMT == LERF_OUT_F32) {
    ((float*)out)[idx_planar] = (float)val;
  } else {
    int q = __double2int_rn(val);  // round half to even; NaN -> 0
    q = min(max(q, 0), 255);
    ((uint8_t*)out)[FMT == LERF_OUT_U8 ? idx_planar : idx_hwc] = (uint8_t)q;
  }
}

template <typename ImgT>
__device__ __forceinline__ float load_img(const ImgT* p) { return (float)__ldg(p); }

// ---------------------------------------------------------------------------------------------
// SR, generic scale: one thread = one output sample
// ---------------------------------------------------------------------------------------------
template <int KIND, int FMT, typename ImgT, typename Hyp>
__global__ void __launch_bounds__(256)
    resize_sr_generic_kernel(const ImgT* __restrict__ img, Hyp hyp, int H, int W, int oH, int oW,
                             const int* __restrict__ left_y, const double* __restrict__ dist_y,
                             const int* __restrict__ left_x, const double* __restrict__ dist_x, int channels,
                             float max_sigma, int oy0, int oy1, void* __restrict__ out) {
  __shared__ HyperTab tab;
  build_hyper_tab(tab, max_sigma, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y);
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = oy0 + blockIdx.y * blockDim.y + threadIdx.y;
  const int p = blockIdx.z;
  if (ox >= oW || oy >= oy1) return;
  const long long plane_sz = (long long)H * W;
  const int ly = left_y[oy], lx = left_x[ox];
  const double dr[2] = {dist_y[2 * oy], dist_y[2 * oy + 1]};
  const double dc[2] = {dist_x[2 * ox], dist_x[2 * ox + 1]};
  double e[4];
  float v[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)      // column tap
#pragma unroll
    for (int b = 0; b < 2; ++b) {  // row tap; patch order a*2+b as in the reference (:95-98)
      const int sy = ly + b, sx = lx + a;
      const int cy = clampi2(sy, 0, H - 1), cx = clampi2(sx, 0, W - 1);  // hypers: 'edge' (:172-174)
      const long long off = (long long)cy * W + cx;
      if (KIND == LERF_KIND_GAUSS) {
        float rho, s_x, s_y;
        hyp.gauss(tab, plane_sz, p, off, max_sigma, rho, s_x, s_y);
        e[a * 2 + b] = gauss_exponent(rho, s_x, s_y, dr[b], dc[a]);
      } else {
        const double al = (double)hyp.alpha(tab, plane_sz, p, off, max_sigma);
        e[a * 2 + b] = lin_alpha(dr[b], al) * lin_alpha(dc[a], al);
      }
      const bool inside = (sy == cy) && (sx == cx);                       // image: 'constant' 0 (:208)
      v[a * 2 + b] = inside ? load_img(img + (long long)p * plane_sz + off) : 0.0f;
    }
  const double val = combine4<KIND>(e, v);
  const long long ip = ((long long)p * oH + oy) * oW + ox;
  const long long ih = (((long long)(p / channels) * oH + oy) * oW + ox) * channels + (p % channels);
  store_sample<FMT>(out, 0, ip, ih, val);
}

// ---------------------------------------------------------------------------------------------
// SR with the reference's NON-DEFAULT operator parameters (r2): any support size, np.pad mode of the image, and the
// antialias scaling that a height factor below 1 turns on (resize_right2d_numpy.py:51-55, :186-193).  One thread = one
// output sample, supp x supp taps, float64 in the reference's operation order (weights / sum first, then the products);
// the Gaussian exponents are max-subtracted before exp (float64 exp: this is the parity path, not a fast one).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pad_src(int i, int n, int mode) {  // np.pad source index of UN-padded position i, -1 = constant 0
  if (i >= 0 && i < n) return i;
  switch (mode) {
    case 0: return -1;
    case 1: return i < 0 ? 0 : n - 1;
    case 2: {
      if (n == 1) return 0;
      const int period = 2 * (n - 1);
      i %= period;
      if (i < 0) i += period;
      return i < n ? i : period - i;
    }
    case 3: {
      const int period = 2 * n;
      i %= period;
      if (i < 0) i += period;
      return i < n ? i : period - 1 - i;
    }
    default: {
      i %= n;
      return i < 0 ? i + n : i;
    }
  }
}

template <int KIND, int FMT, typename ImgT, typename Hyp>
__global__ void __launch_bounds__(256)
    resize_sr_support_kernel(const ImgT* __restrict__ img, Hyp hyp, int H, int W, int oH, int oW, int supp,
                             const int* __restrict__ left_y, const double* __restrict__ dist_y,
                             const int* __restrict__ left_x, const double* __restrict__ dist_x, int pad_mode, double aa,
                             int channels, float max_sigma, int oy0, int oy1, void* __restrict__ out) {
  __shared__ HyperTab tab;
  build_hyper_tab(tab, max_sigma, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y);
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = oy0 + blockIdx.y * blockDim.y + threadIdx.y;
  const int p = blockIdx.z;
  if (ox >= oW || oy >= oy1) return;
  const long long plane_sz = (long long)H * W;
  const int ly = left_y[oy], lx = left_x[ox];
  const double* dr = dist_y + (long long)supp * oy;
  const double* dc = dist_x + (long long)supp * ox;
  // pass 1: the largest exponent (Gaussian) / the weight sum (linear); pass 2: weights / sum * value, in patch order a*supp+b
  double m = -INFINITY, sum = 0.0;
  for (int a = 0; a < supp; ++a)
    for (int b = 0; b < supp; ++b) {
      const int cy = clampi2(ly + b, 0, H - 1), cx = clampi2(lx + a, 0, W - 1);  // hypers: 'edge' (:172-174)
      const long long off = (long long)cy * W + cx;
      if (KIND == LERF_KIND_GAUSS) {
        float rho, s_x, s_y;
        hyp.gauss(tab, plane_sz, p, off, max_sigma, rho, s_x, s_y);
        m = fmax(m, gauss_exponent(rho, s_x, s_y, aa * dr[b], aa * dc[a]));
      } else {
        const double al = (double)hyp.alpha(tab, plane_sz, p, off, max_sigma);
        sum += lin_alpha(dr[b], al) * lin_alpha(dc[a], al);
      }
    }
  if (KIND == LERF_KIND_GAUSS) {
    for (int a = 0; a < supp; ++a)
      for (int b = 0; b < supp; ++b) {
        const int cy = clampi2(ly + b, 0, H - 1), cx = clampi2(lx + a, 0, W - 1);
        float rho, s_x, s_y;
        hyp.gauss(tab, plane_sz, p, (long long)cy * W + cx, max_sigma, rho, s_x, s_y);
        sum += exp(gauss_exponent(rho, s_x, s_y, aa * dr[b], aa * dc[a]) - m);
      }
  }
  double acc = 0.0;
  for (int a = 0; a < supp; ++a)
    for (int b = 0; b < supp; ++b) {
      const int cy = clampi2(ly + b, 0, H - 1), cx = clampi2(lx + a, 0, W - 1);
      const long long off = (long long)cy * W + cx;
      double w;
      if (KIND == LERF_KIND_GAUSS) {
        float rho, s_x, s_y;
        hyp.gauss(tab, plane_sz, p, off, max_sigma, rho, s_x, s_y);
        w = exp(gauss_exponent(rho, s_x, s_y, aa * dr[b], aa * dc[a]) - m);
      } else {
        const double al = (double)hyp.alpha(tab, plane_sz, p, off, max_sigma);
        w = lin_alpha(dr[b], al) * lin_alpha(dc[a], al);
      }
      const int iy = pad_src(ly + b, H, pad_mode), ix = pad_src(lx + a, W, pad_mode);  // image: np.pad(mode) (:208)
      const double v = (iy >= 0 && ix >= 0) ? (double)load_img(img + (long long)p * plane_sz + (long long)iy * W + ix) : 0.0;
      acc += v * (w / sum);  // 0/0 -> NaN exactly where numpy gives NaN
    }
  const long long ip = ((long long)p * oH + oy) * oW + ox;
  const long long ih = (((long long)(p / channels) * oH + oy) * oW + ox) * channels + (p % channels);
  store_sample<FMT>(out, 0, ip, ih, acc);
}

// ---------------------------------------------------------------------------------------------
// homographic warp: one thread = one output pixel, all planes
// ---------------------------------------------------------------------------------------------
struct WarpGeom {
  double m[9];  // inverse homography (output -> input), row-major
  int H, W, oH, oW;
  int pad0_y, pad0_x;            // support-2 leading pads (:363-369)
  int mpad0_y, mpad0_x, border;  // nearest-neighbour mask geometry (eval_lut_warp.py:197-204)
  int support, pad_mode;         // lerf_warp_ex: taps per axis and np.pad mode of the image (2 / 'constant' otherwise)
};

__constant__ double kEps32 = 1.1920928955078125e-07;

template <int KIND, int FMT, typename ImgT, typename Hyp>
__global__ void __launch_bounds__(256)
    warp_kernel(const ImgT* __restrict__ img, Hyp hyp, WarpGeom g, int planes, int channels, float max_sigma,
                void* __restrict__ out, uint8_t* __restrict__ mask) {
  __shared__ HyperTab tab;
  build_hyper_tab(tab, max_sigma, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y);
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= g.oW || oy >= g.oH) return;
  const int H = g.H, W = g.W;
  // get_projected_grid2d (:306-342): float32 output coords, inverse homography, divide, clip to [0, in]
  const double x = (double)(float)ox, y = (double)(float)oy;
  const double g0 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[0], x), __dmul_rn(g.m[1], y)), g.m[2]);
  const double g1 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[3], x), __dmul_rn(g.m[4], y)), g.m[5]);
  const double g2 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[6], x), __dmul_rn(g.m[7], y)), g.m[8]);
  const double pr0 = fmin(fmax(g1 / g2, 0.0), (double)H);  // row coordinate
  const double pc0 = fmin(fmax(g0 / g2, 0.0), (double)W);  // column coordinate

  if (mask) {  // NearestWarp2dNumpy (:460-467): support 1, box2d weight, white frame test
    const int fr = clampi2((int)ceil(pr0 - 0.5 - kEps32) + g.mpad0_y, 0, H - 1);
    const int fc = clampi2((int)ceil(pc0 - 0.5 - kEps32) + g.mpad0_x, 0, W - 1);
    const double dr = (pr0 + (double)g.mpad0_y) - (double)fr, dc = (pc0 + (double)g.mpad0_x) - (double)fc;
    const int sr = fr - g.mpad0_y, sc = fc - g.mpad0_x;
    const bool hit = (-1.0 <= dr && dr <= 1.0) && (-1.0 <= dc && dc <= 1.0) && sr >= g.border &&
                     sr < H - g.border && sc >= g.border && sc < W - g.border;
    mask[(long long)oy * g.oW + ox] = hit ? 1 : 0;
  }
  if (!out) return;

  if (g.support != 2 || g.pad_mode != 0) {  // non-default operator parameters (r2): any support, np.pad mode of the image
    const int supp = g.support;
    const int lr = (int)ceil(pr0 - 0.5 * (double)supp - kEps32) + g.pad0_y;
    const int lc = (int)ceil(pc0 - 0.5 * (double)supp - kEps32) + g.pad0_x;
    const double pr = pr0 + (double)g.pad0_y, pc = pc0 + (double)g.pad0_x;
    const long long plane_sz = (long long)H * W;
    for (int p = 0; p < planes; ++p) {
      double m = -INFINITY, sum = 0.0, acc = 0.0;
      for (int pass = (KIND == LERF_KIND_GAUSS ? 0 : 1); pass < 3; ++pass)  // Gaussian: max exponent, sum, products; linear: sum, products
        for (int a = 0; a < supp; ++a)
          for (int b = 0; b < supp; ++b) {
            const int fr = clampi2(lr + b, 0, H - 1), fc = clampi2(lc + a, 0, W - 1);  // clipped in padded coordinates (:397-398)
            const double dr = pr - (double)fr, dc = pc - (double)fc;
            const int sr = fr - g.pad0_y, sc = fc - g.pad0_x;
            const long long off = (long long)max(sr, 0) * W + max(sc, 0);  // hypers: 'edge'
            double w;
            if (KIND == LERF_KIND_GAUSS) {
              float rho, s_x, s_y;
              hyp.gauss(tab, plane_sz, p, off, max_sigma, rho, s_x, s_y);
              const double e = gauss_exponent(rho, s_x, s_y, dr, dc);
              if (pass == 0) { m = fmax(m, e); continue; }
              w = exp(e - m);
            } else {
              const double al = (double)hyp.alpha(tab, plane_sz, p, off, max_sigma);
              w = lin_alpha(dr, al) * lin_alpha(dc, al);
            }
            if (pass == 1) { sum += w; continue; }
            const int iy = pad_src(sr, H, g.pad_mode), ix = pad_src(sc, W, g.pad_mode);
            const double v = (iy >= 0 && ix >= 0) ? (double)load_img(img + (long long)p * plane_sz + (long long)iy * W + ix) : 0.0;
            acc += v * (w / sum);
          }
      const long long ip = ((long long)p * g.oH + oy) * g.oW + ox;
      const long long ih = (((long long)(p / channels) * g.oH + oy) * g.oW + ox) * channels + (p % channels);
      store_sample<FMT>(out, 0, ip, ih, acc);
    }
    return;
  }
  const int lr = (int)ceil(pr0 - 1.0 - kEps32) + g.pad0_y;  // :347-352, :366
  const int lc = (int)ceil(pc0 - 1.0 - kEps32) + g.pad0_x;
  const double pr = pr0 + (double)g.pad0_y, pc = pc0 + (double)g.pad0_x;  // :367
  int off[4];
  bool inside[4];
  double dr[2], dc[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    dr[k] = pr - (double)clampi2(lr + k, 0, H - 1);  // taps clipped in padded coordinates (:397-403)
    dc[k] = pc - (double)clampi2(lc + k, 0, W - 1);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int sr = clampi2(lr + b, 0, H - 1) - g.pad0_y, sc = clampi2(lc + a, 0, W - 1) - g.pad0_x;
      inside[a * 2 + b] = sr >= 0 && sc >= 0;
      off[a * 2 + b] = max(sr, 0) * W + max(sc, 0);
    }
  const long long plane_sz = (long long)H * W;
  for (int p = 0; p < planes; ++p) {
    double e[4];
    float v[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int t = a * 2 + b;
        if (KIND == LERF_KIND_GAUSS) {
          float rho, s_x, s_y;
          hyp.gauss(tab, plane_sz, p, off[t], max_sigma, rho, s_x, s_y);
          e[t] = gauss_exponent(rho, s_x, s_y, dr[b], dc[a]);
        } else {
          const double al = (double)hyp.alpha(tab, plane_sz, p, off[t], max_sigma);
          e[t] = lin_alpha(dr[b], al) * lin_alpha(dc[a], al);
        }
        v[t] = inside[t] ? load_img(img + (long long)p * plane_sz + off[t]) : 0.0f;
      }
    const double val = combine4<KIND>(e, v);
    const long long ip = ((long long)p * g.oH + oy) * g.oW + ox;
    const long long ih = (((long long)(p / channels) * g.oH + oy) * g.oW + ox) * channels + (p % channels);
    store_sample<FMT>(out, 0, ip, ih, val);
  }
}

template <int KIND, typename ImgT, typename Hyp>
static int launch_sr(const lerf_sr_plan_impl* P, const ImgT* img, Hyp hyp, int planes, int channels, float max_sigma,
                     int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  dim3 block(32, 8), grid((P->oW + 31) / 32, (oy1 - oy0 + 7) / 8, planes);
#define LERF_GO(F)                                                                                           \
  if (P->general)                                                                                            \
    resize_sr_support_kernel<KIND, F, ImgT, Hyp><<<grid, block, 0, st>>>(                                    \
        img, hyp, P->H, P->W, P->oH, P->oW, P->support, P->left_y, P->dist_y, P->left_x, P->dist_x,          \
        P->pad_mode, KIND == LERF_KIND_GAUSS ? P->aa_scale : 1.0, channels, max_sigma, oy0, oy1, out);       \
  else                                                                                                       \
    resize_sr_generic_kernel<KIND, F, ImgT, Hyp><<<grid, block, 0, st>>>(                                    \
        img, hyp, P->H, P->W, P->oH, P->oW, P->left_y, P->dist_y, P->left_x, P->dist_x, channels, max_sigma, \
        oy0, oy1, out)
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

template <int KIND, typename ImgT, typename Hyp>
static int launch_warp(const ImgT* img, Hyp hyp, const WarpGeom& g, int planes, int channels, float max_sigma,
                       void* out, int fmt, uint8_t* mask, cudaStream_t st) {
  dim3 block(32, 8), grid((g.oW + 31) / 32, (g.oH + 7) / 8, 1);
#define LERF_GO(F) warp_kernel<KIND, F, ImgT, Hyp><<<grid, block, 0, st>>>(img, hyp, g, planes, channels, max_sigma, out, mask)
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

// Periodic-geometry detection for integer scales: returns S (2,3,4,8) if out = S*in and the host tables repeat
// with period S (left exactly, distances to 1e-9), else 0.
static int detect_int_scale(int in, int out, const int* left, const double* dist, int& ph, double tab[8][2]) {
  if (in < 2 || out % in) return 0;
  const int S = out / in;
  if (S != 2 && S != 3 && S != 4 && S != 8) return 0;
  int o0 = -1;
  for (int o = 0; o < out && o0 < 0; ++o)
    if (left[o] == 0) o0 = o;
  if (o0 < 0 || o0 >= S || o0 + S > out) return 0;
  ph = o0;
  for (int m = 0; m < S; ++m)
    for (int k = 0; k < 2; ++k) tab[m][k] = dist[2 * (ph + m) + k];
  for (int o = 0; o < out; ++o) {
    const int d = o - ph;
    const int l = d >= 0 ? d / S : -((-d + S - 1) / S);
    const int m = d - l * S;
    if (left[o] != l) return 0;
    for (int k = 0; k < 2; ++k)
      if (fabs(dist[2 * o + k] - tab[m][k]) > 1e-9) return 0;
  }
  return S;
}

static int fill_geom(WarpGeom& g, int H, int W, int oH, int oW, const double minv[9], int pad0_y, int pad0_x,
                     int mpy, int mpx, int border) {
  if (!minv) return fail(LERF_EINVAL, "warp: minv is null");
  if (H < 1 || W < 1 || oH < 0 || oW < 0) return fail(LERF_EINVAL, "warp: bad sizes");
  if (pad0_y < 0 || pad0_x < 0 || mpy < 0 || mpx < 0) return fail(LERF_EINVAL, "warp: pads must be >= 0");
  for (int i = 0; i < 9; ++i) g.m[i] = minv[i];
  g.H = H; g.W = W; g.oH = oH; g.oW = oW;
  g.pad0_y = pad0_y; g.pad0_x = pad0_x; g.mpad0_y = mpy; g.mpad0_x = mpx; g.border = border;
  g.support = 2; g.pad_mode = 0;
  return LERF_OK;
}

}  // namespace lerf

using namespace lerf;


extern "C" {

int lerf_sr_plan_create(int H, int W, int oH, int oW, const int32_t* left_y, const double* dist_y,
                        const int32_t* left_x, const double* dist_x, int device, lerf_sr_plan_t** out) {
  if (!left_y || !dist_y || !left_x || !dist_x || !out) return fail(LERF_EINVAL, "lerf_sr_plan_create: null argument");
  if (H < 1 || W < 1 || oH < 1 || oW < 1) return fail(LERF_EINVAL, "lerf_sr_plan_create: bad sizes");
  for (int o = 0; o < oH; ++o)
    if (left_y[o] < -1 || left_y[o] > H - 1 || (o && left_y[o] < left_y[o - 1]))
      return fail(LERF_EINVAL, "lerf_sr_plan_create: left_y[%d]=%d outside [-1,%d] or not monotone", o, left_y[o], H - 1);
  for (int o = 0; o < oW; ++o)
    if (left_x[o] < -1 || left_x[o] > W - 1 || (o && left_x[o] < left_x[o - 1]))
      return fail(LERF_EINVAL, "lerf_sr_plan_create: left_x[%d]=%d outside [-1,%d] or not monotone", o, left_x[o], W - 1);
  LERF_CUDA(cudaSetDevice(device));
  lerf_sr_plan_impl* P = new lerf_sr_plan_impl();
  memset(P, 0, sizeof(*P));
  P->device = device; P->H = H; P->W = W; P->oH = oH; P->oW = oW;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** d, const void* h, size_t n) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(d, n);
    if (e == cudaSuccess) e = cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice);
  };
  up((void**)&P->left_y, left_y, sizeof(int) * oH);
  up((void**)&P->dist_y, dist_y, sizeof(double) * 2 * oH);
  up((void**)&P->left_x, left_x, sizeof(int) * oW);
  up((void**)&P->dist_x, dist_x, sizeof(double) * 2 * oW);
  if (e != cudaSuccess) {
    lerf_sr_plan_destroy(reinterpret_cast<lerf_sr_plan_t*>(P));
    return fail(LERF_ECUDA, "lerf_sr_plan_create: upload failed: %s", cudaGetErrorString(e));
  }
  P->tile_rows = 32;
  for (int R : {128, 96, 64}) {
    bool ok = true;
    for (int o = 0; o < oH && ok; ++o) ok = left_y[o + R - 1 < oH ? o + R - 1 : oH - 1] - left_y[o] <= 31;
    if (ok) {
      P->tile_rows = R;
      break;
    }
  }
  P->h_left_y = (int*)malloc(sizeof(int) * oH);
  memcpy(P->h_left_y, left_y, sizeof(int) * oH);
  const int sy = detect_int_scale(H, oH, left_y, dist_y, P->ph_y, P->ph_dist_y);
  const int sx = detect_int_scale(W, oW, left_x, dist_x, P->ph_x, P->ph_dist_x);
  P->int_scale = (sy && sy == sx) ? sy : 0;
  P->tile_ok = 1;  // resample_tile.cu: a block's 32 x 32 outputs must find their taps in a 33 x 33 input window,
                   // and the fixed-point exponent assumes |distance| <= 1 (+ eps)
  for (int o = 0; o < oH && P->tile_ok; ++o)
    if (left_y[o + 31 < oH ? o + 31 : oH - 1] - left_y[o] > 31 || fabs(dist_y[2 * o]) > 1.0000005 || fabs(dist_y[2 * o + 1]) > 1.0000005)
      P->tile_ok = 0;
  for (int o = 0; o < oW && P->tile_ok; ++o)
    if (left_x[o + 31 < oW ? o + 31 : oW - 1] - left_x[o] > 31 || fabs(dist_x[2 * o]) > 1.0000005 || fabs(dist_x[2 * o + 1]) > 1.0000005)
      P->tile_ok = 0;
  if (P->tile_ok) {  // cell runs for the cell kernel (resample_tile.cu): the outputs whose first tap is l are contiguous
    auto runs = [&](const int32_t* left, int in, int on, int** dev, int& longest) {
      std::vector<int> cs(in + 2);
      int o = 0;
      for (int l = -1; l <= in; ++l) {
        while (o < on && left[o] < l) ++o;
        cs[l + 1] = o;
      }
      longest = 0;
      for (int l = 0; l <= in; ++l) longest = std::max(longest, cs[l + 1] - cs[l]);
      up((void**)dev, cs.data(), sizeof(int) * (in + 2));
    };
    runs(left_y, H, oH, &P->cell_y, P->cell_max_y);
    runs(left_x, W, oW, &P->cell_x, P->cell_max_x);
    if (e != cudaSuccess) {
      lerf_sr_plan_destroy(reinterpret_cast<lerf_sr_plan_t*>(P));
      return fail(LERF_ECUDA, "lerf_sr_plan_create: upload failed: %s", cudaGetErrorString(e));
    }
  }
  *out = reinterpret_cast<lerf_sr_plan_t*>(P);
  return LERF_OK;
}

int lerf_sr_plan_create_ex(int H, int W, int oH, int oW, int support, const int32_t* left_y, const double* dist_y,
                           const int32_t* left_x, const double* dist_x, int pad_mode, double aa_scale, int device,
                           lerf_sr_plan_t** out) {
  if (!left_y || !dist_y || !left_x || !dist_x || !out) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: null argument");
  if (H < 1 || W < 1 || oH < 1 || oW < 1) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: bad sizes");
  if (support < 1 || support > 64) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: support %d outside [1,64]", support);
  if (pad_mode < LERF_PAD_CONSTANT || pad_mode > LERF_PAD_WRAP) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: unknown pad_mode %d", pad_mode);
  if (!(aa_scale > 0.0)) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: aa_scale must be positive");
  for (int o = 1; o < oH; ++o)
    if (left_y[o] < left_y[o - 1]) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: left_y not monotone at %d", o);
  for (int o = 1; o < oW; ++o)
    if (left_x[o] < left_x[o - 1]) return fail(LERF_EINVAL, "lerf_sr_plan_create_ex: left_x not monotone at %d", o);
  LERF_CUDA(cudaSetDevice(device));
  lerf_sr_plan_impl* P = new lerf_sr_plan_impl();
  memset(P, 0, sizeof(*P));
  P->device = device; P->H = H; P->W = W; P->oH = oH; P->oW = oW;
  P->general = 1; P->support = support; P->pad_mode = pad_mode; P->aa_scale = aa_scale;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** d, const void* h, size_t n) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(d, n);
    if (e == cudaSuccess) e = cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice);
  };
  up((void**)&P->left_y, left_y, sizeof(int) * oH);
  up((void**)&P->dist_y, dist_y, sizeof(double) * support * oH);
  up((void**)&P->left_x, left_x, sizeof(int) * oW);
  up((void**)&P->dist_x, dist_x, sizeof(double) * support * oW);
  if (e != cudaSuccess) {
    lerf_sr_plan_destroy(reinterpret_cast<lerf_sr_plan_t*>(P));
    return fail(LERF_ECUDA, "lerf_sr_plan_create_ex: upload failed: %s", cudaGetErrorString(e));
  }
  P->h_left_y = (int*)malloc(sizeof(int) * oH);
  memcpy(P->h_left_y, left_y, sizeof(int) * oH);
  P->tile_rows = 32;
  *out = reinterpret_cast<lerf_sr_plan_t*>(P);
  return LERF_OK;
}

void lerf_sr_plan_destroy(lerf_sr_plan_t* plan) {
  if (!plan) return;
  lerf_sr_plan_impl* P = reinterpret_cast<lerf_sr_plan_impl*>(plan);
  cudaSetDevice(P->device);
  cudaFree(P->left_y); cudaFree(P->dist_y); cudaFree(P->left_x); cudaFree(P->dist_x);
  cudaFree(P->cell_y); cudaFree(P->cell_x);
  for (int i = 0; i < P->coef_n; ++i) cudaFree(P->coef_dev[i]);
  free(P->h_left_y);
  delete P;
}

static int sr_args_ok(const char* who, int kind, const void* plan, const void* a, const void* b, const void* out,
                      int planes, int channels, int oy0, int oy1, int oH) {
  if (kind != LERF_KIND_GAUSS && kind != LERF_KIND_LINEAR) return fail(LERF_EINVAL, "%s: unknown kind %d", who, kind);
  if (!plan || !a || !b || !out) return fail(LERF_EINVAL, "%s: null pointer", who);
  if (planes < 0 || planes > 65535) return fail(LERF_EINVAL, "%s: planes=%d outside [0,65535]", who, planes);
  if (channels < 1 || (planes % channels)) return fail(LERF_EINVAL, "%s: planes=%d not a multiple of channels=%d", who, planes, channels);
  if (oy0 < 0 || oy1 > oH || oy0 > oy1) return fail(LERF_EINVAL, "%s: bad output band [%d,%d) of %d", who, oy0, oy1, oH);
  return LERF_OK;
}

int lerf_resize_sr(int kind, const lerf_sr_plan_t* plan, const uint8_t* feat, const uint8_t* codes, int planes,
                   int channels, float max_sigma, int oy0, int oy1, void* out, int out_format, lerf_stream_t stream) {
  const lerf_sr_plan_impl* P = reinterpret_cast<const lerf_sr_plan_impl*>(plan);
  int rc = sr_args_ok("lerf_resize_sr", kind, plan, feat, codes, out, planes, channels, oy0, oy1, P ? P->oH : 0);
  if (rc) return rc;
  if (planes == 0 || oy0 == oy1) return LERF_OK;
  CodeSrc hyp{codes};
  if (P->general) {  // non-default support / pad mode / antialias: the float64 support kernel
    if (kind == LERF_KIND_GAUSS)
      return launch_sr<LERF_KIND_GAUSS>(P, feat, hyp, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
    return launch_sr<LERF_KIND_LINEAR>(P, feat, hyp, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
  }
  if (kind == LERF_KIND_GAUSS) {
    if (P->int_scale && g_dbg.force_generic == 0) {  // periodic geometry: integer-scale cell-owner kernel (resample_int.cu)
      rc = resize_sr_int_gauss(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
      if (rc != -1) return rc;
    }
  }
  if (kind == LERF_KIND_LINEAR && P->int_scale && g_dbg.force_generic == 0) {  // periodic geometry, LeRF-L (resample_int.cu)
    rc = resize_sr_int_linear(P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  if (g_dbg.force_generic == 0 || g_dbg.force_generic == 3) {  // scales in [3, 4] per axis: any-scale cell kernel (resample_tile.cu)
    rc = resize_sr_cell(kind, P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, out_format, g_dbg.force_generic == 3,
                        (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  if (g_dbg.force_generic != 1) {  // any scale >= 1: tile kernel (resample_tile.cu); the float64 kernels below are the parity path
    rc = resize_sr_tile(kind, P, feat, codes, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  if (kind == LERF_KIND_GAUSS) {
    return launch_sr<LERF_KIND_GAUSS>(P, feat, hyp, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
  }
  return launch_sr<LERF_KIND_LINEAR>(P, feat, hyp, planes, channels, max_sigma, oy0, oy1, out, out_format, (cudaStream_t)stream);
}

/* Testing hook: route integer scales through the generic kernel too (parity tests compare both). */
void lerf_debug_force_generic(int on) { g_dbg.force_generic = on; }

/* Testing hook: 0 = the fast warp kernel reads img/codes + tables, 1 (default) = per-sample records (resample_tile.cu). */
void lerf_debug_warp_records(int on) { g_dbg.warp_records = on != 0; }

/* Testing / tuning hook for the integer-scale kernel (see lerf_b200.h). */
void lerf_debug_resize_variant(int variant) {  // 10: production arithmetic with the byte-store uint8 epilogue of r1
  g_dbg.u8_staged = variant != 10;
  g_dbg.resize_variant = variant == 10 ? 0 : variant;
}

int lerf_resize_sr_f32(int kind, const lerf_sr_plan_t* plan, const float* img, const float* h0, const float* h1,
                       const float* h2, int planes, float max_sigma, float* out, lerf_stream_t stream) {
  const lerf_sr_plan_impl* P = reinterpret_cast<const lerf_sr_plan_impl*>(plan);
  if (kind == LERF_KIND_LINEAR) { h1 = h0; h2 = h0; }
  int rc = sr_args_ok("lerf_resize_sr_f32", kind, plan, img, h0, out, planes, 1, 0, P ? P->oH : 0, P ? P->oH : 0);
  if (rc) return rc;
  if (!h1 || !h2) return fail(LERF_EINVAL, "lerf_resize_sr_f32: null hyper plane");
  if (planes == 0) return LERF_OK;
  FloatSrc hyp{h0, h1, h2};
  if (kind == LERF_KIND_GAUSS)
    return launch_sr<LERF_KIND_GAUSS>(P, img, hyp, planes, 1, max_sigma, 0, P->oH, out, LERF_OUT_F32, (cudaStream_t)stream);
  return launch_sr<LERF_KIND_LINEAR>(P, img, hyp, planes, 1, max_sigma, 0, P->oH, out, LERF_OUT_F32, (cudaStream_t)stream);
}

int lerf_warp(int kind, const uint8_t* feat, const uint8_t* codes, int planes, int channels, int H, int W, int oH,
              int oW, const double minv[9], int pad0_y, int pad0_x, float max_sigma, void* out, int out_format,
              uint8_t* mask, int mask_pad0_y, int mask_pad0_x, int mask_border, lerf_stream_t stream) {
  if (kind != LERF_KIND_GAUSS && kind != LERF_KIND_LINEAR) return fail(LERF_EINVAL, "lerf_warp: unknown kind %d", kind);
  if (out && (!feat || !codes)) return fail(LERF_EINVAL, "lerf_warp: null input");
  if (!out && !mask) return fail(LERF_EINVAL, "lerf_warp: neither out nor mask requested");
  if (planes < 0 || channels < 1 || (planes % channels)) return fail(LERF_EINVAL, "lerf_warp: bad planes/channels");
  WarpGeom g;
  int rc = fill_geom(g, H, W, oH, oW, minv, pad0_y, pad0_x, mask_pad0_y, mask_pad0_x, mask_border);
  if (rc) return rc;
  if (oH == 0 || oW == 0) return LERF_OK;
  if (g_dbg.force_generic != 1) {  // fast path (resample_tile.cu); the float64 kernel below is the parity path
    rc = warp_fast(kind, feat, codes, planes, channels, H, W, oH, oW, minv, pad0_y, pad0_x, mask_pad0_y, mask_pad0_x,
                   mask_border, max_sigma, out, out_format, mask, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  CodeSrc hyp{codes};
  if (kind == LERF_KIND_GAUSS)
    return launch_warp<LERF_KIND_GAUSS>(feat, hyp, g, planes, channels, max_sigma, out, out_format, mask, (cudaStream_t)stream);
  return launch_warp<LERF_KIND_LINEAR>(feat, hyp, g, planes, channels, max_sigma, out, out_format, mask, (cudaStream_t)stream);
}

int lerf_warp_f32(int kind, const float* img, const float* h0, const float* h1, const float* h2, int planes, int H,
                  int W, int oH, int oW, const double minv[9], int pad0_y, int pad0_x, float max_sigma, float* out,
                  lerf_stream_t stream) {
  if (kind != LERF_KIND_GAUSS && kind != LERF_KIND_LINEAR) return fail(LERF_EINVAL, "lerf_warp_f32: unknown kind %d", kind);
  if (kind == LERF_KIND_LINEAR) { h1 = h0; h2 = h0; }
  if (!img || !h0 || !h1 || !h2 || !out) return fail(LERF_EINVAL, "lerf_warp_f32: null pointer");
  if (planes < 0) return fail(LERF_EINVAL, "lerf_warp_f32: bad planes");
  WarpGeom g;
  int rc = fill_geom(g, H, W, oH, oW, minv, pad0_y, pad0_x, 0, 0, 0);
  if (rc) return rc;
  if (oH == 0 || oW == 0 || planes == 0) return LERF_OK;
  FloatSrc hyp{h0, h1, h2};
  if (kind == LERF_KIND_GAUSS)
    return launch_warp<LERF_KIND_GAUSS>(img, hyp, g, planes, 1, max_sigma, out, LERF_OUT_F32, nullptr, (cudaStream_t)stream);
  return launch_warp<LERF_KIND_LINEAR>(img, hyp, g, planes, 1, max_sigma, out, LERF_OUT_F32, nullptr, (cudaStream_t)stream);
}

int lerf_warp_ex(int kind, const uint8_t* feat, const uint8_t* codes, const float* img, const float* h0, const float* h1,
                 const float* h2, int planes, int channels, int H, int W, int oH, int oW, const double minv[9], int support,
                 int pad_mode, int pad0_y, int pad0_x, float max_sigma, void* out, int out_format, lerf_stream_t stream) {
  if (kind != LERF_KIND_GAUSS && kind != LERF_KIND_LINEAR) return fail(LERF_EINVAL, "lerf_warp_ex: unknown kind %d", kind);
  const bool u8 = feat != nullptr;
  if (u8 ? !codes : (!img || !h0)) return fail(LERF_EINVAL, "lerf_warp_ex: give feat + codes (uint8) or img + h0[, h1, h2] (float32)");
  if (kind == LERF_KIND_LINEAR) { h1 = h0; h2 = h0; }
  if (!u8 && (!h1 || !h2)) return fail(LERF_EINVAL, "lerf_warp_ex: null hyper plane");
  if (!out) return fail(LERF_EINVAL, "lerf_warp_ex: null output");
  if (support < 1 || support > 64) return fail(LERF_EINVAL, "lerf_warp_ex: support %d outside [1,64]", support);
  if (pad_mode < LERF_PAD_CONSTANT || pad_mode > LERF_PAD_WRAP) return fail(LERF_EINVAL, "lerf_warp_ex: unknown pad_mode %d", pad_mode);
  if (planes < 0 || channels < 1 || (planes % channels)) return fail(LERF_EINVAL, "lerf_warp_ex: bad planes/channels");
  if (!u8 && out_format != LERF_OUT_F32) return fail(LERF_EINVAL, "lerf_warp_ex: float32 inputs give float32 output");
  WarpGeom g;
  int rc = fill_geom(g, H, W, oH, oW, minv, pad0_y, pad0_x, 0, 0, 0);
  if (rc) return rc;
  if (oH == 0 || oW == 0 || planes == 0) return LERF_OK;
  g.support = support;
  g.pad_mode = pad_mode;
  if (u8) {
    CodeSrc hyp{codes};
    if (kind == LERF_KIND_GAUSS) return launch_warp<LERF_KIND_GAUSS>(feat, hyp, g, planes, channels, max_sigma, out, out_format, nullptr, (cudaStream_t)stream);
    return launch_warp<LERF_KIND_LINEAR>(feat, hyp, g, planes, channels, max_sigma, out, out_format, nullptr, (cudaStream_t)stream);
  }
  FloatSrc hyp{h0, h1, h2};
  if (kind == LERF_KIND_GAUSS) return launch_warp<LERF_KIND_GAUSS>(img, hyp, g, planes, 1, max_sigma, out, LERF_OUT_F32, nullptr, (cudaStream_t)stream);
  return launch_warp<LERF_KIND_LINEAR>(img, hyp, g, planes, 1, max_sigma, out, LERF_OUT_F32, nullptr, (cudaStream_t)stream);
}

}  // extern "C"

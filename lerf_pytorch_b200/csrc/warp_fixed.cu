// Fixed-kernel homographic warps: nearest / bilinear / bicubic / Lanczos-2 / Lanczos-3.
// Reference being replaced (ddlee-cn/LeRF-PyTorch): resize_right/resize_right2d_numpy.py:451-494
// (Bicubic/Nearest/Bilinear/Lanczos2/Lanczos3Warp2dNumpy) = Warp2dNumpy.warp :409-449 on the geometry of :292-407 with
// the separable kernels of resize_right/interp_methods.py:32-100.  These are the baselines the reference compares LeRF
// against; they share the geometry code path of lerf_warp.
//
// One thread = one output pixel, all planes.  The inverse homography, the tap positions (clipped in padded coordinates)
// and the 2 x SUPP kernel values are float64 in the reference's operation order; the kernel is separable, so the
// normalised S x S patch sum is (sum_j wr_j (sum_i wc_i v_ji)) / (sum wr * sum wc).  0/0 gives NaN where numpy does.
#include <math.h>

#include "common.cuh"

namespace lerf {
namespace {

__constant__ double kEps32w = 1.1920928955078125e-07;
constexpr double kPi = 3.141592653589793;

struct FixedGeom {
  double m[9];
  int H, W, oH, oW, pad0_y, pad0_x;
};

template <int KERN>
__device__ __forceinline__ double kern(double x) {
  if (KERN == LERF_WARP_NEAREST) return ((-1.0 <= x && x < 0.0) ? 1.0 : 0.0) + ((0.0 <= x && x <= 1.0) ? 1.0 : 0.0);
  if (KERN == LERF_WARP_BILINEAR)
    return (x + 1.0) * ((-1.0 <= x && x < 0.0) ? 1.0 : 0.0) + (1.0 - x) * ((0.0 <= x && x <= 1.0) ? 1.0 : 0.0);
  if (KERN == LERF_WARP_BICUBIC) {
    const double a = fabs(x), a2 = a * a, a3 = a * a * a;
    double r = 0.0;
    if (a <= 1.0) r += 1.5 * a3 - 2.5 * a2 + 1.0;
    if (1.0 < a && a <= 2.0) r += -0.5 * a3 + 2.5 * a2 - 4.0 * a + 2.0;
    return r;
  }
  const double n = KERN == LERF_WARP_LANCZOS2 ? 2.0 : 3.0;
  const double v = (sin(kPi * x) * sin(kPi * x / n) + kEps32w) / ((kPi * kPi * x * x / n) + kEps32w);
  return fabs(x) < n ? v : 0.0;
}

template <int SUPP, int KERN, typename ImgT>
__global__ void __launch_bounds__(256)
    warp_fixed_kernel(const ImgT* __restrict__ img, const FixedGeom g, int planes, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= g.oW || oy >= g.oH) return;
  const int H = g.H, W = g.W;
  // get_projected_grid2d (:306-342): float32 output coords, inverse homography, divide, clip to [0, in]
  const double x = (double)(float)ox, y = (double)(float)oy;
  const double g0 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[0], x), __dmul_rn(g.m[1], y)), g.m[2]);
  const double g1 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[3], x), __dmul_rn(g.m[4], y)), g.m[5]);
  const double g2 = __dadd_rn(__dadd_rn(__dmul_rn(g.m[6], x), __dmul_rn(g.m[7], y)), g.m[8]);
  const double pr0 = fmin(fmax(g1 / g2, 0.0), (double)H);
  const double pc0 = fmin(fmax(g0 / g2, 0.0), (double)W);
  const int lr = (int)ceil(pr0 - 0.5 * SUPP - kEps32w) + g.pad0_y;  // :347-352, :366
  const int lc = (int)ceil(pc0 - 0.5 * SUPP - kEps32w) + g.pad0_x;
  const double pr = pr0 + (double)g.pad0_y, pc = pc0 + (double)g.pad0_x;  // :367
  double wr[SUPP], wc[SUPP];
  int sr[SUPP], sc[SUPP];
  double sum_r = 0.0, sum_c = 0.0;
#pragma unroll
  for (int k = 0; k < SUPP; ++k) {
    const int fr = min(max(lr + k, 0), H - 1), fc = min(max(lc + k, 0), W - 1);  // clipped in padded coordinates (:397-398)
    wr[k] = kern<KERN>(pr - (double)fr);
    wc[k] = kern<KERN>(pc - (double)fc);
    sum_r += wr[k];
    sum_c += wc[k];
    sr[k] = fr - g.pad0_y;  // < 0: inside the zero pad (:433)
    sc[k] = fc - g.pad0_x;
  }
  const double den = sum_r * sum_c;
  const long long plane_sz = (long long)H * W;
  for (int p = 0; p < planes; ++p) {
    const ImgT* ip = img + (long long)p * plane_sz;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < SUPP; ++j) {
      if (sr[j] < 0) continue;
      double row = 0.0;
#pragma unroll
      for (int i = 0; i < SUPP; ++i)
        if (sc[i] >= 0) row = fma(wc[i], (double)__ldg(ip + (long long)sr[j] * W + sc[i]), row);
      acc = fma(wr[j], row, acc);
    }
    out[((long long)p * g.oH + oy) * g.oW + ox] = (float)(acc / den);
  }
}

template <typename ImgT>
int launch_fixed(int kernel, const ImgT* img, const FixedGeom& g, int planes, float* out, cudaStream_t st) {
  dim3 block(32, 8), grid((g.oW + 31) / 32, (g.oH + 7) / 8, 1);
  switch (kernel) {
    case LERF_WARP_NEAREST: warp_fixed_kernel<1, LERF_WARP_NEAREST, ImgT><<<grid, block, 0, st>>>(img, g, planes, out); break;
    case LERF_WARP_BILINEAR: warp_fixed_kernel<2, LERF_WARP_BILINEAR, ImgT><<<grid, block, 0, st>>>(img, g, planes, out); break;
    case LERF_WARP_BICUBIC: warp_fixed_kernel<4, LERF_WARP_BICUBIC, ImgT><<<grid, block, 0, st>>>(img, g, planes, out); break;
    case LERF_WARP_LANCZOS2: warp_fixed_kernel<4, LERF_WARP_LANCZOS2, ImgT><<<grid, block, 0, st>>>(img, g, planes, out); break;
    case LERF_WARP_LANCZOS3: warp_fixed_kernel<6, LERF_WARP_LANCZOS3, ImgT><<<grid, block, 0, st>>>(img, g, planes, out); break;
    default: return fail(LERF_EINVAL, "lerf_warp_fixed: unknown kernel %d", kernel);
  }
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // namespace
}  // namespace lerf

using namespace lerf;

extern "C" {

int lerf_warp_fixed_support(int kernel) {
  switch (kernel) {
    case LERF_WARP_NEAREST: return 1;
    case LERF_WARP_BILINEAR: return 2;
    case LERF_WARP_BICUBIC: return 4;
    case LERF_WARP_LANCZOS2: return 4;
    case LERF_WARP_LANCZOS3: return 6;
    default: return -1;
  }
}

int lerf_warp_fixed(int kernel, const void* img, int img_is_u8, int planes, int H, int W, int oH, int oW, const double minv[9],
                    int pad0_y, int pad0_x, float* out, lerf_stream_t stream) {
  if (lerf_warp_fixed_support(kernel) < 0) return fail(LERF_EINVAL, "lerf_warp_fixed: unknown kernel %d", kernel);
  if (!img || !out || !minv) return fail(LERF_EINVAL, "lerf_warp_fixed: null pointer");
  if (planes < 0 || H < 1 || W < 1 || oH < 0 || oW < 0 || pad0_y < 0 || pad0_x < 0)
    return fail(LERF_EINVAL, "lerf_warp_fixed: bad sizes");
  if (planes == 0 || oH == 0 || oW == 0) return LERF_OK;
  FixedGeom g;
  for (int i = 0; i < 9; ++i) g.m[i] = minv[i];
  g.H = H; g.W = W; g.oH = oH; g.oW = oW; g.pad0_y = pad0_y; g.pad0_x = pad0_x;
  if (img_is_u8) return launch_fixed<uint8_t>(kernel, (const uint8_t*)img, g, planes, out, (cudaStream_t)stream);
  return launch_fixed<float>(kernel, (const float*)img, g, planes, out, (cudaStream_t)stream);
}

}  // extern "C"

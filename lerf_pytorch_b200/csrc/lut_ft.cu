// LUT fine-tuning operators: forward and backward of the two differentiable pieces the reference
// trains through when it fine-tunes its LUTs (resample/train_model.py --lutft):
//   * SWF2LUT.InterpTorchBatch, resample/model.py:172-385 -- 4-simplex interpolation of a TRAINABLE float table at
//     integer-valued pixels; the gradient flows to the table only (the pixels enter through floor_divide / %).
//   * SteeringGaussianResize2dTorch.resize, resize_right/resize_right2d_torch.py:140-197 -- the steerable Gaussian
//     resampler; gradients flow to the image and to the three hyper-parameter maps.
// Both are float32 like the reference's torch path.  The backward kernels scatter with atomicAdd (float), so gradients
// are reproducible to rounding, not bit for bit -- as with torch's own index_put / gather backward on CUDA.
//
// One deliberate option: model.py:229-243 reads the LSBs of modes c and t from the pixels of mode y
// ((1,1), (1,2), (2,1) instead of the mode's own taps).  `lsb_like_reference` = 1 reproduces that
// (what a checkpoint fine-tuned with the reference saw), 0 uses the mode's own taps (what the inference path
// FourSimplexInterpFaster, eval_lut_sr.py:24-470, does).
#include "common.cuh"

namespace lerf {
namespace {

struct FtTaps {
  int mi[4], mj[4];  // taps whose MSBs index the table
  int li[4], lj[4];  // taps whose LSBs weight the simplex
};

__device__ __forceinline__ void simplex_ft(const float* __restrict__ pl, int wp, int i, int j, const FtTaps& tp, int idx[5], float w[5]) {
  int msb[4], key[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int vm = (int)pl[(i + tp.mi[k]) * wp + j + tp.mj[k]];
    const int vl = (int)pl[(i + tp.li[k]) * wp + j + tp.lj[k]];
    msb[k] = vm >> 4;
    key[k] = ((vl & 15) << 4) | (3 - k);  // sort descending by lsb; ties by tap order (tied vertices have weight 0)
  }
  int t;
#define LERF_CE(a, b) if (key[a] < key[b]) { t = key[a]; key[a] = key[b]; key[b] = t; }
  LERF_CE(0, 1) LERF_CE(2, 3) LERF_CE(0, 2) LERF_CE(1, 3) LERF_CE(1, 2)
#undef LERF_CE
  const int strides[4] = {kStrideA, kStrideB, kStrideC, 1};
  int base = ((msb[0] * kL + msb[1]) * kL + msb[2]) * kL + msb[3];
  int f[5];
  f[4] = 0;
  idx[0] = base;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[k] = key[k] >> 4;
    base += strides[3 - (key[k] & 3)];
    idx[k + 1] = base;
  }
  w[0] = (float)(16 - f[0]);
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k + 1] = (float)(f[k] - f[k + 1]);
}

// out[p][c][i][j] = sum_k w_k * weight[idx_k][c] / 16
__global__ void lut_ft_forward_kernel(const float* __restrict__ weight, int oC, const float* __restrict__ img, int h, int w,
                                      int hp, int wp, FtTaps tp, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y, p = blockIdx.z;
  if (i >= h || j >= w) return;
  int idx[5];
  float wt[5];
  simplex_ft(img + (long long)p * hp * wp, wp, i, j, tp, idx, wt);
  for (int c = 0; c < oC; ++c) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s = fmaf(wt[k], __ldg(weight + (long long)idx[k] * oC + c), s);
    out[(((long long)p * oC + c) * h + i) * w + j] = s * 0.0625f;
  }
}

// grad_weight[idx_k][c] += w_k / 16 * grad_out[p][c][i][j]
__global__ void lut_ft_backward_kernel(const float* __restrict__ gout, int oC, const float* __restrict__ img, int h, int w,
                                       int hp, int wp, FtTaps tp, float* __restrict__ gweight) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y, p = blockIdx.z;
  if (i >= h || j >= w) return;
  int idx[5];
  float wt[5];
  simplex_ft(img + (long long)p * hp * wp, wp, i, j, tp, idx, wt);
  for (int c = 0; c < oC; ++c) {
    const float g = gout[(((long long)p * oC + c) * h + i) * w + j] * 0.0625f;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (wt[k] != 0.0f) atomicAdd(gweight + (long long)idx[k] * oC + c, wt[k] * g);
  }
}

bool ft_taps(char mode, int lsb_like_reference, FtTaps& t, int& pad) {
  static const int S[2][4] = {{0, 0, 1, 1}, {0, 1, 0, 1}};
  static const int D[2][4] = {{0, 0, 2, 2}, {0, 2, 0, 2}};
  static const int Y[2][4] = {{0, 1, 1, 2}, {0, 1, 2, 1}};
  static const int Cm[2][4] = {{0, 0, 0, 0}, {0, 1, 2, 3}};
  static const int T[2][4] = {{0, 1, 2, 3}, {0, 1, 2, 3}};
  const int(*m)[4];
  switch (mode) {  // model.py:186-245 (= eval_lut_sr.py:30-81 for the MSB taps)
    case 's': m = S; pad = 1; break;
    case 'd': m = D; pad = 2; break;
    case 'y': m = Y; pad = 2; break;
    case 'c': m = Cm; pad = 3; break;
    case 't': m = T; pad = 3; break;
    default: return false;
  }
  const bool bug = lsb_like_reference && (mode == 'c' || mode == 't');  // model.py:229-232, :240-243
  for (int k = 0; k < 4; ++k) {
    t.mi[k] = m[0][k];
    t.mj[k] = m[1][k];
    t.li[k] = bug ? Y[0][k] : m[0][k];
    t.lj[k] = bug ? Y[1][k] : m[1][k];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// steerable Gaussian SR resampling, float32, backward
// ---------------------------------------------------------------------------------------------------------------
__global__ void resize_gauss_backward_kernel(const float* __restrict__ img, const float* __restrict__ h0, const float* __restrict__ h1,
                                             const float* __restrict__ h2, int H, int W, int oH, int oW, const int* __restrict__ left_y,
                                             const int* __restrict__ left_x, const double* __restrict__ dist_y,
                                             const double* __restrict__ dist_x, float max_sigma, const float* __restrict__ gout,
                                             float* __restrict__ gimg, float* __restrict__ g0, float* __restrict__ g1,
                                             float* __restrict__ g2) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y, p = blockIdx.z;
  if (ox >= oW || oy >= oH) return;
  const long long pl = (long long)p * H * W;
  const int ly = left_y[oy], lx = left_x[ox];
  float wt[4], v[4], X[4], Y[4], rho[4], dxs[4], dys[4];
  long long hoff[4], ioff[4];
  float W_ = 0.0f, num = 0.0f, emax = -INFINITY, e[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int b = t >> 1, a = t & 1;  // tap row ly + b, column lx + a
    const int sy = ly + b, sx = lx + a;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'replicate' (:165-167)
    hoff[t] = pl + (long long)cy * W + cx;
    const bool inside = sy == cy && sx == cx;
    ioff[t] = inside ? hoff[t] : -1;
    v[t] = inside ? img[hoff[t]] : 0.0f;  // image: 'constant' 0 (:189)
    rho[t] = h0[hoff[t]] * 2.0f - 1.0f;
    const float sgx = h1[hoff[t]] * max_sigma, sgy = h2[hoff[t]] * max_sigma;
    dxs[t] = (float)dist_y[2 * oy + b];  // "x" of sk_weight is the ROW distance (:145-153)
    dys[t] = (float)dist_x[2 * ox + a];
    X[t] = sgx * dxs[t];
    Y[t] = sgy * dys[t];
    e[t] = -0.5f * (X[t] * X[t] - 2.0f * rho[t] * X[t] * Y[t] + Y[t] * Y[t]);
    emax = fmaxf(emax, e[t]);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    wt[t] = __expf(e[t] - emax);  // the common factor exp(emax) cancels in every normalised quantity below
    W_ += wt[t];
    num = fmaf(wt[t], v[t], num);
  }
  const float out = num / W_;
  const float g = gout[((long long)p * oH + oy) * oW + ox] / W_;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (gimg && ioff[t] >= 0) atomicAdd(gimg + ioff[t], g * wt[t]);
    const float gw = g * (v[t] - out) * wt[t];  // dL/dw_t * w_t (every d w_t / d hyper carries a factor w_t)
    if (g0) atomicAdd(g0 + hoff[t], gw * X[t] * Y[t] * 2.0f);                                   // rho = 2 h0 - 1
    if (g1) atomicAdd(g1 + hoff[t], gw * dxs[t] * (rho[t] * Y[t] - X[t]) * max_sigma);          // sigma_x = max_sigma h1
    if (g2) atomicAdd(g2 + hoff[t], gw * dys[t] * (rho[t] * X[t] - Y[t]) * max_sigma);          // sigma_y = max_sigma h2
  }
}

}  // namespace
}  // namespace lerf

using namespace lerf;

extern "C" {

int lerf_lut_ft_forward(const float* weight, int oC, const float* img, int planes, int h, int w, char mode,
                        int lsb_like_reference, float* out, lerf_stream_t stream) {
  FtTaps tp;
  int pad;
  if (!ft_taps(mode, lsb_like_reference, tp, pad)) return fail(LERF_EINVAL, "Mode %c not implemented.", mode);
  if (!weight || !img || !out) return fail(LERF_EINVAL, "lerf_lut_ft_forward: null pointer");
  if (oC < 1 || planes < 0 || h < 0 || w < 0) return fail(LERF_EINVAL, "lerf_lut_ft_forward: bad sizes");
  if (planes == 0 || h == 0 || w == 0) return LERF_OK;
  if (planes > 65535) return fail(LERF_EINVAL, "lerf_lut_ft_forward: more than 65535 planes");
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, planes);
  lut_ft_forward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(weight, oC, img, h, w, h + pad, w + pad, tp, out);
  LERF_LAUNCHED();
  return LERF_OK;
}

int lerf_lut_ft_backward(const float* grad_out, int oC, const float* img, int planes, int h, int w, char mode,
                         int lsb_like_reference, float* grad_weight, lerf_stream_t stream) {
  FtTaps tp;
  int pad;
  if (!ft_taps(mode, lsb_like_reference, tp, pad)) return fail(LERF_EINVAL, "Mode %c not implemented.", mode);
  if (!grad_out || !img || !grad_weight) return fail(LERF_EINVAL, "lerf_lut_ft_backward: null pointer");
  if (oC < 1 || planes < 0 || h < 0 || w < 0) return fail(LERF_EINVAL, "lerf_lut_ft_backward: bad sizes");
  if (planes == 0 || h == 0 || w == 0) return LERF_OK;
  if (planes > 65535) return fail(LERF_EINVAL, "lerf_lut_ft_backward: more than 65535 planes");
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, planes);
  lut_ft_backward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(grad_out, oC, img, h, w, h + pad, w + pad, tp, grad_weight);
  LERF_LAUNCHED();
  return LERF_OK;
}

int lerf_resize_sr_f32_backward(int kind, const lerf_sr_plan_t* plan, const float* img, const float* h0, const float* h1,
                                const float* h2, int planes, float max_sigma, const float* grad_out, float* grad_img,
                                float* grad_h0, float* grad_h1, float* grad_h2, lerf_stream_t stream) {
  if (kind != LERF_KIND_GAUSS) return fail(LERF_EUNSUPPORTED, "lerf_resize_sr_f32_backward: only LERF_KIND_GAUSS has a backward");
  if (!plan || !img || !h0 || !h1 || !h2 || !grad_out) return fail(LERF_EINVAL, "lerf_resize_sr_f32_backward: null pointer");
  const lerf_sr_plan_impl* P = reinterpret_cast<const lerf_sr_plan_impl*>(plan);
  if (P->general) return fail(LERF_EUNSUPPORTED, "lerf_resize_sr_f32_backward: default operator parameters only (support 2, 'constant' pad)");
  if (planes < 0) return fail(LERF_EINVAL, "lerf_resize_sr_f32_backward: bad planes");
  if (planes == 0 || P->oH == 0 || P->oW == 0) return LERF_OK;
  if (planes > 65535) return fail(LERF_EINVAL, "lerf_resize_sr_f32_backward: more than 65535 planes");
  dim3 block(32, 8), grid((P->oW + 31) / 32, (P->oH + 7) / 8, planes);
  resize_gauss_backward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(img, h0, h1, h2, P->H, P->W, P->oH, P->oW, P->left_y,
                                                                         P->left_x, P->dist_y, P->dist_x, max_sigma, grad_out,
                                                                         grad_img, grad_h0, grad_h1, grad_h2);
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // extern "C"

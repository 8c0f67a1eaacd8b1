// Fast resamplers for the geometries the cell-owner kernel (resample_int.cu) does not cover, uint8 feat + hyper codes in:
//   * resize_sr_tile_kernel : arbitrary (non-integer, anisotropic) scale >= 1 SR, Gaussian (LeRF-G) or amplified linear
//                             (LeRF-L; BASELINE.json cfg-2)
//   * warp_fast_kernel      : homographic warp, both kinds (cfg-4)
// Reference being replaced (ddlee-cn/LeRF-PyTorch, resize_right/resize_right2d_numpy.py):
//   SteeringGaussianResize2dNumpy.resize :162-223   AmplifiedLinearResize2dNumpy.resize :243-282
//   SteeringGaussianWarp2dNumpy.warp     :516-577   AmplifiedLinearWarp2dNumpy.warp     :597-635   geometry :292-407
//
// Same arithmetic as resample_int.cuh: per-tap coefficients are exact float64 promotions of the reference's float32
// hyper values (per-code tables), the exponent is evaluated in float64 on top of the rounding constant of an unsigned
// fixed point (SR kernels: column terms hoisted with the taps, the ROWQ order of resample_int.cuh, three roundings at the
// fixed-point scale; warp: magic-number add last, one rounding), the smallest of the four becomes weight 1, only differences go through ex2.approx, and
// the output is v0 + sum w_t (v_t - v0) / sum w_t with exact integer differences.  The exact-operation-order float64
// kernels of resample.cu stay as the parity path (float32-hyper API, lerf_debug_force_generic).
//
// SR tile kernel: a block owns 32 x 32 outputs of one plane (4 rows per thread).  For a scale >= 1 the first tap advances
// by at most one input sample per output, so the block's taps lie in a window of at most 33 x 33 input samples: their
// coefficients are decoded once into shared memory (the float64 kernel decodes them once per output sample and tap).
#include <math.h>

#include <mutex>
#include <type_traits>

#include "resample_int.cuh"

namespace lerf {

using namespace rsi;

namespace {

constexpr int kOX = 32;             // output columns per block; 256 threads, thread (tx, ty) owns rows ty + 8 j of the block
constexpr int kTY = 8;
constexpr int kWX = 33, kWY = 33;   // input window capacity; a block covers as many output rows (32 .. 128, plan->tile_rows)
                                    // as keep its taps inside 33 input rows, so the per-block set-up is amortised

struct FixQ {
  int fb;           // fraction bits (host only)
  double magic;     // 2^(52-FB) + 16 * 2^-FB
  float neg_scale;  // -2^-FB
  double dlim;      // warp: distances are clamped to +-dlim for the exponent (only outside the validity mask)
};

FixQ make_fixq(float max_sigma) {
  FixQ q;
  q.dlim = 1.0 + 1.0 / 1048576.0;
  const double reach = (double)max_sigma * 2.0 * q.dlim;  // |dr|, |dc| <= 1 (+ eps): taps are the two samples around p
  const double bound = 0.5 * kLog2e * reach * reach + 2.0;
  int fb = 26;
  while (fb > 8 && bound * (double)(1u << fb) >= 4294967000.0) --fb;
  q.fb = fb;
  q.magic = (double)(1ull << (52 - fb)) + 16.0 / (double)(1u << fb);
  q.neg_scale = -1.0f / (float)(1u << fb);
  return q;
}

// amplified-linear weights (float64 like the reference, :233-241) -> one output sample; w = 0/0 -> NaN like numpy
__device__ __forceinline__ float combine_lin(const float w[4], const float dv[4], float v0) {
  const float den = (w[0] + w[1]) + (w[2] + w[3]);
  const float num = fmaf(w[1], dv[1], fmaf(w[2], dv[2], w[3] * dv[3]));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float qn = num * r;
  qn = fmaf(fmaf(-den, qn, num), r, qn);
  return v0 + qn;
}

// linear_weight (:233-241) of one tap: max(lin(dr), 0) * max(lin(dc), 0) with lin(x) = alpha x + 1 on [-1, 0), 1 - alpha x
// on [0, 1], 0 elsewhere -- i.e. (1 - alpha |x|) * [|x| <= 1].  `valid` = [|dr| <= 1] * [|dc| <= 1] as 0.0f / 1.0f.
// CLAMP is only needed for max_sigma > 1 (|alpha| <= 1 keeps 1 - alpha |x| >= 0 on [-1, 1]); there the two factors
// are clamped in float32.
template <bool CLAMP>
__device__ __forceinline__ float lin_weight(double alpha, double dr, double dc, float valid) {
  const double lr = fma(-alpha, fabs(dr), 1.0), lc = fma(-alpha, fabs(dc), 1.0);
  if (CLAMP) return fmaxf((float)lr, 0.0f) * fmaxf((float)lc, 0.0f) * valid;
  return (float)(lr * lc) * valid;
}

struct SmemG {
  CoefTabs tab;
  double sA[kWY][kWX], sB[kWY][kWX], sC[kWY][kWX];
  float sV[kWY][kWX];
};
struct SmemL {
  double al[256];
  double sA[kWY][kWX];
  float sV[kWY][kWX];
};

template <int KIND, int FMT, bool CLAMP>
__global__ void __launch_bounds__(kOX* kTY)
    resize_sr_tile_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH, int oW,
                          const int* __restrict__ left_y, const double* __restrict__ dist_y, const int* __restrict__ left_x,
                          const double* __restrict__ dist_x, const CoefTabs* __restrict__ ct, const FixQ fq, float max_sigma,
                          int channels, int oy0, int oy1, int rows_per_block, void* __restrict__ out) {
  __shared__ typename std::conditional<KIND == LERF_KIND_GAUSS, SmemG, SmemL>::type sm;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int p = blockIdx.z;
  const int ox_first = blockIdx.x * kOX, ox_last = min(ox_first + kOX - 1, oW - 1);
  const int oy_first = oy0 + blockIdx.y * rows_per_block, oy_last = min(oy_first + rows_per_block - 1, oy1 - 1);
  const int r0 = __ldg(left_y + oy_first), c0 = __ldg(left_x + ox_first);
  const int nr = __ldg(left_y + oy_last) - r0 + 2, nc = __ldg(left_x + ox_last) - c0 + 2;  // <= kWY, kWX (checked by the host)
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  if constexpr (KIND == LERF_KIND_GAUSS) {
    sm.tab.s2[tid] = __ldg(ct->s2 + tid);
    sm.tab.sg[tid] = __ldg(ct->sg + tid);
    sm.tab.rl[tid] = __ldg(ct->rl + tid);
  } else {  // alpha = fl(max_sigma * rho) in float32 like numpy (:249-250), promoted exactly
    const float h = __fdiv_rn((float)tid, 255.0f);
    sm.al[tid] = (double)__fmul_rn(max_sigma, __fsub_rn(__fmul_rn(h, 2.0f), 1.0f));
  }
  __syncthreads();
  const uint8_t* cp = codes + (long long)p * (KIND == LERF_KIND_GAUSS ? 3 : 1) * plane_sz;
  const float inv_nc = 1.0f / (float)nc;
  for (int i = tid; i < nr * nc; i += kOX * kTY) {
    // i / nc for i < 33 * 33: (i + 0.5) / nc is at least 0.5 / 33 away from an integer, far above the float error
    const int r = (int)(((float)i + 0.5f) * inv_nc), c = i - r * nc;
    const int sy = r0 + r, sx = c0 + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'edge' (:172-174)
    const long long off = (long long)cy * W + cx;
    if constexpr (KIND == LERF_KIND_GAUSS) {
      const int kr = __ldcg(cp + off), kx = __ldcg(cp + plane_sz + off), ky = __ldcg(cp + 2 * plane_sz + off);
      sm.sA[r][c] = sm.tab.s2[kx];
      sm.sC[r][c] = sm.tab.s2[ky];
      sm.sB[r][c] = sm.tab.rl[kr] * sm.tab.sg[kx] * sm.tab.sg[ky];
    } else {
      sm.sA[r][c] = sm.al[__ldcg(cp + off)];
    }
    sm.sV[r][c] = (sy == cy && sx == cx) ? (float)__ldcg(fp + off) : 0.0f;  // image: 'constant' 0 (:208)
  }
  __syncthreads();
  const int ox = ox_first + tx;
  if (ox > ox_last) return;
  const int lx = __ldg(left_x + ox) - c0;
  const double dc[2] = {__ldg(dist_x + 2 * ox), __ldg(dist_x + 2 * ox + 1)};
  const double ndc2[2] = {-dc[0] * dc[0], -dc[1] * dc[1]};
  const float vc[2] = {fabs(dc[0]) <= 1.0 ? 1.0f : 0.0f, fabs(dc[1]) <= 1.0 ? 1.0f : 0.0f};
  // Thread (tx, ty) owns rows_per_block / 8 CONSECUTIVE output rows: at scale s about s of them share their first tap row,
  // so the four taps' coefficients are re-read from shared memory only when that row changes (a warp-uniform branch:
  // the 32 lanes of a warp work on the same output row).
  // The linear kind keeps one row per pass and rows ty + 8 j (its taps are one value each; caching them costs
  // occupancy: measured 1.26 vs 1.22 ms on cfg-2).
  constexpr bool kCache = KIND == LERF_KIND_GAUSS;
  const int rpt = rows_per_block / kTY;
  const int row_step = kCache ? 1 : kTY;
  int oy = oy_first + (kCache ? ty * rpt : ty);
  const int oy_end = kCache ? min(oy + rpt - 1, oy_last) : oy_last;
  long long ip = ((long long)p * oH + oy) * oW + ox;                                                  // planar index
  long long ih = (((long long)(p / channels) * oH + oy) * oW + ox) * channels + (p % channels);      // interleaved index
  const long long ip_step = (long long)row_step * oW, ih_step = ip_step * channels;
  const int* lyp = left_y + oy;
  const double* dyp = dist_y + 2 * oy;
  int ly_cur = -1 << 30;
  double ca[4], cb[4], cq[4];  // Gaussian: a', b' * -dc, magic + c' * -dc^2 per tap;  linear: ca = alpha
  float dv[4], v0 = 0.0f;
#pragma unroll 1
  for (; oy <= oy_end; oy += row_step, ip += ip_step, ih += ih_step, lyp += row_step, dyp += 2 * row_step) {
    const int ly = __ldg(lyp) - r0;
    const double dr[2] = {__ldg(dyp), __ldg(dyp + 1)};
    if (!kCache || ly != ly_cur) {
      ly_cur = ly;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {  // patch order a*2+b as in the reference (:95-98)
          const int t = a * 2 + b;
          ca[t] = sm.sA[ly + b][lx + a];
          if constexpr (KIND == LERF_KIND_GAUSS) {  // column terms hoisted with the taps (r2: the ROWQ form of resample_int.cuh)
            cb[t] = sm.sB[ly + b][lx + a] * -dc[a];             // b' * -dc
            cq[t] = fma(sm.sC[ly + b][lx + a], ndc2[a], fq.magic);  // magic + c' * -dc^2  (>= magic)
          }
          dv[t] = sm.sV[ly + b][lx + a];
        }
      v0 = dv[0];
      dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;  // exact: integers in [-255, 255]
    }
    float res;
    if constexpr (KIND == LERF_KIND_GAUSS) {
      const double ndr2[2] = {-dr[0] * dr[0], -dr[1] * dr[1]};
      unsigned q[4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int t = a * 2 + b;
          // the two non-negative terms first, the signed cross term last: no partial sum leaves magic's binade
          const double e = fma(cb[t], dr[b], fma(ca[t], ndr2[b], cq[t]));
          q[t] = (unsigned)__double2loint(e);  // round(-log2 w * 2^FB) + 16
        }
      res = combine_uq(q, dv, v0, fq.neg_scale);
    } else {
      const float vr[2] = {fabs(dr[0]) <= 1.0 ? 1.0f : 0.0f, fabs(dr[1]) <= 1.0 ? 1.0f : 0.0f};
      float w[4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) w[a * 2 + b] = lin_weight<CLAMP>(ca[a * 2 + b], dr[b], dc[a], vr[b] * vc[a]);
      res = combine_lin(w, dv, v0);
    }
    store1<FMT>(out, ip, ih, res);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SR cell kernel (r2): the cell-owner layout of resample_int.cuh for ANY scale in [1, NMAX] per axis (non-integer,
// anisotropic, non-periodic).  For a scale >= 1 the outputs whose first tap is input sample l form a contiguous run
// [start[l], start[l + 1]) on each axis (plan->cell_y / cell_x, built from the caller's tables); one thread owns one cell
// = one (row run) x (column run), reads its 2 x 2 taps ONCE, keeps the column runs' distances in registers (at most
// NMAX columns) and walks its rows.  What this saves over the tile kernel above: the ~300-instruction per-thread set-up
// per 4..16 samples, the tap re-read test per row, and half the float64 work of an exponent -- the ROWQ form of
// resample_int.cuh, `rowq[t] = fma(a'_t, -dr^2, magic)` and `b'_t * -dr` once per output row, then
// `fma(b'dr_t, dc, fma(c'_t, -dc^2, rowq[t]))` per exponent.  The amplified-linear kind evaluates exactly the tile
// kernel's expression (bit-identical results).  A block = 32 x 8 cells, its taps a 33 x 9 window in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGX = 32, kGY = 8;
constexpr int kCellMaxY = 8;  // longest row run the kernel takes (its block stages kGY * kCellMaxY row distances)

struct SmemCG {
  CoefTabs tab;
  double sA[kGY + 1][kGX + 1], sB[kGY + 1][kGX + 1], sC[kGY + 1][kGX + 1];
  float sV[kGY + 1][kGX + 1];
};
struct SmemCL {
  double al[256];
  double sA[kGY + 1][kGX + 1];
  float sV[kGY + 1][kGX + 1];
};

template <int KIND, int FMT, int NMAX, bool CLAMP>
__global__ void __launch_bounds__(kGX* kGY, 3)
    resize_sr_cell_kernel(const uint8_t* __restrict__ feat, const uint8_t* __restrict__ codes, int H, int W, int oH, int oW,
                          const int* __restrict__ cell_y, const double* __restrict__ dist_y, const int* __restrict__ cell_x,
                          const double* __restrict__ dist_x, const CoefTabs* __restrict__ ct, const FixQ fq, float max_sigma,
                          int channels, int ly0, int oy0, int oy1, void* __restrict__ out) {
  __shared__ typename std::conditional<KIND == LERF_KIND_GAUSS, SmemCG, SmemCL>::type sm;
  __shared__ double2 s_dy[kGY * kCellMaxY];  // the block's output rows: distances to tap 0 / tap 1
  __shared__ double2 s_dx[kGX * NMAX];       // the block's output columns
  __shared__ int s_cx[kGX + 1];              // column runs of the block's cells
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int p = blockIdx.z;
  const int lxb = (int)blockIdx.x * kGX - 1, lyb = ly0 + (int)blockIdx.y * kGY;  // first cell of the block (cells start at -1)
  const long long plane_sz = (long long)H * W;
  const uint8_t* fp = feat + (long long)p * plane_sz;
  if constexpr (KIND == LERF_KIND_GAUSS) {
    sm.tab.s2[tid] = __ldg(ct->s2 + tid);
    sm.tab.sg[tid] = __ldg(ct->sg + tid);
    sm.tab.rl[tid] = __ldg(ct->rl + tid);
  } else {  // alpha = fl(max_sigma * rho) in float32 like numpy (:249-250), promoted exactly
    const float h = __fdiv_rn((float)tid, 255.0f);
    sm.al[tid] = (double)__fmul_rn(max_sigma, __fsub_rn(__fmul_rn(h, 2.0f), 1.0f));
  }
  const int yb0 = __ldg(cell_y + lyb + 1);                              // first output row of the block's cells
  const int yb1 = __ldg(cell_y + min(lyb + kGY, H) + 1);                // one past their last
  const int xb0 = __ldg(cell_x + lxb + 1);
  const int xb1 = __ldg(cell_x + min(lxb + kGX, W) + 1);
  if (tid < yb1 - yb0) s_dy[tid] = __ldg(reinterpret_cast<const double2*>(dist_y) + yb0 + tid);  // <= kGY * kCellMaxY (host check)
  if (tid < xb1 - xb0) s_dx[tid] = __ldg(reinterpret_cast<const double2*>(dist_x) + xb0 + tid);  // <= kGX * NMAX
  if (tid <= kGX) s_cx[tid] = __ldg(cell_x + min(lxb + tid, W) + 1);
  __syncthreads();
  const uint8_t* cp = codes + (long long)p * (KIND == LERF_KIND_GAUSS ? 3 : 1) * plane_sz;
  for (int i = tid; i < (kGY + 1) * (kGX + 1); i += kGX * kGY) {
    const int r = i / (kGX + 1), c = i - r * (kGX + 1);
    const int sy = lyb + r, sx = lxb + c;
    const int cy = min(max(sy, 0), H - 1), cx = min(max(sx, 0), W - 1);  // hypers: 'edge' (:172-174, :252-254)
    const long long off = (long long)cy * W + cx;
    if constexpr (KIND == LERF_KIND_GAUSS) {
      const int kr = __ldcg(cp + off), kx = __ldcg(cp + plane_sz + off), ky = __ldcg(cp + 2 * plane_sz + off);
      sm.sA[r][c] = sm.tab.s2[kx];
      sm.sC[r][c] = sm.tab.s2[ky];
      sm.sB[r][c] = sm.tab.rl[kr] * sm.tab.sg[kx] * sm.tab.sg[ky];
    } else {
      sm.sA[r][c] = sm.al[__ldcg(cp + off)];
    }
    sm.sV[r][c] = (sy == cy && sx == cx) ? (float)__ldcg(fp + off) : 0.0f;  // image: 'constant' 0 (:208, :268)
  }
  __syncthreads();
  const int lx = lxb + tx, ly = lyb + ty;
  if (lx > W - 1 || ly > H - 1) return;
  const int xs = s_cx[tx], nx = s_cx[tx + 1] - xs;  // the cell's column run (nx <= NMAX: host check)
  const int ys = max(__ldg(cell_y + ly + 1), oy0), ye = min(__ldg(cell_y + ly + 2), oy1);
  if (nx <= 0 || ys >= ye) return;
  double ca[4], cb[4], cc[4];  // Gaussian: a', b', -c' per tap;  linear: ca = alpha
  float dv[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {  // patch order a*2+b as in the reference (:95-98)
      const int t = a * 2 + b;
      ca[t] = sm.sA[ty + b][tx + a];
      if constexpr (KIND == LERF_KIND_GAUSS) {
        cb[t] = sm.sB[ty + b][tx + a];
        cc[t] = -sm.sC[ty + b][tx + a];
      }
      dv[t] = sm.sV[ty + b][tx + a];
    }
  const float v0 = dv[0];
  dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;  // exact: integers in [-255, 255]
  double dc[NMAX][2];     // Gaussian: dc;  linear: |dc|
  float vc[NMAX][2];      // linear only: [|dc| <= 1]
#pragma unroll
  for (int mc = 0; mc < NMAX; ++mc) {
    const double2 d = s_dx[min(xs - xb0 + mc, kGX * NMAX - 1)];  // columns past the run: any finite value (never stored)
    if constexpr (KIND == LERF_KIND_GAUSS) {
      dc[mc][0] = d.x; dc[mc][1] = d.y;
    } else {
      dc[mc][0] = fabs(d.x); dc[mc][1] = fabs(d.y);
      vc[mc][0] = dc[mc][0] <= 1.0 ? 1.0f : 0.0f;
      vc[mc][1] = dc[mc][1] <= 1.0 ? 1.0f : 0.0f;
    }
  }
  const long long ip = ((long long)p * oH + ys) * oW + xs;                                             // planar index
  const long long ih = (((long long)(p / channels) * oH + ys) * oW + xs) * channels + (p % channels);  // interleaved index
  // one pointer per thread, advanced row by row; column mc is a constant offset from it
  constexpr int kEl = FMT == LERF_OUT_F32 ? 4 : 1;
  unsigned char* op = (unsigned char*)out + (FMT == LERF_OUT_U8_HWC ? ih : ip) * kEl;
  const long long row_step = (long long)oW * kEl * (FMT == LERF_OUT_U8_HWC ? channels : 1);
  const int col_step = FMT == LERF_OUT_U8_HWC ? channels : 1;
  const double2* dyp = s_dy + (ys - yb0);
#pragma unroll 1
  for (int oy = ys; oy < ye; ++oy, op += row_step, ++dyp) {
    const double2 dr = *dyp;  // one address per warp: a shared-memory broadcast
    float res[NMAX];
    if constexpr (KIND == LERF_KIND_GAUSS) {
      const double drb[2] = {dr.x, dr.y};
      double rowq[4], bdr[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        rowq[t] = fma(ca[t], -drb[t & 1] * drb[t & 1], fq.magic);  // >= magic: the quadratic form's row term is >= 0
        bdr[t] = cb[t] * -drb[t & 1];
      }
#pragma unroll
      for (int mc = 0; mc < NMAX; ++mc) {
        unsigned q[4];
#pragma unroll
        for (int t = 0; t < 4; ++t)  // dc (b' -dr + -c' dc) + rowq: Horner in dc, ONE rounding at the fixed-point scale here
          q[t] = (unsigned)__double2loint(fma(dc[mc][t >> 1], fma(cc[t], dc[mc][t >> 1], bdr[t]), rowq[t]));
        res[mc] = combine_uq(q, dv, v0, fq.neg_scale);
      }
    } else {
      const double adr[2] = {fabs(dr.x), fabs(dr.y)};
      const float vr[2] = {adr[0] <= 1.0 ? 1.0f : 0.0f, adr[1] <= 1.0 ? 1.0f : 0.0f};
      double lr[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) lr[t] = fma(-ca[t], adr[t & 1], 1.0);
#pragma unroll
      for (int mc = 0; mc < NMAX; ++mc) {
        float w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // lin_weight<CLAMP> with the row factor hoisted
          const double lc = fma(-ca[t], dc[mc][t >> 1], 1.0);
          const float valid = vr[t & 1] * vc[mc][t >> 1];
          w[t] = CLAMP ? fmaxf((float)lr[t], 0.0f) * fmaxf((float)lc, 0.0f) * valid : (float)(lr[t] * lc) * valid;
        }
        res[mc] = combine_lin(w, dv, v0);
      }
    }
#pragma unroll
    for (int mc = 0; mc < NMAX; ++mc)
      if (mc < nx) {
        if (FMT == LERF_OUT_F32) {
          __stcg(reinterpret_cast<float*>(op) + mc, res[mc]);
        } else {
          const int qv = min(max(__float2int_rn(res[mc]), 0), 255);  // round half to even, like store1
          __stcg(op + mc * col_step, (unsigned char)qv);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// warp: one thread = one output pixel, all planes; float64 geometry in the reference's operation order
// ---------------------------------------------------------------------------------------------------------------
// One 32-byte record per input sample for the Gaussian warp: the tap's three exponent coefficients and its value, so
// the warp kernel fetches a tap with ONE 256-bit gather instead of four byte gathers, three table reads and two DMULs.
struct __align__(32) TapRec {
  double a, b, c;  // -L/2 sx^2, L rho sx sy, -L/2 sy^2   (L = log2 e)
  float v, pad;
};

__global__ void __launch_bounds__(256)
    warp_records_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ codes, long long plane_sz, float max_sigma,
                        TapRec* __restrict__ rec) {
  __shared__ double t_s2[256], t_sg[256], t_rl[256];
  {
    const int c = threadIdx.x;  // float32 decode exactly like numpy (eval_lut_warp.py:186-191, resize_right2d_numpy.py:522-524)
    const float h = __fdiv_rn((float)c, 255.0f);
    const float rho = __fsub_rn(__fmul_rn(h, 2.0f), 1.0f);
    const float sig = __fmul_rn(h, max_sigma);
    t_s2[c] = __dmul_rn(-0.5 * kLog2e, __dmul_rn((double)sig, (double)sig));
    t_sg[c] = (double)sig;
    t_rl[c] = __dmul_rn(kLog2e, (double)rho);
  }
  __syncthreads();
  const int p = blockIdx.y;
  const uint8_t* cp = codes + (long long)p * 3 * plane_sz;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < plane_sz; i += (long long)gridDim.x * 256) {
    const int kr = __ldcg(cp + i), kx = __ldcg(cp + plane_sz + i), ky = __ldcg(cp + 2 * plane_sz + i);
    TapRec r;
    r.a = t_s2[kx];
    r.c = t_s2[ky];
    r.b = t_rl[kr] * t_sg[kx] * t_sg[ky];
    r.v = (float)__ldcg(img + (long long)p * plane_sz + i);
    r.pad = 0.0f;
    rec[(long long)p * plane_sz + i] = r;
  }
}

struct WarpGeomF {
  double m[9];
  int H, W, oH, oW;
  int pad0_y, pad0_x;
  int mpad0_y, mpad0_x, border;
};

__constant__ double kEps32f = 1.1920928955078125e-07;

__device__ __forceinline__ int clampi3(int v, int lo, int hi) { return min(max(v, lo), hi); }

// REC: Gaussian taps come from the TapRec array (`rec`) instead of img/codes + tables.
template <int KIND, int FMT, bool CLAMP, bool REC>
__global__ void __launch_bounds__(256)
    warp_fast_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ codes, const TapRec* __restrict__ rec,
                     const WarpGeomF g, int planes, int channels, float max_sigma, const FixQ fq, void* __restrict__ out,
                     uint8_t* __restrict__ mask) {
  __shared__ double t_s2[REC ? 1 : 256], t_sg[REC ? 1 : 256], t_rl[REC ? 1 : 256];  // Gauss: -L/2 sigma^2, sigma, L rho;  linear: t_s2 = alpha
  if (!REC) {
    const int c = threadIdx.y * 32 + threadIdx.x;  // 256 threads = 256 codes; float32 decode exactly like numpy
    const float h = __fdiv_rn((float)c, 255.0f);
    const float rho = __fsub_rn(__fmul_rn(h, 2.0f), 1.0f);
    if (KIND == LERF_KIND_GAUSS) {
      const float sig = __fmul_rn(h, max_sigma);
      t_s2[c] = __dmul_rn(-0.5 * kLog2e, __dmul_rn((double)sig, (double)sig));
      t_sg[c] = (double)sig;
      t_rl[c] = __dmul_rn(kLog2e, (double)rho);
    } else {
      t_s2[c] = (double)__fmul_rn(max_sigma, rho);
    }
  }
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

  if (mask) {  // NearestWarp2dNumpy (:460-467): support 1, box2d weight, white frame test (eval_lut_warp.py:197-204)
    const int fr = clampi3((int)ceil(pr0 - 0.5 - kEps32f) + g.mpad0_y, 0, H - 1);
    const int fc = clampi3((int)ceil(pc0 - 0.5 - kEps32f) + g.mpad0_x, 0, W - 1);
    const double dr = (pr0 + (double)g.mpad0_y) - (double)fr, dc = (pc0 + (double)g.mpad0_x) - (double)fc;
    const int sr = fr - g.mpad0_y, sc = fc - g.mpad0_x;
    const bool hit = (-1.0 <= dr && dr <= 1.0) && (-1.0 <= dc && dc <= 1.0) && sr >= g.border &&
                     sr < H - g.border && sc >= g.border && sc < W - g.border;
    mask[(long long)oy * g.oW + ox] = hit ? 1 : 0;
  }
  if (!out) return;

  const int lr = (int)ceil(pr0 - 1.0 - kEps32f) + g.pad0_y;  // :347-352, :366
  const int lc = (int)ceil(pc0 - 1.0 - kEps32f) + g.pad0_x;
  const double pr = pr0 + (double)g.pad0_y, pc = pc0 + (double)g.pad0_x;  // :367
  int off[4];
  bool inside[4];
  double dr[2], dc[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    dr[k] = pr - (double)clampi3(lr + k, 0, H - 1);  // taps clipped in padded coordinates (:397-403)
    dc[k] = pc - (double)clampi3(lc + k, 0, W - 1);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int sr = clampi3(lr + b, 0, H - 1) - g.pad0_y, sc = clampi3(lc + a, 0, W - 1) - g.pad0_x;
      inside[a * 2 + b] = sr >= 0 && sc >= 0;
      off[a * 2 + b] = max(sr, 0) * W + max(sc, 0);
    }
  // Gauss: geometry factors of the exponent, shared by all planes.  Where a tap was clipped the distance can exceed 1
  // (only outside the validity mask); it is clamped so the fixed-point exponent cannot wrap.
  // Where a tap was clipped (image borders, outside the validity mask) a distance can exceed 1 and the fixed-point
  // exponent could wrap: those pixels take the float64 max-subtracted form instead (far).
  double ndr2[2], ndc2[2], npp[2][2];
  bool far = false;
  if (KIND == LERF_KIND_GAUSS) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      far = far || fabs(dr[k]) > fq.dlim || fabs(dc[k]) > fq.dlim;
      ndr2[k] = -dr[k] * dr[k];
      ndc2[k] = -dc[k] * dc[k];
    }
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) npp[b][a] = -(dr[b] * dc[a]);
  }
  float valid[4];  // linear: [|dr| <= 1] * [|dc| <= 1] per tap
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) valid[a * 2 + b] = (fabs(dr[b]) <= 1.0 && fabs(dc[a]) <= 1.0) ? 1.0f : 0.0f;
  const long long plane_sz = (long long)H * W;
  const long long osz = (long long)g.oH * g.oW, opix = (long long)oy * g.oW + ox;
  for (int p = 0; p < planes; ++p) {
    const uint8_t* fp = img + (long long)p * plane_sz;
    const uint8_t* cp = codes + (long long)p * (KIND == LERF_KIND_GAUSS ? 3 : 1) * plane_sz;
    float dv[4];
    float res;
    if (KIND == LERF_KIND_GAUSS) {
      unsigned q[4];
      double ef[4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int t = a * 2 + b;
          double ca, cb, cc;
          if (REC) {
            uint32_t r0, r1, r2, r3, r4, r5, r6, r7;  // one 256-bit load of the tap's record
            asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                : "l"(rec + (long long)p * plane_sz + off[t]));
            ca = __hiloint2double((int)r1, (int)r0);
            cb = __hiloint2double((int)r3, (int)r2);
            cc = __hiloint2double((int)r5, (int)r4);
            dv[t] = inside[t] ? __uint_as_float(r6) : 0.0f;
          } else {
            const int kr = __ldg(cp + off[t]), kx = __ldg(cp + plane_sz + off[t]), ky = __ldg(cp + 2 * plane_sz + off[t]);
            ca = t_s2[kx];
            cc = t_s2[ky];
            cb = t_rl[kr] * t_sg[kx] * t_sg[ky];
            dv[t] = inside[t] ? (float)__ldg(fp + off[t]) : 0.0f;
          }
          double e = ca * ndr2[b];
          e = fma(cc, ndc2[a], e);
          e = fma(cb, npp[b][a], e);
          ef[t] = e;  // -log2 w >= 0
          q[t] = (unsigned)__double2loint(e + fq.magic);
        }
      const float v0 = dv[0];
      dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;
      if (!far) {
        res = combine_uq(q, dv, v0, fq.neg_scale);
      } else {
        const double m = fmin(fmin(ef[0], ef[1]), fmin(ef[2], ef[3]));
        float w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float xx = (float)(m - ef[t]);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[t]) : "f"(xx));
        }
        res = combine_lin(w, dv, v0);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) dv[t] = inside[t] ? (float)__ldg(fp + off[t]) : 0.0f;
      const float v0 = dv[0];
      dv[1] -= v0; dv[2] -= v0; dv[3] -= v0;
      float w[4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int t = a * 2 + b;
          w[t] = lin_weight<CLAMP>(t_s2[__ldg(cp + off[t])], dr[b], dc[a], valid[t]);
        }
      res = combine_lin(w, dv, v0);
    }
    const long long ip = (long long)p * osz + opix;
    const long long ih = FMT == LERF_OUT_U8_HWC ? ((long long)(p / channels) * osz + opix) * channels + (p % channels) : 0;
    store1<FMT>(out, ip, ih, res);
  }
}

}  // namespace

// Called by lerf_resize_sr for uint8 code inputs when the plan's tap windows fit the tile (any scale >= 1).
// Returns -1 when this path does not apply.
int resize_sr_tile(int kind, const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                   float max_sigma, int oy0, int oy1, void* out, int fmt, cudaStream_t st) {
  if (!P->tile_ok || !(max_sigma >= 0.0f) || max_sigma > 64.0f) return -1;
  const FixQ fq = make_fixq(max_sigma);
  if (kind == LERF_KIND_GAUSS && fq.fb < kMinFracBits) return -1;  // too wide an exponent range for the fixed point: the float64 kernels take it
  const CoefTabs* ct = nullptr;
  if (kind == LERF_KIND_GAUSS) {
    ct = plan_coef_tabs(P, max_sigma, st);
    if (!ct) return fail(LERF_ECUDA, "uploading the hyper decode tables failed");
  }
  const int rpb = P->tile_rows;
  dim3 block(kOX * kTY), grid((P->oW + kOX - 1) / kOX, (oy1 - oy0 + rpb - 1) / rpb, planes);
#define LERF_GK(K, F, CL)                                                                                                 \
  resize_sr_tile_kernel<K, F, CL><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, P->left_y, P->dist_y,   \
                                                          P->left_x, P->dist_x, ct, fq, max_sigma, channels, oy0, oy1, rpb, out)
#define LERF_GO(F)                                                  \
  if (kind == LERF_KIND_GAUSS) LERF_GK(LERF_KIND_GAUSS, F, false);   \
  else if (max_sigma > 1.0f) LERF_GK(LERF_KIND_LINEAR, F, true);    \
  else LERF_GK(LERF_KIND_LINEAR, F, false)
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


// Called by lerf_resize_sr before the tile kernel.  It applies to any scale whose column runs hold at most 4 outputs and
// whose row runs at most 8, and it PAYS from x3 per axis up (measured, 8 frames 2040x1356, G samples/s cell / tile: x3.5
// Gaussian 325 / 265, linear 381 / 305; x2.5 207 / 218 and 248 / 282; x1.5 95 / 141 -- a cell of 2 x 2 outputs does not
// amortise its set-up), so smaller scales stay on the tile kernel unless `any_scale` (lerf_debug_force_generic(3), the
// parity tests).  Returns -1 when this path does not apply.
int resize_sr_cell(int kind, const lerf_sr_plan_impl* P, const uint8_t* feat, const uint8_t* codes, int planes, int channels,
                   float max_sigma, int oy0, int oy1, void* out, int fmt, bool any_scale, cudaStream_t st) {
  if (!any_scale && (P->oW < 3 * (long long)P->W || P->oH < 3 * (long long)P->H)) return -1;
  if (!P->tile_ok || !P->cell_x || P->cell_max_x > 4 || P->cell_max_y > kCellMaxY || !(max_sigma >= 0.0f) || max_sigma > 64.0f) return -1;
  const FixQ fq = make_fixq(max_sigma);
  if (kind == LERF_KIND_GAUSS && fq.fb < kMinFracBits) return -1;  // too wide an exponent range for the fixed point: the float64 kernels take it
  const CoefTabs* ct = nullptr;
  if (kind == LERF_KIND_GAUSS) {
    ct = plan_coef_tabs(P, max_sigma, st);
    if (!ct) return fail(LERF_ECUDA, "uploading the hyper decode tables failed");
  }
  const int ly0 = P->h_left_y[oy0], ly1 = P->h_left_y[oy1 - 1];
  dim3 block(kGX * kGY), grid((P->W + 1 + kGX - 1) / kGX, (ly1 - ly0 + 1 + kGY - 1) / kGY, planes);
#define LERF_GK(K, F, N, CL)                                                                                              \
  resize_sr_cell_kernel<K, F, N, CL><<<grid, block, 0, st>>>(feat, codes, P->H, P->W, P->oH, P->oW, P->cell_y, P->dist_y, \
                                                             P->cell_x, P->dist_x, ct, fq, max_sigma, channels, ly0, oy0, oy1, out)
#define LERF_GN(K, F, CL)                         \
  if (P->cell_max_x <= 2) LERF_GK(K, F, 2, CL);      \
  else if (P->cell_max_x == 3) LERF_GK(K, F, 3, CL); \
  else LERF_GK(K, F, 4, CL)
#define LERF_GO(F)                                                  \
  if (kind == LERF_KIND_GAUSS) { LERF_GN(LERF_KIND_GAUSS, F, false); }   \
  else if (max_sigma > 1.0f) { LERF_GN(LERF_KIND_LINEAR, F, true); }    \
  else { LERF_GN(LERF_KIND_LINEAR, F, false); }
  switch (fmt) {
    case LERF_OUT_F32: LERF_GO(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GO(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GO(LERF_OUT_U8_HWC); break;
    default: return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GO
#undef LERF_GN
#undef LERF_GK
  LERF_LAUNCHED();
  return LERF_OK;
}

// Stream-ordered scratch for the records: a library-owned memory pool per device that keeps its memory between calls
// (the default pool hands it back to the driver at every synchronisation: 6 ms per 100 MB call, measured).
static cudaMemPool_t scratch_pool() {
  static cudaMemPool_t pools[64] = {};
  static std::mutex mu;  // lerf_warp may be called from several host threads (one per stream)
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = pool;
  }
  return pools[dev];
}

// Called by lerf_warp for uint8 code inputs.  Returns -1 when this path does not apply.
int warp_fast(int kind, const uint8_t* feat, const uint8_t* codes, int planes, int channels, int H, int W, int oH, int oW,
              const double minv[9], int pad0_y, int pad0_x, int mpad0_y, int mpad0_x, int border, float max_sigma, void* out,
              int fmt, uint8_t* mask, cudaStream_t st) {
  if (!(max_sigma >= 0.0f) || max_sigma > 64.0f || (long long)H * W >= 2147483647LL) return -1;
  WarpGeomF g;
  for (int i = 0; i < 9; ++i) g.m[i] = minv[i];
  g.H = H; g.W = W; g.oH = oH; g.oW = oW;
  g.pad0_y = pad0_y; g.pad0_x = pad0_x; g.mpad0_y = mpad0_y; g.mpad0_x = mpad0_x; g.border = border;
  const FixQ fq = make_fixq(max_sigma);
  if (kind == LERF_KIND_GAUSS && fq.fb < kMinFracBits) return -1;  // too wide an exponent range for the fixed point: the float64 kernels take it
  dim3 block(32, 8), grid((oW + 31) / 32, (oH + 7) / 8, 1);
  // Gaussian: decode every input sample once into a 32-byte record (stream-ordered scratch), then gather records.
  TapRec* rec = nullptr;
  if (kind == LERF_KIND_GAUSS && out && planes > 0 && g_dbg.warp_records) {
    const long long n = (long long)planes * H * W;
    cudaMemPool_t pool = scratch_pool();
    if (pool && cudaMallocFromPoolAsync((void**)&rec, (size_t)n * sizeof(TapRec), pool, st) == cudaSuccess) {
      dim3 rgrid((unsigned)min((long long)((long long)H * W + 255) / 256, 4096LL), planes);
      warp_records_kernel<<<rgrid, 256, 0, st>>>(feat, codes, (long long)H * W, max_sigma, rec);
      ++g_launches;
      const cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) {  // give the scratch back before reporting
        cudaFreeAsync(rec, st);
        return fail(LERF_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
      }
    } else {
      (void)cudaGetLastError();  // no scratch: the table form below needs none
      rec = nullptr;
    }
  }
#define LERF_GK(K, F, CL, RC) \
  warp_fast_kernel<K, F, CL, RC><<<grid, block, 0, st>>>(feat, codes, rec, g, planes, channels, max_sigma, fq, out, mask)
#define LERF_GO(F)                                                         \
  if (kind == LERF_KIND_GAUSS && rec) LERF_GK(LERF_KIND_GAUSS, F, false, true); \
  else if (kind == LERF_KIND_GAUSS) LERF_GK(LERF_KIND_GAUSS, F, false, false);  \
  else if (max_sigma > 1.0f) LERF_GK(LERF_KIND_LINEAR, F, true, false);     \
  else LERF_GK(LERF_KIND_LINEAR, F, false, false)
  switch (fmt) {
    case LERF_OUT_F32: LERF_GO(LERF_OUT_F32); break;
    case LERF_OUT_U8: LERF_GO(LERF_OUT_U8); break;
    case LERF_OUT_U8_HWC: LERF_GO(LERF_OUT_U8_HWC); break;
    default:
      if (rec) cudaFreeAsync(rec, st);
      return fail(LERF_EINVAL, "unknown out_format %d", fmt);
  }
#undef LERF_GO
#undef LERF_GK
  if (rec) cudaFreeAsync(rec, st);  // stream-ordered: released after the warp kernel
  LERF_LAUNCHED();
  return LERF_OK;
}


}  // namespace lerf

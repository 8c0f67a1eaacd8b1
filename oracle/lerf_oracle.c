/*
 * lerf_oracle.c -- CPU ORACLE for the LeRF LUT inference hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (lerf_pytorch_b200/) never links, imports
 * or falls back to anything in oracle/.
 *
 * It is a plain-C restatement of the reference's numpy algorithm, function by
 * function, keeping the reference's operation order so float64 results agree
 * to rounding noise and the integer stages agree bit for bit.  Citations are
 * file:line into the upstream reference (ddlee-cn/LeRF-PyTorch).
 *
 * PARITY PINNING: the reference ships no tests.  This oracle is pinned against
 * golden vectors produced by importing and running the reference's own Python
 * code (tests/golden/make_golden.py, fixtures under tests/golden/), including
 * the Set5 fixtures whose published PSNR the reference quotes in scripts.sh:33-47.
 * tests/test_oracle_golden.py is the pin.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LERF_Q 16 /* q = 2**interval, interval = 4   (eval_lut_sr.py:27) */
#define LERF_L 17 /* L = 2**(8-interval) + 1         (eval_lut_sr.py:28) */

void lerf_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int lerf_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Tap offsets (row, col) of samples a,b,c,d inside the rotated, edge-padded
 * image and the bottom/right pad each mode needs.
 * eval_lut_sr.py:12-18 (mode_pad_dict) and :30-81 (the slices). */
static int mode_taps(char mode, int di[4], int dj[4]) {
  static const int S[2][4] = {{0, 0, 1, 1}, {0, 1, 0, 1}};
  static const int D[2][4] = {{0, 0, 2, 2}, {0, 2, 0, 2}};
  static const int Y[2][4] = {{0, 1, 1, 2}, {0, 1, 2, 1}};
  static const int Cc[2][4] = {{0, 0, 0, 0}, {0, 1, 2, 3}};
  static const int T[2][4] = {{0, 1, 2, 3}, {0, 1, 2, 3}};
  const int(*m)[4];
  int pad;
  switch (mode) {
    case 's': m = S; pad = 1; break;
    case 'd': m = D; pad = 2; break;
    case 'y': m = Y; pad = 2; break;
    case 'c': m = Cc; pad = 3; break;
    case 't': m = T; pad = 3; break;
    default: return -1; /* eval_lut_sr.py:82-84 raises ValueError */
  }
  for (int k = 0; k < 4; ++k) { di[k] = m[0][k]; dj[k] = m[1][k]; }
  return pad;
}

int lerf_oracle_mode_pad(char mode) {
  int di[4], dj[4];
  return mode_taps(mode, di, dj);
}

/* The 24 mutually exclusive simplex cases of eval_lut_sr.py:218-462, kept as
 * the reference's own decision structure (strict '>' comparisons, the
 * "overflow bug fix" ordering of cases 10/11 at :310-336).  Writes the tap
 * order (0=a,1=b,2=c,3=d), largest LSB first. */
static void simplex_order(int fa, int fb, int fc, int fd, int ord[4]) {
  const int fab = fa > fb, fac = fa > fc, fad = fa > fd;
  const int fbc = fb > fc, fbd = fb > fd, fcd = fc > fd;
#define ORD(w, x, y, z) do { ord[0] = w; ord[1] = x; ord[2] = y; ord[3] = z; } while (0)
  if (fab && fbc) {                 /* i1..i4   :226-262 */
    if (fcd)      ORD(0, 1, 2, 3);
    else if (fbd) ORD(0, 1, 3, 2);
    else if (fad) ORD(0, 3, 1, 2);
    else          ORD(3, 0, 1, 2);
  } else if (fab && fac) {          /* i5..i8   :264-300  (~fbc) */
    if (fbd)      ORD(0, 2, 1, 3);
    else if (fcd) ORD(0, 2, 3, 1);
    else if (fad) ORD(0, 3, 2, 1);
    else          ORD(3, 0, 2, 1);
  } else if (fab) {                 /* i9..i12  :302-346  (~fbc, ~fac) */
    if (fbd)      ORD(2, 0, 1, 3);
    else if (fad) ORD(2, 0, 3, 1);  /* c > a > d > b, :315-324 */
    else if (fcd) ORD(2, 3, 0, 1);  /* c > d > a > b, :325-336 */
    else          ORD(3, 2, 0, 1);
  } else if (fac) {                 /* i13..i16 :348-384  (~fab) */
    if (fcd)      ORD(1, 0, 2, 3);
    else if (fad) ORD(1, 0, 3, 2);
    else if (fbd) ORD(1, 3, 0, 2);
    else          ORD(3, 1, 0, 2);
  } else if (fbc) {                 /* i17..i20 :386-423  (~fab, ~fac) */
    if (fad)      ORD(1, 2, 0, 3);
    else if (fcd) ORD(1, 2, 3, 0);
    else if (fbd) ORD(1, 3, 2, 0);
    else          ORD(3, 1, 2, 0);
  } else {                          /* i21..i24 :425-462  (~fab, ~fac, ~fbc) */
    if (fad)      ORD(2, 1, 0, 3);
    else if (fbd) ORD(2, 1, 3, 0);
    else if (fcd) ORD(2, 3, 1, 0);
    else          ORD(3, 2, 1, 0);
  }
#undef ORD
}

/* One LUT pass over one rotated, edge-padded plane set.
 * Restates FourSimplexInterpFaster (eval_lut_sr.py:24-470) up to, but not
 * including, the final np.rot90 (:468) -- see lerf_oracle_rot90_planes.
 *   weight : int8 [17^4][oC]        (the reference holds the same values as fp32)
 *   img    : uint8 [C][h+pad][w+pad] (the reference holds the same values as fp32)
 *   out    : double [C*oC][h][w] = N/16, channel c*oC+k  (:464-469)
 * The reference's arithmetic is exact in fp32 (integers < 2^24), so int32 here
 * gives the identical value; the divide by q is done in double as at :469. */
int lerf_oracle_lut_pass(const int8_t* weight, const uint8_t* img, int C, int h,
                         int w, char mode, int oC, double* out) {
  int di[4], dj[4];
  const int pad = mode_taps(mode, di, dj);
  if (pad < 0) return 1;
  const int hp = h + pad, wp = w + pad;
  static const int stride[4] = {LERF_L * LERF_L * LERF_L, LERF_L * LERF_L, LERF_L, 1};
#pragma omp parallel for collapse(2) schedule(static)
  for (int c = 0; c < C; ++c) {
    for (int i = 0; i < h; ++i) {
      const uint8_t* plane = img + (size_t)c * hp * wp;
      for (int j = 0; j < w; ++j) {
        int msb[4], f[4];
        for (int k = 0; k < 4; ++k) {
          const int v = plane[(size_t)(i + di[k]) * wp + (j + dj[k])];
          msb[k] = v / LERF_Q; /* :32-35 */
          f[k] = v % LERF_Q;   /* :38-41 */
        }
        int ord[4];
        simplex_order(f[0], f[1], f[2], f[3], ord);
        /* vertices p0000 -> p1111 along the sorted taps (:91-187), weights
         * (q-f1), (f1-f2), (f2-f3), (f3-f4), f4 (:227-233 and siblings) */
        int base = msb[0] * stride[0] + msb[1] * stride[1] + msb[2] * stride[2] + msb[3] * stride[3];
        int idx[5], wt[5];
        idx[0] = base;
        wt[0] = LERF_Q - f[ord[0]];
        for (int k = 0; k < 4; ++k) {
          idx[k + 1] = idx[k] + stride[ord[k]];
          wt[k + 1] = f[ord[k]] - (k < 3 ? f[ord[k + 1]] : 0);
        }
        for (int o = 0; o < oC; ++o) {
          int32_t acc = 0;
          for (int k = 0; k < 5; ++k) acc += wt[k] * (int32_t)weight[(size_t)idx[k] * oC + o];
          out[((size_t)(c * oC + o) * h + i) * w + j] = (double)acc / (double)LERF_Q; /* :469 */
        }
      }
    }
  }
  return 0;
}

/* np.rot90(m, k, axes=(1,2)) on [P][h][w] planes (numpy semantics: k counter-
 * clockwise quarter turns).  dst is [P][h'][w'] with (h',w') = (w,h) for odd k.
 * Used for img rotation (eval_lut_sr.py:549) and un-rotation (:468). */
#define DEFINE_ROT90(NAME, T)                                                         \
  void NAME(const T* src, int P, int h, int w, int k, T* dst) {                       \
    k = ((k % 4) + 4) % 4;                                                            \
    const int oh = (k & 1) ? w : h, ow = (k & 1) ? h : w;                             \
    for (int p = 0; p < P; ++p) {                                                     \
      const T* s = src + (size_t)p * h * w;                                           \
      T* d = dst + (size_t)p * oh * ow;                                               \
      for (int i = 0; i < oh; ++i)                                                    \
        for (int j = 0; j < ow; ++j) {                                                \
          int si, sj;                                                                 \
          switch (k) {                                                                \
            case 0: si = i; sj = j; break;                                            \
            case 1: si = j; sj = w - 1 - i; break;                                    \
            case 2: si = h - 1 - i; sj = w - 1 - j; break;                            \
            default: si = h - 1 - j; sj = i; break;                                   \
          }                                                                           \
          d[(size_t)i * ow + j] = s[(size_t)si * w + sj];                             \
        }                                                                             \
    }                                                                                 \
  }
DEFINE_ROT90(lerf_oracle_rot90_u8, uint8_t)
DEFINE_ROT90(lerf_oracle_rot90_f64, double)

/* np.pad(img, ((0,pad),(0,pad)), mode="edge") per plane (eval_lut_sr.py:551-553). */
static void edge_pad_br(const uint8_t* src, int P, int h, int w, int pad, uint8_t* dst) {
  const int hp = h + pad, wp = w + pad;
  for (int p = 0; p < P; ++p)
    for (int i = 0; i < hp; ++i)
      for (int j = 0; j < wp; ++j) {
        const int si = i < h ? i : h - 1, sj = j < w ? j : w - 1;
        dst[((size_t)p * hp + i) * wp + j] = src[((size_t)p * h + si) * w + sj];
      }
}

/* Rotation-ensembled LUT stage: the loops of eval_lut_sr.py:541-577 (stage 1)
 * and :579-628 (stage 2); identical copies at eval_lut_warp.py:104-191.
 *   tables : stage 1: n_modes tables (key s1_<mode>r0), used for r = 0..3
 *            stage 2: 2*n_modes tables ordered [mode][r0, r1]; r in {0,2} uses
 *            r0, r in {1,3} uses r1 (:582-619)
 *   img    : uint8 planar [C][H][W]
 *   pred   : double [C*oC][H][W] = sum of the passes (the reference's `pred`)
 *   out    : uint8 [C*oC][H][W] = round(clip(pred/avg + bias, 0, 255))
 *            stage 1: avg = n_modes, bias 0 (:566); stage 2: avg = 4*n_modes,
 *            bias = 255//2 = 127 (:621-628; the /255 to float is the caller's). */
int lerf_oracle_stage(int stage, const int8_t* const* tables, const char* modes,
                      int n_modes, int oC, const uint8_t* img, int C, int H, int W,
                      double* pred, uint8_t* out) {
  const size_t np_ = (size_t)C * oC * H * W;
  memset(pred, 0, np_ * sizeof(double));
  uint8_t* rot = (uint8_t*)malloc((size_t)C * H * W);
  uint8_t* padded = (uint8_t*)malloc((size_t)C * (H + 3) * (W + 3));
  double* pass = (double*)malloc(np_ * sizeof(double));
  double* back = (double*)malloc(np_ * sizeof(double));
  if (!rot || !padded || !pass || !back) { free(rot); free(padded); free(pass); free(back); return 2; }
  int rc = 0;
  for (int m = 0; m < n_modes && !rc; ++m) {
    const int pad = lerf_oracle_mode_pad(modes[m]);
    if (pad < 0) { rc = 1; break; }
    for (int r = 0; r < 4; ++r) {
      const int8_t* weight = (stage == 1) ? tables[m] : tables[2 * m + (r & 1)];
      const int h = (r & 1) ? W : H, w = (r & 1) ? H : W;
      lerf_oracle_rot90_u8(img, C, H, W, r, rot);                       /* :549 */
      edge_pad_br(rot, C, h, w, pad, padded);                           /* :551-553 */
      rc = lerf_oracle_lut_pass(weight, padded, C, h, w, modes[m], oC, pass); /* :554-564 */
      if (rc) break;
      lerf_oracle_rot90_f64(pass, C * oC, h, w, 4 - r, back);           /* :468, rot = 4 - r */
      for (size_t i = 0; i < np_; ++i) pred[i] += back[i];              /* pred += ... */
    }
  }
  if (!rc) {
    const double avg = (stage == 1) ? (double)n_modes : (double)(n_modes * 4);
    const double bias = (stage == 1) ? 0.0 : 127.0;
    for (size_t i = 0; i < np_; ++i) {
      double v = pred[i] / avg + bias;        /* :574 / :624 */
      v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);   /* np.clip */
      out[i] = (uint8_t)rint(v);              /* np.round = half to even */
    }
  }
  free(rot); free(padded); free(pass); free(back);
  return rc;
}

/* ---------------------------------------------------------------------------
 * Resampling (float64), resize_right/resize_right2d_numpy.py
 * ------------------------------------------------------------------------- */

static const double EPS32 = 1.1920928955078125e-07; /* np.finfo(np.float32).eps, :12 */

/* Gaussian weight: SteeringGaussianResize2dNumpy.sk_weight, :150-160 (same at :504-514).
 * rho, sx, sy are float32 values promoted to double exactly as numpy does. */
static inline double sk_weight(float rho, float sx, float sy, double x, double y) {
  const double sxx = (double)sx * x;
  const double x_nominal = sxx * sxx;                 /* (sigma_x * x) ** 2 */
  const double syy = (double)sy * y;
  const double y_nominal = syy * syy;
  const double xy_nominal = sxx * (double)sy * y;     /* sigma_x * x * sigma_y * y */
  const float two_rho = 2.0f * rho;                   /* 2 * rho stays float32 */
  const double exp_term = -0.5 * (x_nominal - (double)two_rho * xy_nominal + y_nominal);
  return exp(exp_term);
}

/* AmplifiedLinearResize2dNumpy.linear_alpha / linear_weight, :233-241 (same at :587-595). */
static inline double linear_alpha(double x, float alpha) {
  const double a = (double)alpha;
  return (a * x + 1.0) * (double)((-1.0 <= x) && (x < 0.0)) +
         (1.0 - a * x) * (double)((0.0 <= x) && (x <= 1.0));
}
static inline double linear_weight(float alpha, double x, double y) {
  double lx = linear_alpha(x, alpha), ly = linear_alpha(y, alpha);
  lx = lx < 0.0 ? 0.0 : lx; /* np.clip(., 0, None) */
  ly = ly < 0.0 ? 0.0 : ly;
  return lx * ly;
}

/* hyper decode in float32: rho = h*2-1, sigma = h*max_sigma (:168-170, :249-250) */
static inline float dec_rho(float h) { float t = h * 2.0f; return t - 1.0f; }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One axis of Resize2dNumpy.get_distance (:106-140) for support 2:
 * projected grid (:57-80), left boundary (:82-98), pad (:100-104).
 * p[o] is the padded projected coordinate, left[o] the padded left tap. */
static int sr_axis(int in_sz, int out_sz, double scale, int supp, double* p, int* left,
                   int* pad0, int* pad1) {
  for (int o = 0; o < out_sz; ++o) {
    const double g = (double)o / scale + (double)(in_sz - 1) / 2 - (double)(out_sz - 1) / (2 * scale);
    p[o] = g;
    left[o] = (int)ceil(g - (double)supp / 2 - EPS32);
  }
  *pad0 = -left[0];
  *pad1 = left[out_sz - 1] + (supp - 1) - in_sz + 1;
  if (*pad0 < 0 || *pad1 < 0) return 3; /* np.pad would raise on a negative width */
  for (int o = 0; o < out_sz; ++o) { left[o] += *pad0; p[o] += (double)*pad0; }
  return 0;
}

/* np.pad source index for padded position ip of an axis of length n padded by pad0 in front:
 * mode 0 'constant' (-1 = outside), 1 'edge', 2 'reflect', 3 'symmetric', 4 'wrap' (numpy.pad semantics). */
static int pad_src(int ip, int pad0, int n, int mode) {
  int i = ip - pad0;
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

/* SteeringGaussianResize2dNumpy.resize (:162-223) when kind == 0, with hypers
 * h0=rho, h1=sigma_x, h2=sigma_y in [0,1]; AmplifiedLinearResize2dNumpy.resize
 * (:243-282) when kind == 1, with h0 = alpha (h1, h2 ignored).  Any support size (the
 * caller applies the antialias growth of :51-55), np.pad mode `pad_mode` for the image, 'edge'
 * for the hypers; aa_scale = min_scale_factor (:186-193: the Gaussian kind scales its
 * distances by it when antialiasing; its weight factor cancels in the normalisation), 1 otherwise.
 *   img, h* : float32 [C][H][W];   out : double [C][oH][oW]
 * (oH, oW) = ceil(scale * in) is computed by the caller as at :41-45. */
int lerf_oracle_resize_sr_ex(int kind, const float* img, const float* h0, const float* h1,
                             const float* h2, int C, int H, int W, double scale_h,
                             double scale_w, int oH, int oW, float max_sigma, int supp,
                             int pad_mode, double aa_scale, double* out) {
  if (supp < 1 || supp > 64) return 2;
  double* px = (double*)malloc(sizeof(double) * oH);
  double* py = (double*)malloc(sizeof(double) * oW);
  int* lx = (int*)malloc(sizeof(int) * oH);
  int* ly = (int*)malloc(sizeof(int) * oW);
  int p0x, p1x, p0y, p1y, rc;
  rc = sr_axis(H, oH, scale_h, supp, px, lx, &p0x, &p1x);
  if (!rc) rc = sr_axis(W, oW, scale_w, supp, py, ly, &p0y, &p1y);
  if (rc) { free(px); free(py); free(lx); free(ly); return rc; }
  const int n = supp * supp;
#pragma omp parallel for collapse(2) schedule(static)
  for (int c = 0; c < C; ++c) {
    for (int ox = 0; ox < oH; ++ox) {
      const size_t pl = (size_t)c * H * W;
      double* wts = (double*)malloc(sizeof(double) * 2 * n);
      double* nb = wts + n;
      for (int oy = 0; oy < oW; ++oy) {
        /* flattened patch order a*supp+b: row tap b, column tap a (np.meshgrid 'xy', :95-98) */
        for (int a = 0; a < supp; ++a)
          for (int b = 0; b < supp; ++b) {
            const int fx = lx[ox] + b, fy = ly[oy] + a;       /* padded tap index */
            const double dx = px[ox] - (double)fx, dy = py[oy] - (double)fy; /* :131-134 */
            const int sx_ = fx - p0x, sy_ = fy - p0y;         /* un-padded source index */
            const int cx = clampi(sx_, 0, H - 1), cy = clampi(sy_, 0, W - 1); /* 'edge', :172-174 */
            const size_t hi = pl + (size_t)cx * W + cy;
            double wgt;
            if (kind == 0)
              wgt = sk_weight(dec_rho(h0[hi]), h1[hi] * max_sigma, h2[hi] * max_sigma, aa_scale * dx, aa_scale * dy);
            else
              wgt = linear_weight(max_sigma * dec_rho(h0[hi]), dx, dy);
            wts[a * supp + b] = wgt;
            const int ix = pad_src(fx, p0x, H, pad_mode), iy = pad_src(fy, p0y, W, pad_mode);
            nb[a * supp + b] = (ix >= 0 && iy >= 0) ? (double)img[pl + (size_t)ix * W + iy] : 0.0; /* :208 */
          }
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += wts[k];                 /* :205 */
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += nb[k] * (wts[k] / s); /* :206, :220-221 */
        out[((size_t)c * oH + ox) * oW + oy] = acc;
      }
      free(wts);
    }
  }
  free(px); free(py); free(lx); free(ly);
  return 0;
}

/* The default configuration of the eval scripts: support 2, 'constant' image pad, no antialias. */
int lerf_oracle_resize_sr(int kind, const float* img, const float* h0, const float* h1,
                          const float* h2, int C, int H, int W, double scale_h,
                          double scale_w, int oH, int oW, float max_sigma, double* out) {
  return lerf_oracle_resize_sr_ex(kind, img, h0, h1, h2, C, H, W, scale_h, scale_w, oH, oW, max_sigma, 2, 0, 1.0, out);
}

/* Warp2dNumpy.get_projected_grid2d (:306-342) for one output pixel: the inverse
 * homography on float32 output coordinates, perspective divide, flip back to
 * (row, col), clip to [0, in].  minv = np.linalg.inv(matrix), computed by the
 * caller with numpy exactly as the reference does (:327). */
static inline void warp_project(const double* minv, int ox, int oy, int H, int W,
                                double* px, double* py) {
  const double x = (double)(float)oy, y = (double)(float)ox; /* h -> y, w -> x, :322-325 */
  const double g0 = minv[0] * x + minv[1] * y + minv[2] * 1.0;
  const double g1 = minv[3] * x + minv[4] * y + minv[5] * 1.0;
  const double g2 = minv[6] * x + minv[7] * y + minv[8] * 1.0;
  const double xi = g0 / g2, yi = g1 / g2;                   /* :330-331 */
  double r = yi, c = xi;                                     /* reverse back, :335 */
  r = r < 0.0 ? 0.0 : (r > (double)H ? (double)H : r);       /* .clip(0, in_sz[0]), :338 */
  c = c < 0.0 ? 0.0 : (c > (double)W ? (double)W : c);
  *px = r; *py = c;
}

static inline int warp_left(double p, int supp) {
  return (int)ceil(p - (double)supp / 2 - EPS32); /* :347-352 */
}

/* Warp2dNumpy.calc_pad_sz (:363-369): pads come from fov[0,0] and fov[-1,-1] only. */
static void warp_pads(const double* minv, int H, int W, int oH, int oW, int supp,
                      int* p0x, int* p0y) {
  double px, py;
  warp_project(minv, 0, 0, H, W, &px, &py);
  const int l0x = warp_left(px, supp), l0y = warp_left(py, supp);
  *p0x = -l0x > 0 ? -l0x : 0;
  *p0y = -l0y > 0 ? -l0y : 0;
  (void)oH; (void)oW; /* the trailing pad only sizes np.pad's output; taps are clipped to in-1 */
}

/* SteeringGaussianWarp2dNumpy.warp (:516-577) for kind 0, AmplifiedLinearWarp2dNumpy.warp
 * (:597-635) for kind 1, on the geometry of Warp2dNumpy.get_distance (:371-407):
 * taps clipped to [0, in-1] in PADDED coordinates after the pad shift (:396-398),
 * distances against the clipped taps (:400-403).  0/0 gives NaN as in numpy.
 * Any support size; pad_mode = the np.pad mode of the image (only the LEADING pad can be
 * reached: taps are clipped to in-1), the hypers replicate. */
int lerf_oracle_warp_ex(int kind, const float* img, const float* h0, const float* h1,
                        const float* h2, int C, int H, int W, const double* minv, int oH,
                        int oW, float max_sigma, int supp, int pad_mode, double* out) {
  if (supp < 1 || supp > 64) return 2;
  int p0x, p0y;
  warp_pads(minv, H, W, oH, oW, supp, &p0x, &p0y);
  const int n = supp * supp;
#pragma omp parallel for schedule(static)
  for (int ox = 0; ox < oH; ++ox) {
    double* wts = (double*)malloc(sizeof(double) * 2 * n);
    double* nb = wts + n;
    for (int oy = 0; oy < oW; ++oy) {
      double px, py;
      warp_project(minv, ox, oy, H, W, &px, &py);
      const int lx = warp_left(px, supp) + p0x, ly = warp_left(py, supp) + p0y; /* :366 */
      px += (double)p0x; py += (double)p0y;                                      /* :367 */
      for (int c = 0; c < C; ++c) {
        const size_t pl = (size_t)c * H * W;
        for (int a = 0; a < supp; ++a)
          for (int b = 0; b < supp; ++b) {
            const int fx = clampi(lx + b, 0, H - 1), fy = clampi(ly + a, 0, W - 1); /* :397-398 */
            const double dx = px - (double)fx, dy = py - (double)fy;
            const int sx_ = fx - p0x, sy_ = fy - p0y;
            const int cx = clampi(sx_, 0, H - 1), cy = clampi(sy_, 0, W - 1);
            const size_t hi = pl + (size_t)cx * W + cy;
            double wgt;
            if (kind == 0)
              wgt = sk_weight(dec_rho(h0[hi]), h1[hi] * max_sigma, h2[hi] * max_sigma, dx, dy);
            else
              wgt = linear_weight(max_sigma * dec_rho(h0[hi]), dx, dy);
            wts[a * supp + b] = wgt;
            const int ix = pad_src(fx, p0x, H, pad_mode), iy = pad_src(fy, p0y, W, pad_mode);
            nb[a * supp + b] = (ix >= 0 && iy >= 0) ? (double)img[pl + (size_t)ix * W + iy] : 0.0;
          }
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += wts[k];
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += nb[k] * (wts[k] / s);
        out[((size_t)c * oH + ox) * oW + oy] = acc;
      }
    }
    free(wts);
  }
  return 0;
}

int lerf_oracle_warp(int kind, const float* img, const float* h0, const float* h1,
                     const float* h2, int C, int H, int W, const double* minv, int oH,
                     int oW, float max_sigma, double* out) {
  return lerf_oracle_warp_ex(kind, img, h0, h1, h2, C, H, W, minv, oH, oW, max_sigma, 2, 0, out);
}

/* NearestWarp2dNumpy (:460-467): support 1, box2d weight (interp_methods.py:67-70,
 * 83-85) through Warp2dNumpy.warp (:409-449).  out = img*w/w -> img value or NaN. */
int lerf_oracle_nearest_warp(const float* img, int C, int H, int W, const double* minv,
                             int oH, int oW, double* out) {
  const int supp = 1;
  int p0x, p0y;
  warp_pads(minv, H, W, oH, oW, supp, &p0x, &p0y);
#pragma omp parallel for schedule(static)
  for (int ox = 0; ox < oH; ++ox) {
    for (int oy = 0; oy < oW; ++oy) {
      double px, py;
      warp_project(minv, ox, oy, H, W, &px, &py);
      const int fx = clampi(warp_left(px, supp) + p0x, 0, H - 1);
      const int fy = clampi(warp_left(py, supp) + p0y, 0, W - 1);
      const double dx = (px + (double)p0x) - (double)fx, dy = (py + (double)p0y) - (double)fy;
      const double bx = (double)((-1.0 <= dx) && (dx < 0.0)) + (double)((0.0 <= dx) && (dx <= 1.0));
      const double by = (double)((-1.0 <= dy) && (dy < 0.0)) + (double)((0.0 <= dy) && (dy <= 1.0));
      const double wgt = bx * by;
      const double wn = wgt / wgt; /* weights / weights_patch_sum, :429-431 (NaN for 0/0) */
      const int sx_ = fx - p0x, sy_ = fy - p0y;
      for (int c = 0; c < C; ++c) {
        const double v = (sx_ >= 0 && sy_ >= 0) ? (double)img[((size_t)c * H + sx_) * W + sy_] : 0.0;
        out[((size_t)c * oH + ox) * oW + oy] = v * wn;
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------
 * Fixed-kernel warps: Bilinear / Bicubic / Lanczos2 / Lanczos3Warp2dNumpy
 * (resize_right2d_numpy.py:451-494) = Warp2dNumpy.warp (:409-449) with the separable kernels of
 * resize_right/interp_methods.py:32-100.  kernel: 0 box (support 1), 1 linear (2), 2 cubic (4),
 * 3 lanczos2 (4), 4 lanczos3 (6).
 * ------------------------------------------------------------------------- */
static const double LERF_PI = 3.141592653589793;

static inline double k_cubic(double x) { /* interp_methods.py:32-41 */
  const double a = fabs(x), a2 = a * a, a3 = a * a * a;
  double r = 0.0;
  if (a <= 1.0) r += 1.5 * a3 - 2.5 * a2 + 1.0;
  if (1.0 < a && a <= 2.0) r += -0.5 * a3 + 2.5 * a2 - 4.0 * a + 2.0;
  return r;
}
static inline double k_lanczos(double x, double n) { /* :44-55: ((sin(pi x) sin(pi x / n) + eps) / (pi^2 x^2 / n + eps)) [|x| < n] */
  const double v = (sin(LERF_PI * x) * sin(LERF_PI * x / n) + EPS32) / ((LERF_PI * LERF_PI * x * x / n) + EPS32);
  return fabs(x) < n ? v : 0.0;
}
static inline double k_linear(double x) { /* :58-62 */
  return (x + 1.0) * (double)((-1.0 <= x) && (x < 0.0)) + (1.0 - x) * (double)((0.0 <= x) && (x <= 1.0));
}
static inline double k_box(double x) { /* :65-68 */
  return (double)((-1.0 <= x) && (x < 0.0)) + (double)((0.0 <= x) && (x <= 1.0));
}
static inline double k_eval(int kernel, double x) {
  switch (kernel) {
    case 0: return k_box(x);
    case 1: return k_linear(x);
    case 2: return k_cubic(x);
    case 3: return k_lanczos(x, 2.0);
    default: return k_lanczos(x, 3.0);
  }
}

int lerf_oracle_fixed_warp_support(int kernel) {
  static const int S[5] = {1, 2, 4, 4, 6};
  return (kernel < 0 || kernel > 4) ? -1 : S[kernel];
}

int lerf_oracle_fixed_warp(int kernel, const float* img, int C, int H, int W, const double* minv,
                           int oH, int oW, double* out) {
  const int supp = lerf_oracle_fixed_warp_support(kernel);
  if (supp < 0) return 1;
  int p0x, p0y;
  warp_pads(minv, H, W, oH, oW, supp, &p0x, &p0y);
#pragma omp parallel for schedule(static)
  for (int ox = 0; ox < oH; ++ox) {
    for (int oy = 0; oy < oW; ++oy) {
      double px, py;
      warp_project(minv, ox, oy, H, W, &px, &py);
      const int lx = warp_left(px, supp) + p0x, ly = warp_left(py, supp) + p0y; /* :366 */
      px += (double)p0x; py += (double)p0y;                                      /* :367 */
      double wx[6], wy[6];
      int sx[6], sy[6];
      for (int k = 0; k < supp; ++k) {
        const int fx = clampi(lx + k, 0, H - 1), fy = clampi(ly + k, 0, W - 1);  /* :397-398 */
        wx[k] = k_eval(kernel, px - (double)fx);                                 /* :400-403, weight(dis_x, dis_y) */
        wy[k] = k_eval(kernel, py - (double)fy);
        sx[k] = fx - p0x; sy[k] = fy - p0y;                                      /* index into the un-padded image */
      }
      /* patch element (i, j): row tap j, column tap i (the meshgrid of :357 is 'xy'); sum over the patch (:428) */
      double s = 0.0;
      for (int i = 0; i < supp; ++i)
        for (int j = 0; j < supp; ++j) s += wx[j] * wy[i];
      for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int i = 0; i < supp; ++i)
          for (int j = 0; j < supp; ++j) {
            const double v = (sx[j] >= 0 && sy[i] >= 0) ? (double)img[((size_t)c * H + sx[j]) * W + sy[i]] : 0.0;
            acc += v * ((wx[j] * wy[i]) / s);                                    /* :431, :446-447 */
          }
        out[((size_t)c * oH + ox) * oW + oy] = acc;
      }
    }
  }
  return 0;
}

// PNG writer on the device for the result images of the LeRF path (sm_100a) -- SURVEY.md 8f item 2, the output side of the
// image I/O step: the reference hands every result to PIL on the host (resample/eval_lut_sr.py:667-708, Image.save), which
// deflates a 2K x4 result for seconds per image.  Here the uint8 HWC image that the resampler left in device memory is
// wrapped into a complete, valid PNG file ON THE GPU -- filter type 0 per scanline, a zlib stream of STORED deflate blocks
// (no compression), Adler-32 and CRC-32 computed by parallel kernels -- so the host only copies the bytes to disk.
// HBM-bound byte work: one pass over the image for the payload, one for Adler-32, one over the file for CRC-32.
//
// File layout (H rows of W pixels, C = 1, 2, 3, 4 channels of 8 bits; row = 1 + W*C, R = H*row raw bytes,
// nblk = ceil(R / 65535) stored blocks, Z = 2 + 5*nblk + R + 4 bytes of zlib stream):
//   [0,8) signature   [8,33) IHDR chunk   [33,37) IDAT length   [37,41) "IDAT"   [41,41+Z) zlib   [41+Z,45+Z) IDAT CRC
//   [45+Z,57+Z) IEND chunk
// CRC-32 of a long range = XOR over segments of crc(segment) * x^(8 * bytes after it) mod P (the identity behind zlib's
// crc32_combine); Adler-32: A = 1 + sum d_i, B = R + sum (R - i) d_i (mod 65521) straight from the image.
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace lerf {
namespace {

constexpr uint32_t kPoly = 0xedb88320u;
constexpr int kSeg = 1024;           // bytes of the file per CRC thread (aligned to kSeg in the file)
constexpr uint32_t kAdlerMod = 65521u;

// a * b mod P over GF(2), reflected representation (x^0 = 0x80000000)
__host__ __device__ inline uint32_t multmodp(uint32_t a, uint32_t b) {
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ kPoly : b >> 1;
  }
  return p;
}

struct X2n {
  uint32_t t[32];  // x^(2^k) mod P
};

// x^(n * 2^k) mod P
__host__ __device__ inline uint32_t x2nmodp(const X2n& x, unsigned long long n, unsigned k) {
  uint32_t p = 1u << 31;
  while (n) {
    if (n & 1) p = multmodp(x.t[k & 31], p);
    n >>= 1;
    ++k;
  }
  return p;
}

inline uint32_t crc32_host(const uint8_t* d, size_t n) {
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kPoly : c >> 1;
  }
  return ~c;
}

struct PngGeom {
  long long row, R, Z, nblk;       // bytes per raw scanline, raw bytes, zlib bytes, stored blocks
  long long zoff, file;            // start of the zlib stream, file size
  uint8_t head[41];                // signature + IHDR chunk + IDAT length + "IDAT"
  uint8_t tail[12];                // IEND chunk
};

// File byte at position p for every p but the Adler-32 and the IDAT CRC fields (written by the finalisers).  Z < 2^31, so
// positions inside the zlib stream are 32-bit (the division by the block size is by a constant, the one by the scanline
// length is a 32-bit division).
__device__ __forceinline__ uint8_t file_byte(const PngGeom& g, const uint8_t* __restrict__ img, long long p) {
  if (p < 41) return g.head[p];
  if (p >= g.zoff + g.Z) {
    const long long t = p - (g.zoff + g.Z + 4);
    return t >= 0 && t < 12 ? g.tail[t] : 0;
  }
  uint32_t j = (uint32_t)(p - g.zoff);
  if (j < 2) return j == 0 ? 0x78 : 0x01;  // zlib header: deflate, 32 KiB window, no preset dictionary, check bits
  j -= 2;
  const uint32_t blk = j / 65540u, o = j - blk * 65540u, nblk = (uint32_t)g.nblk, R = (uint32_t)g.R;
  if (blk >= nblk) return 0;  // the Adler-32 field
  if (o < 5) {
    const uint32_t l = min(65535u, R - blk * 65535u), nl = ~l & 0xffffu;
    return o == 0 ? (blk == nblk - 1 ? 1 : 0) : (o == 1 ? l & 255u : (o == 2 ? l >> 8 : (o == 3 ? nl & 255u : nl >> 8)));
  }
  const uint32_t r = blk * 65535u + (o - 5);
  if (r >= R) return 0;  // the Adler-32 field after a short last block
  const uint32_t row = (uint32_t)g.row, y = r / row, c = r - y * row;
  return c == 0 ? 0 : __ldg(img + (size_t)y * (row - 1) + (c - 1));  // filter type 0 (None), then the scanline
}

// One aligned 32-bit word of the file per thread.
__global__ void __launch_bounds__(256) png_fill_kernel(const PngGeom g, const uint8_t* __restrict__ img, uint32_t* __restrict__ out) {
  const long long words = (g.file + 3) >> 2;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
    const long long p = w << 2;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v |= (uint32_t)(p + k < g.file ? file_byte(g, img, p + k) : 0) << (8 * k);
    out[w] = v;
  }
}

// Adler-32 partial sums over the image: acc[0] += sum d, acc[1] += sum (R - i) d (both mod 65521), i = raw index of the byte
// = image index e + y + 1 (y + 1 filter bytes precede it).  A thread takes 64 image bytes (four 128-bit loads when the run
// is aligned and does not cross a scanline: sum (w0 - k) d_k = w0 * s1 - sum k d_k, all 32-bit but one product).
__global__ void __launch_bounds__(256) png_adler_kernel(const PngGeom g, const uint8_t* __restrict__ img, long long n,
                                                        unsigned long long* __restrict__ acc) {
  constexpr int kRun = 64;
  unsigned long long a1 = 0, a2 = 0;
  const long long wc = g.row - 1;
  const bool aligned = ((uintptr_t)img & 15) == 0;
  for (long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kRun; e0 < n; e0 += (long long)gridDim.x * blockDim.x * kRun) {
    long long y = e0 / wc, c = e0 - y * wc;
    const long long e1 = min(e0 + kRun, n);
    if (aligned && e1 - e0 == kRun && c + kRun <= wc) {
      uint32_t s1 = 0, sk = 0;
#pragma unroll
      for (int q = 0; q < kRun / 16; ++q) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + e0) + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t d = (w[k >> 2] >> (8 * (k & 3))) & 255u;
          s1 += d;
          sk += (uint32_t)(q * 16 + k) * d;
        }
      }
      const unsigned long long w0 = (unsigned long long)(g.R - (e0 + y + 1));  // weight of the run's first byte
      a1 += s1;
      a2 += (w0 % kAdlerMod) * s1 + (unsigned long long)kAdlerMod * 4096ull - sk;  // sk <= 64 * 63 * 255 < 4096 * 65521
    } else {
      unsigned long long s1 = 0, s2 = 0;
      for (long long e = e0; e < e1; ++e) {
        const unsigned d = __ldg(img + e);
        s1 += d;
        s2 += (unsigned long long)(g.R - (e + y + 1)) * d;
        if (++c == wc) { c = 0; ++y; }
      }
      a1 += s1;
      a2 += s2 % kAdlerMod;
    }
    a2 %= kAdlerMod;
  }
  a1 %= kAdlerMod;
  for (int o = 16; o; o >>= 1) {
    a1 += __shfl_down_sync(0xffffffffu, a1, o);
    a2 += __shfl_down_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, a1);
    atomicAdd(acc + 1, a2);
  }
}

__global__ void png_adler_final_kernel(const PngGeom g, const unsigned long long* __restrict__ acc, uint8_t* __restrict__ out) {
  const uint32_t a = (uint32_t)((1 + acc[0] % kAdlerMod) % kAdlerMod);
  const uint32_t b = (uint32_t)((g.R % kAdlerMod + acc[1] % kAdlerMod) % kAdlerMod);
  uint8_t* f = out + g.zoff + g.Z - 4;
  f[0] = b >> 8; f[1] = b & 255u; f[2] = a >> 8; f[3] = a & 255u;  // big-endian (B << 16 | A)
}

// CRC-32 of the file range [lo, hi): one thread per kSeg-aligned segment, combined by crc_i * x^(8 * bytes after segment i).
__global__ void __launch_bounds__(256) png_crc_kernel(const uint8_t* __restrict__ file, long long lo, long long hi, const X2n x2n,
                                                      uint32_t* __restrict__ acc) {
  __shared__ uint32_t tab[256];
  {
    uint32_t c = threadIdx.x;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kPoly : c >> 1;
    tab[threadIdx.x] = c;
  }
  __syncthreads();
  const long long seg0 = lo / kSeg, nseg = (hi - 1) / kSeg - seg0 + 1;
  uint32_t part = 0;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += (long long)gridDim.x * blockDim.x) {
    const long long a = max(lo, (seg0 + s) * kSeg), b = min(hi, (seg0 + s + 1) * kSeg);
    uint32_t c = 0xffffffffu;
    long long p = a;
    for (; p < b && (p & 15); ++p) c = tab[(c ^ file[p]) & 255u] ^ (c >> 8);
    for (; p + 16 <= b; p += 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(file + p);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        c ^= w[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) c = tab[c & 255u] ^ (c >> 8);
      }
    }
    for (; p < b; ++p) c = tab[(c ^ file[p]) & 255u] ^ (c >> 8);
    c = ~c;
    const long long after = hi - b;
    part ^= after ? multmodp(x2nmodp(x2n, (unsigned long long)after, 3), c) : c;
  }
  for (int o = 16; o; o >>= 1) part ^= __shfl_down_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0 && part) atomicXor(acc, part);
}

__global__ void png_crc_final_kernel(const uint32_t* __restrict__ acc, uint8_t* __restrict__ out, long long off) {
  const uint32_t c = *acc;
  out[off] = c >> 24; out[off + 1] = (c >> 16) & 255u; out[off + 2] = (c >> 8) & 255u; out[off + 3] = c & 255u;
}

int make_geom(int H, int W, int C, PngGeom& g) {
  if (H < 1 || W < 1 || C < 1 || C > 4) return fail(LERF_EINVAL, "png: bad image shape %d x %d x %d", H, W, C);
  g.row = 1 + (long long)W * C;
  g.R = (long long)H * g.row;
  g.nblk = (g.R + 65534) / 65535;
  g.Z = 2 + 5 * g.nblk + g.R + 4;
  if (g.Z > 0x7fffffffLL) return fail(LERF_EUNSUPPORTED, "png: %lld bytes of image data do not fit one IDAT chunk", g.Z);
  g.zoff = 41;
  g.file = 41 + g.Z + 4 + 12;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  static const uint8_t ctype[5] = {0, 0, 4, 2, 6};  // grey, grey + alpha, RGB, RGBA
  memcpy(g.head, sig, 8);
  uint8_t* h = g.head + 8;
  auto be32 = [](uint8_t* d, uint32_t v) { d[0] = v >> 24; d[1] = (v >> 16) & 255u; d[2] = (v >> 8) & 255u; d[3] = v & 255u; };
  be32(h, 13);
  memcpy(h + 4, "IHDR", 4);
  be32(h + 8, (uint32_t)W);
  be32(h + 12, (uint32_t)H);
  h[16] = 8; h[17] = ctype[C]; h[18] = 0; h[19] = 0; h[20] = 0;
  be32(h + 21, crc32_host(h + 4, 17));
  be32(g.head + 33, (uint32_t)g.Z);
  memcpy(g.head + 37, "IDAT", 4);
  be32(g.tail, 0);
  memcpy(g.tail + 4, "IEND", 4);
  be32(g.tail + 8, crc32_host(g.tail + 4, 4));
  return LERF_OK;
}

}  // namespace
}  // namespace lerf

using namespace lerf;

extern "C" {

long long lerf_png_stored_bytes(int H, int W, int channels) {
  PngGeom g;
  if (make_geom(H, W, channels, g) != LERF_OK) return -1;
  return g.file;
}

int lerf_png_encode_stored(const uint8_t* img, int H, int W, int channels, uint8_t* png, long long png_capacity, void* scratch32,
                           lerf_stream_t stream) {
  if (!img || !png || !scratch32) return fail(LERF_EINVAL, "lerf_png_encode_stored: null pointer");
  PngGeom g;
  int rc = make_geom(H, W, channels, g);
  if (rc) return rc;
  if (png_capacity < ((g.file + 3) & ~3LL)) return fail(LERF_EINVAL, "lerf_png_encode_stored: output holds %lld bytes, the file rounded up to 4 needs %lld", png_capacity, (g.file + 3) & ~3LL);
  if (((uintptr_t)png & 15) || ((uintptr_t)scratch32 & 7)) return fail(LERF_EINVAL, "lerf_png_encode_stored: png must be 16-byte aligned, scratch 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  X2n x2n;
  x2n.t[0] = 0x40000000u;
  for (int k = 1; k < 32; ++k) x2n.t[k] = multmodp(x2n.t[k - 1], x2n.t[k - 1]);
  LERF_CUDA(cudaMemsetAsync(scratch32, 0, 32, st));
  unsigned long long* adler = (unsigned long long*)scratch32;
  uint32_t* crc = (uint32_t*)((uint8_t*)scratch32 + 16);
  const long long words = (g.file + 3) >> 2, n = (long long)H * W * channels;
  const int sm = 148;
  png_fill_kernel<<<(unsigned)std::min<long long>((words + 255) / 256, sm * 16), 256, 0, st>>>(g, img, (uint32_t*)png);
  LERF_LAUNCHED();
  png_adler_kernel<<<(unsigned)std::min<long long>((n + 64 * 256 - 1) / (64 * 256), sm * 16), 256, 0, st>>>(g, img, n, adler);
  LERF_LAUNCHED();
  png_adler_final_kernel<<<1, 1, 0, st>>>(g, adler, png);
  LERF_LAUNCHED();
  const long long lo = 37, hi = g.zoff + g.Z;  // chunk type + data
  const long long nseg = (hi - 1) / kSeg - lo / kSeg + 1;
  png_crc_kernel<<<(unsigned)std::min<long long>((nseg + 255) / 256, sm * 8), 256, 0, st>>>(png, lo, hi, x2n, crc);
  LERF_LAUNCHED();
  png_crc_final_kernel<<<1, 1, 0, st>>>(crc, png, hi);
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // extern "C"

// "Max-tap block" LUT format for the oC = 3 stage-2 tables of the LeRF hot path (sm_100a): lookup primitive, host
// repack and the stage kernel body.  Replaces the stage-2 loop resample/eval_lut_sr.py:579-628 over
// FourSimplexInterpFaster (:24-470) of the reference.
//
// Why.  Measured on B200 (scripts/microbench/l1gather.cu): a warp-wide scattered load that hits L1 costs ~13 cycles
// for 4, 8 or 16 bytes per lane and ~19 cycles for 32 bytes per lane.  A row-major lookup is five such loads per
// channel triple (65 cycles); a 48-byte cell (lut_cell.cuh) is three (39 cycles) but 2 sectors and 16x the
// footprint.  This format needs ONE 32-byte load per lookup: the 4-D cell is split by WHICH TAP HAS THE LARGEST
// LSB (t1).  Of the 16 corners only 9 can be simplex vertices once t1 is known -- corner 0, {t1}, the three
// {t1,x}, the three "all but z", and corner 15 -- i.e. 27 bytes for three channels.
//
// Block (cell, t1), 8 words:
//   word 2c   (channel c): byte x = corner {t1,x} for x != t1,  byte t1 = corner 0      (vertex 2 candidates, vertex 0)
//   word 2c+1 (channel c): byte z = corner 15-{z} for z != t1,  byte t1 = corner 15     (vertex 3 candidates, vertex 4)
//   word 6               : bytes (c0, c1, c2, 0) of corner {t1}                        (vertex 1)
// so PRMT(word 2c, word 2c+1, [t2, 4+t4, t1, 4+t1]) = [P2, P3, P0, P4] and one DP4A blends them.
// Tap ids: a=0, b=1, c=2, d=3; corner bit of tap t is 8 >> t.
//
// The kernel walks a 32 x (8*NJ) tile TABLE BY TABLE (all pixels' two passes on table 0, then table 1, ...), so
// at any time a CTA's loads fall in one table's blocks and L1 keeps them between neighbouring pixels.
#pragma once
#include <stdint.h>

#include "lut_cell.cuh"

namespace lerf {
namespace mt {

using cell::prmt;
using cell::dp4a_ss;

constexpr int kBlockBytes = 32;
constexpr size_t kTableBytes = (size_t)65536 * 4 * kBlockBytes;  // 8 MiB per table

struct Lookup {
  uint32_t block;  // cell * 4 + t1
  uint32_t sel;    // PRMT selector [t2, 4+t4, t1, 4+t1]
  uint32_t w;      // byte weights [w2, w3, w0, w4]
  uint32_t w1[3];  // w1 in byte lane c
};

// Pixel as kept in shared memory: .x = lsb << 24 (sort key base, low 16 bits free), .y = msb.
LERF_HD void split_px2(uint32_t v, uint32_t& key, uint32_t& msb) {
  key = (v & 15u) << 24;
  msb = v >> 4;
}

// Taps a, b, c, d in table-axis order.
LERF_HD Lookup prepare(uint32_t ka, uint32_t ma, uint32_t kb, uint32_t mb, uint32_t kc, uint32_t mc, uint32_t kd, uint32_t md) {
  Lookup L;
  const uint32_t cellidx = ((ma * 16u + mb) * 16u + mc) * 16u + md;
  // key = lsb<<24 | t*0x1111 + 0x4040: nibbles (t, 4+t, t, 4+t)
  int k1 = (int)(ka | 0x4040u), k2 = (int)(kb | 0x5151u), k3 = (int)(kc | 0x6262u), k4 = (int)(kd | 0x7373u);
  int t;
  t = cell::imax(k1, k2); k2 = cell::imin(k1, k2); k1 = t;
  t = cell::imax(k3, k4); k4 = cell::imin(k3, k4); k3 = t;
  t = cell::imax(k1, k3); k3 = cell::imin(k1, k3); k1 = t;
  t = cell::imax(k2, k4); k4 = cell::imin(k2, k4); k2 = t;
  t = cell::imax(k2, k3); k3 = cell::imin(k2, k3); k2 = t;
  L.block = cellidx * 4u + ((uint32_t)k1 & 3u);
  const uint32_t s24 = ((uint32_t)k2 & 0x000Fu) | ((uint32_t)k4 & ~0x000Fu);  // nibble 0 = t2, nibble 1 = 4+t4
  L.sel = (s24 & 0x00FFu) | ((uint32_t)k1 & ~0x00FFu);                        // nibbles 2,3 = t1, 4+t1
  // sorted lsbs f1 >= f2 >= f3 >= f4 sit in byte 3 of the keys
  const uint32_t f12 = prmt((uint32_t)k1, (uint32_t)k2, 0x0073u);   // [f1, f2, x, x]
  const uint32_t f34 = prmt((uint32_t)k3, (uint32_t)k4, 0x0073u);   // [f3, f4, x, x]
  const uint32_t F = prmt(f12, f34, 0x5410u);                       // [f1, f2, f3, f4]
  const uint32_t Wd = F - (F >> 8);                                 // [f1-f2, f2-f3, f3-f4, f4]   no borrows (sorted)
  const uint32_t G = 0x10101010u - F;                               // [16-f1, ...]                no borrows (f <= 15)
  L.w = prmt(Wd, G, 0x3421u);                                       // [w2, w3, w0, w4]
  L.w1[0] = Wd & 0xFFu;
  L.w1[1] = prmt(Wd, 0u, 0x4404u);
  L.w1[2] = prmt(Wd, 0u, 0x4044u);
  return L;
}

// Single-word form (production since r1e): taps as cell::split_px words (lsb << 24 | msb << 8, the stage-1 tile format),
// so a tap costs one 4-byte shared-memory word instead of a uint2.  The low byte of a sort key carries the tap id twice,
// (4+t) << 4 | t; the msb bits 8..11 ride along in the keys and only break lsb ties (tied vertices have weight 0).
// L.block is returned as a BYTE offset (cell * 128 + t1 * 32) in this form.
LERF_HD Lookup prepare1(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t wd) {
  Lookup L;
  const uint32_t acc = ((wa * 16u + wb) * 16u + wc) * 16u + wd;  // bits 8..23 = cell index (the lsb bytes land above)
  int k1 = (int)(wa | 0x40u), k2 = (int)(wb | 0x51u), k3 = (int)(wc | 0x62u), k4 = (int)(wd | 0x73u);
  int t;
  t = cell::imax(k1, k2); k2 = cell::imin(k1, k2); k1 = t;
  t = cell::imax(k3, k4); k4 = cell::imin(k3, k4); k3 = t;
  t = cell::imax(k1, k3); k3 = cell::imin(k1, k3); k1 = t;
  t = cell::imax(k2, k4); k4 = cell::imin(k2, k4); k2 = t;
  t = cell::imax(k2, k3); k3 = cell::imin(k2, k3); k2 = t;
  L.block = ((acc & 0x00FFFF00u) >> 1) + (((uint32_t)k1 & 0x30u) << 1);  // bits 4,5 of the key's low byte = t1
  const uint32_t s24 = ((uint32_t)k2 & 0x0Fu) | ((uint32_t)k4 & ~0x0Fu);  // byte 0: nibble 0 = t2, nibble 1 = 4+t4
  L.sel = prmt(s24, (uint32_t)k1, 0x4440u);                              // byte 1: nibbles 2,3 = t1, 4+t1
  const uint32_t f12 = prmt((uint32_t)k1, (uint32_t)k2, 0x0073u);   // [f1, f2, x, x]
  const uint32_t f34 = prmt((uint32_t)k3, (uint32_t)k4, 0x0073u);   // [f3, f4, x, x]
  const uint32_t F = prmt(f12, f34, 0x5410u);                       // [f1, f2, f3, f4]
  const uint32_t Wd = F - (F >> 8);                                 // [f1-f2, f2-f3, f3-f4, f4]
  const uint32_t G = 0x10101010u - F;                               // [16-f1, ...]
  L.w = prmt(Wd, G, 0x3421u);                                       // [w2, w3, w0, w4]
  L.w1[0] = Wd & 0xFFu;
  L.w1[1] = prmt(Wd, 0u, 0x4404u);
  L.w1[2] = prmt(Wd, 0u, 0x4044u);
  return L;
}

// q[0..7] = the block's 8 words.
LERF_HD void blend3(const uint32_t q[8], const Lookup& L, int& n0, int& n1, int& n2) {
  n0 = dp4a_ss(prmt(q[0], q[1], L.sel), L.w, dp4a_ss(q[6], L.w1[0], n0));
  n1 = dp4a_ss(prmt(q[2], q[3], L.sel), L.w, dp4a_ss(q[6], L.w1[1], n1));
  n2 = dp4a_ss(prmt(q[4], q[5], L.sel), L.w, dp4a_ss(q[6], L.w1[2], n2));
}

// Host: row-major T[17^4][3] (int8) -> max-tap blocks, dst = kTableBytes.
inline void repack_maxtap(const int8_t* T, uint8_t* dst) {
  for (int cellidx = 0; cellidx < 65536; ++cellidx) {
    const int msb[4] = {(cellidx >> 12) & 15, (cellidx >> 8) & 15, (cellidx >> 4) & 15, cellidx & 15};
    auto row_of = [&](int m) {  // corner mask m (bit 8>>t = tap t bumped) -> table row
      int row = 0;
      for (int k = 0; k < 4; ++k) row = row * 17 + msb[k] + ((m >> (3 - k)) & 1);
      return row;
    };
    for (int t1 = 0; t1 < 4; ++t1) {
      uint8_t* b = dst + ((size_t)cellidx * 4 + t1) * kBlockBytes;
      const int bit1 = 8 >> t1;
      for (int c = 0; c < 3; ++c) {
        for (int x = 0; x < 4; ++x) {
          const int lo = x == t1 ? 0 : (bit1 | (8 >> x));
          const int hi = x == t1 ? 15 : (15 & ~(8 >> x));
          b[8 * c + x] = (uint8_t)T[(size_t)row_of(lo) * 3 + c];
          b[8 * c + 4 + x] = (uint8_t)T[(size_t)row_of(hi) * 3 + c];
        }
        b[24 + c] = (uint8_t)T[(size_t)row_of(bit1) * 3 + c];
      }
      b[27] = 0;
      for (int i = 28; i < 32; ++i) b[i] = 0;
    }
  }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------
// stage-2 kernel body: 32 x (8*NJ) tile, table by table
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTX = 32, kHalo = 3, kPitch = 40;  // uint2 per pixel; 40 mod 16 = 8 keeps the 8x4 patch rows apart

struct MtTables {
  const uint8_t* t[6];  // s r0, s r1, c r0, c r1, t r0, t r1
};

template <int MODE, int R, int K>
struct Tap {  // mode pattern (eval_lut_sr.py:30-81) composed with the rotation
  static constexpr int di = MODE == 0 ? (K >> 1) : (MODE == 1 ? 0 : K);
  static constexpr int dj = MODE == 0 ? (K & 1) : K;
  static constexpr int dy = R == 0 ? di : (R == 1 ? dj : (R == 2 ? -di : -dj));
  static constexpr int dx = R == 0 ? dj : (R == 1 ? -di : (R == 2 ? -dj : di));
};

// LD: how the 32-byte block is fetched.  0 = ld.global.nc (allocates in L1), 1 = nc + L1::no_allocate, 2 = ld.global.cg
// (L2 only), 3 = nc + L1::evict_first.
template <int LD>
__device__ __forceinline__ void load_block(const uint8_t* p, uint32_t q[8]) {
#define LERF_LD8(OP)                                                                                  \
  asm(OP " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                          \
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) \
      : "l"(p))
  if (LD == 1) LERF_LD8("ld.global.nc.L1::no_allocate.v8.u32");
  else if (LD == 2) LERF_LD8("ld.global.cg.v8.u32");
  else if (LD == 3) LERF_LD8("ld.global.nc.L1::evict_first.v8.u32");
  else LERF_LD8("ld.global.nc.v8.u32");
#undef LERF_LD8
}

template <int MODE, int R, int LD>
__device__ __forceinline__ void pass(const uint8_t* __restrict__ tab, const uint2* c, int& n0, int& n1, int& n2) {
  const uint2 a = c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx];
  const uint2 b = c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx];
  const uint2 cc = c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx];
  const uint2 d = c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx];
  const Lookup L = prepare(a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y);
  uint32_t q[8];
  load_block<LD>(tab + (size_t)L.block * kBlockBytes, q);
  blend3(q, L, n0, n1, n2);
}

template <int MODE, int R, int LD>
__device__ __forceinline__ void pass(const uint8_t* __restrict__ tab, const uint32_t* c, int& n0, int& n1, int& n2) {
  const Lookup L = prepare1(c[Tap<MODE, R, 0>::dy * kPitch + Tap<MODE, R, 0>::dx], c[Tap<MODE, R, 1>::dy * kPitch + Tap<MODE, R, 1>::dx],
                            c[Tap<MODE, R, 2>::dy * kPitch + Tap<MODE, R, 2>::dx], c[Tap<MODE, R, 3>::dy * kPitch + Tap<MODE, R, 3>::dx]);
  uint32_t q[8];
  load_block<LD>(tab + L.block, q);  // byte offset in the single-word form
  blend3(q, L, n0, n1, n2);
}

__device__ __forceinline__ void store_px(uint2* t, uint32_t v) {
  uint2 w;
  split_px2(v, w.x, w.y);
  *t = w;
}
__device__ __forceinline__ void store_px(uint32_t* t, uint32_t v) { *t = cell::split_px(v); }

__device__ __forceinline__ int rhe_div192(int num) {  // round_half_even(num / 192), num > 0
  const int t = num + 96;
  int q = t / 192;
  if (t - q * 192 == 0 && (q & 1)) --q;
  return q;
}

// tile: (8*NJ + 6) * kPitch pixels of shared memory, PX = uint32_t (single-word form, production) or uint2 (r1d form).
// (bxi, byi, p) = tile column, tile row, plane.  256 threads.
template <int NJ, int LD, unsigned TABMASK = 0x3Fu, typename PX = uint2>
__device__ __forceinline__ void lut_stage2_mt_body(const MtTables& t, const uint8_t* __restrict__ feat, int H, int W, int y0,
                                                   int y1, uint8_t* __restrict__ out, int bxi, int byi, int p, PX* tile) {
  constexpr int TY = 8 * NJ;
  const int bx = bxi * kTX, by = y0 + byi * TY;
  const uint8_t* src = feat + (long long)p * H * W;
  const int tid = threadIdx.x;
  for (int i = tid; i < (TY + 2 * kHalo) * (kTX + 2 * kHalo); i += 256) {
    const int r = i / (kTX + 2 * kHalo), c = i - r * (kTX + 2 * kHalo);
    const int gy = min(max(by + r - kHalo, 0), H - 1), gx = min(max(bx + c - kHalo, 0), W - 1);
    store_px(tile + r * kPitch + c, __ldcg(src + (long long)gy * W + gx));
  }
  __syncthreads();
  const int lane = tid & 31, wrp = tid >> 5;
  const int tx = (wrp & 3) * 8 + (lane & 7), ty = (wrp >> 2) * 4 + (lane >> 3);  // a warp = an 8x4 pixel patch
  const int x = bx + tx;
  if (x >= W) return;
  int acc[NJ][3];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0;
  const PX* c0 = tile + (ty + kHalo) * kPitch + tx + kHalo;
#define LERF_TAB(M, PAR)                                                          \
  if ((TABMASK >> (2 * M + PAR)) & 1u)                                            \
  _Pragma("unroll") for (int j = 0; j < NJ; ++j) {                                \
    if (by + ty + 8 * j < y1) {                                                   \
      const PX* c = c0 + 8 * j * kPitch;                                          \
      pass<M, PAR, LD>(t.t[2 * M + PAR], c, acc[j][0], acc[j][1], acc[j][2]);         \
      pass<M, PAR + 2, LD>(t.t[2 * M + PAR], c, acc[j][0], acc[j][1], acc[j][2]);     \
    }                                                                             \
  }
  LERF_TAB(0, 0) LERF_TAB(0, 1) LERF_TAB(1, 0) LERF_TAB(1, 1) LERF_TAB(2, 0) LERF_TAB(2, 1)
#undef LERF_TAB
  const long long ps = (long long)H * W;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int y = by + ty + 8 * j;
    if (y >= y1) continue;
    const long long o = ((long long)p * 3 * H + y) * W + x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int v = acc[j][k] + 127 * 192;
      __stcg(out + o + k * ps, (uint8_t)(v <= 0 ? 0 : min(rhe_div192(v), 255)));
    }
  }
}
#endif  // __CUDACC__

}  // namespace mt
}  // namespace lerf

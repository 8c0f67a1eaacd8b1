// Cell-packed LUT lookup primitive of the LeRF hot path (sm_100a), shared by the stage kernels in
// lut_cell.cu.  Replaces the body of FourSimplexInterpFaster (resample/eval_lut_sr.py:86-462 of the
// reference: 16 corner gathers + 24 boolean masks) for the rotation-ensembled stages.
//
// Idea.  The 17^4 table is repacked by CELL: the 16 corners of the 4-D cell (msb_a, msb_b, msb_c, msb_d)
// sit in one 16-byte block, so a lookup is ONE 128-bit load instead of five scattered byte loads.  The
// five simplex vertices are then picked out of the four registers with two PRMTs whose selectors fall
// out of the sort of the LSB keys, and blended with two DP4As whose byte weights are differences of
// the sorted LSBs.  All arithmetic is exact integer.
//
// Corner m = (ca<<3 | cb<<2 | cc<<1 | cd) is stored at byte  (popcount(m) odd ? 0 : 8) + (m & 7):
//   X = bytes 0..7  : the odd corners  -- vertex 1 (one tap bumped) and vertex 3 (three bumped)
//   Y = bytes 8..15 : the even corners -- vertex 0 (m=0 -> Y[0]), vertex 2 (two bumped), vertex 4 (m=15 -> Y[7])
// (m & 7 is injective on each parity class.)
//
// This header is also compiled by g++ for the CPU emulation test of the bit tricks (tests/csrc/): every
// device intrinsic used here has a host twin below.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LERF_HD __host__ __device__ __forceinline__
#else
#define LERF_HD inline
#endif

namespace lerf {
namespace cell {

// PTX prmt.b32, generic mode: result byte n = bytes{a,b}[sel nibble n & 7], or its sign replicated when
// the nibble's msb is set.
LERF_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
  return d;
#else
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int n = 0; n < 4; ++n) {
    const uint32_t nib = (s >> (4 * n)) & 15u;
    uint32_t byte = (uint32_t)(src >> (8 * (nib & 7u))) & 255u;
    if (nib & 8u) byte = (byte & 128u) ? 255u : 0u;
    d |= byte << (8 * n);
  }
  return d;
#endif
}

LERF_HD int dp4a_ss(uint32_t a, uint32_t b, int c) {  // signed bytes x signed bytes
#if defined(__CUDA_ARCH__)
  return __dp4a((int)a, (int)b, c);
#else
  for (int n = 0; n < 4; ++n) c += (int)(int8_t)(a >> (8 * n)) * (int)(int8_t)(b >> (8 * n));
  return c;
#endif
}

// (a & ~mask) | (c & mask) as ONE LOP3 (ptxas splits the two-immediate source form into two).
LERF_HD uint32_t bitsel(uint32_t a, uint32_t c, uint32_t mask) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xB8;" : "=r"(d) : "r"(a), "r"(mask), "r"(c));
  return d;
#else
  return (a & ~mask) | (c & mask);
#endif
}

LERF_HD int imax(int a, int b) { return a > b ? a : b; }
LERF_HD int imin(int a, int b) { return a < b ? a : b; }

// A pixel value v (0..255) as the kernels keep it in shared memory: lsb in byte 3, msb in bits 8..11.
// Bits 0..7 and 12..23 are zero: the low byte takes the tap's selector bits, byte 2 stays a zero source.
LERF_HD uint32_t split_px(uint32_t v) { return ((v & 15u) << 24) | ((v >> 4) << 8); }

struct Simplex {
  uint32_t cell;        // (msb_a<<12 | msb_b<<8 | msb_c<<4 | msb_d): index of the 16-byte block
  uint32_t selX, selY;  // PRMT selectors into (X0,X1) and (Y0,Y1)
  uint32_t wX, wY;      // byte weights matching the PRMT results:  [w1, w3, 0, 0] and [w2, w4, 0, w0]
};

// Block swizzle.  A 128-bit load is served in groups of 8 lanes; two lanes of a group that read different 128-byte
// lines at the same 16-byte slot collide in the L1 data banks.  Neighbouring pixels often share msb_d (the slot
// of the plain layout) while differing in the other msbs, which made every lookup ~4-way conflicted (ncu, r1c).
// The block of cell (a,b,c,d) is therefore stored at (cell & ~15) | ((d + ha*a + hb*b + hc*c) & 15): with odd
// weights every single-coordinate step, and every step along the diagonal, changes the slot.
struct Hash {
  uint32_t ha, hb, hc;  // 0,0,0 = plain layout
};

LERF_HD uint32_t swizzled_cell(uint32_t cellidx, const Hash& h) {
  const uint32_t a = (cellidx >> 12) & 15u, b = (cellidx >> 8) & 15u, c = (cellidx >> 4) & 15u, d = cellidx & 15u;
  return (cellidx & ~15u) | ((d + h.ha * a + h.hb * b + h.hc * c) & 15u);
}

// Taps a, b, c, d (split_px words) in table-axis order.
LERF_HD Simplex simplex_of(uint32_t xa, uint32_t xb, uint32_t xc, uint32_t xd, const Hash& h) {
  Simplex s;
  // msb fields land in bits 8..23; the lsb bytes only reach bits >= 24 (or overflow out).
  const uint32_t acc = ((xa * 16u + xb) * 16u + xc) * 16u + xd;
  // bits 8..11 of acc + ha*xa + hb*xb + hc*xc = (d + ha*a + hb*b + hc*c) mod 16: the X words are zero below bit 8
  const uint32_t mix = xa * h.ha + (xb * h.hb + (xc * h.hc + acc));
  s.cell = prmt(bitsel(acc, mix, 0xF00u), 0u, 0x4421u);
  // key = lsb<<24 | msb<<8 | (tap's corner bit & 7) replicated in nibbles 0 and 1.  Sorting descending orders the
  // taps by lsb (ties: any order, the tied vertices get weight 0).
  int k1 = (int)(xa | 0x00u), k2 = (int)(xb | 0x44u), k3 = (int)(xc | 0x22u), k4 = (int)(xd | 0x11u);
  int t;
  t = imax(k1, k2); k2 = imin(k1, k2); k1 = t;
  t = imax(k3, k4); k4 = imin(k3, k4); k3 = t;
  t = imax(k1, k3); k3 = imin(k1, k3); k1 = t;
  t = imax(k2, k4); k4 = imin(k2, k4); k2 = t;
  t = imax(k2, k3); k3 = imin(k2, k3); k2 = t;
  const uint32_t u2 = (uint32_t)(k1 | k2), u3 = u2 | (uint32_t)k3;
  // X: nibble 0 = corner of vertex 1, nibble 1 = corner of vertex 3 (nibble 2 is junk, nibble 3 is 0)
  s.selX = bitsel(u3, (uint32_t)k1, 0xFu);
  // Y: nibble 0 = corner of vertex 2, nibble 1 forced to 7 (vertex 4), nibble 3 = 0 (vertex 0)
  s.selY = u2 | 0x70u;
  const uint32_t a1 = prmt((uint32_t)k1, (uint32_t)k3, 0x2273u);  // [f1, f3, 0, 0]
  const uint32_t a2 = prmt((uint32_t)k2, (uint32_t)k4, 0x2273u);  // [f2, f4, 0, 0]
  const uint32_t a3 = prmt((uint32_t)k3, (uint32_t)k1, 0x7223u);  // [f3, 0, 0, f1]
  s.wX = a1 - a2;                // [f1-f2, f3-f4, 0, 0]          no borrows: the f are sorted
  s.wY = a2 + 0x10000000u - a3;  // [f2-f3, f4,    0, 16-f1]
  return s;
}

// Unsigned max / min of the two 16-bit halves (VIMNMX.U16x2 on sm_90+).
LERF_HD uint32_t vmax2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
#else
  const uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
  return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}
LERF_HD uint32_t vmin2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
#else
  const uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
  return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}

// Cell index of one lookup (first half of simplex_of).
LERF_HD uint32_t cell_of(uint32_t xa, uint32_t xb, uint32_t xc, uint32_t xd, const Hash& h) {
  const uint32_t acc = ((xa * 16u + xb) * 16u + xc) * 16u + xd;
  const uint32_t mix = xa * h.ha + (xb * h.hb + (xc * h.hc + acc));
  return prmt(bitsel(acc, mix, 0xF00u), 0u, 0x4421u);
}

// TWO lookups through ONE sorting network (r2).  The stage-1 kernel is bound by the ALU pipe and the ten
// min / max of a lookup's sort are its largest item.  The sort only needs the lsb (4 bits) and the tap's selector bits (one
// byte), so the keys of two lookups fit the two halves of a register -- half = lsb << 8 | selector bits -- and the 16x2
// min / max sorts both at once: per pair 4 PRMT to pack + 10 min / max instead of 20.  The selectors come out in the low
// byte of each sorted half exactly as in simplex_of (a PRMT reads only the low 16 bits of its selector operand; the
// second lookup's are shifted down); nibble 2 of a selector now holds an lsb, which is junk of weight 0 like before,
// and nibble 3 stays 0.  The weight bytes are picked from byte 1 (first lookup) / byte 3 (second) of the sorted halves; the zero
// bytes come from PRMT's sign replication of an lsb byte (<= 15).
LERF_HD void simplex_pair_of(const uint32_t x1[4], const uint32_t x2[4], const Hash& h, Simplex& s1, Simplex& s2) {
  s1.cell = cell_of(x1[0], x1[1], x1[2], x1[3], h);
  s2.cell = cell_of(x2[0], x2[1], x2[2], x2[3], h);
  // half = [selector bits, lsb]: bytes (x.b2 = 0, x.b3 = lsb) of each lookup's split_px word
  uint32_t k1 = prmt(x1[0], x2[0], 0x7632u);
  uint32_t k2 = prmt(x1[1], x2[1], 0x7632u) | 0x00440044u;
  uint32_t k3 = prmt(x1[2], x2[2], 0x7632u) | 0x00220022u;
  uint32_t k4 = prmt(x1[3], x2[3], 0x7632u) | 0x00110011u;
  uint32_t t;
  t = vmax2(k1, k2); k2 = vmin2(k1, k2); k1 = t;
  t = vmax2(k3, k4); k4 = vmin2(k3, k4); k3 = t;
  t = vmax2(k1, k3); k3 = vmin2(k1, k3); k1 = t;
  t = vmax2(k2, k4); k4 = vmin2(k2, k4); k2 = t;
  t = vmax2(k2, k3); k3 = vmin2(k2, k3); k2 = t;
  const uint32_t u2 = k1 | k2, u3 = u2 | k3;
  const uint32_t selX = bitsel(u3, k1, 0x000F000Fu), selY = u2 | 0x00700070u;
  s1.selX = selX;
  s1.selY = selY;
  s2.selX = selX >> 16;
  s2.selY = selY >> 16;
  {
    const uint32_t a1 = prmt(k1, k3, 0x9951u);  // [f1, f3, 0, 0]
    const uint32_t a2 = prmt(k2, k4, 0x9951u);  // [f2, f4, 0, 0]
    const uint32_t a3 = prmt(k3, k1, 0x5991u);  // [f3, 0, 0, f1]
    s1.wX = a1 - a2;
    s1.wY = a2 + 0x10000000u - a3;
  }
  {
    const uint32_t a1 = prmt(k1, k3, 0xBB73u);
    const uint32_t a2 = prmt(k2, k4, 0xBB73u);
    const uint32_t a3 = prmt(k3, k1, 0x7BB3u);
    s2.wX = a1 - a2;
    s2.wY = a2 + 0x10000000u - a3;
  }
}

// q = the cell's 16 bytes (x,y = X; z,w = Y).  Returns N = sum_k w_k * vertex_k, |N| <= 2048.
LERF_HD int blend(uint32_t qx, uint32_t qy, uint32_t qz, uint32_t qw, const Simplex& s) {
  const uint32_t r1 = prmt(qx, qy, s.selX);  // [P1, P3, junk, X[0]]
  const uint32_t r2 = prmt(qz, qw, s.selY);  // [P2, P4, junk, P0]
  return dp4a_ss(r1, s.wX, dp4a_ss(r2, s.wY, 0));
}

// Position of corner m inside the 16-byte block.
LERF_HD int corner_pos(int m) {
  const int odd = ((m >> 3) ^ (m >> 2) ^ (m >> 1) ^ m) & 1;
  return (odd ? 0 : 8) + (m & 7);
}

// Host-side repack of one row-major table T[17^4][oC] (int8) into cell blocks.
//   dst[swizzled_cell(cell) * cell_stride + slot_off + ch * 16 + corner_pos(m)] = T[perm-ed row][ch]
// perm[k] = which of the lookup's taps (0=a..3=d, in the order the kernel passes them) feeds table axis k, so a
// kernel that passes taps in ANCHOR order (A,B,C,D) reads the value T[tap perm[0], tap perm[1], ...].
inline void repack_cells(const int8_t* T, int oC, const int perm[4], const Hash& h, uint8_t* dst, size_t cell_stride,
                         size_t slot_off) {
  for (int cellidx = 0; cellidx < 65536; ++cellidx) {
    const size_t block = swizzled_cell((uint32_t)cellidx, h);
    const int msb[4] = {(cellidx >> 12) & 15, (cellidx >> 8) & 15, (cellidx >> 4) & 15, cellidx & 15};
    for (int m = 0; m < 16; ++m) {
      const int bump[4] = {(m >> 3) & 1, (m >> 2) & 1, (m >> 1) & 1, m & 1};
      int row = 0;
      for (int k = 0; k < 4; ++k) row = row * 17 + msb[perm[k]] + bump[perm[k]];
      for (int ch = 0; ch < oC; ++ch)
        dst[block * cell_stride + slot_off + (size_t)ch * 16 + corner_pos(m)] = (uint8_t)T[(size_t)row * oC + ch];
    }
  }
}

}  // namespace cell
}  // namespace lerf

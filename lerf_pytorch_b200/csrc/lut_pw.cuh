// "Paired-window" LUT format and lookup primitive of the LeRF hot path (sm_100a): one sort and ONE table block serve
// two of the twelve rotation-ensembled lookups.  Replaces the ensembling loops resample/eval_lut_sr.py:541-628 over
// FourSimplexInterpFaster (:24-470) of the reference.
//
// Why.  Rotation r at pixel z and rotation r+2 at the far end of the same window read the SAME four pixels in
// reverse order: the horizontal segment [x, x+3] is mode c / r0 for pixel x and mode c / r2 for pixel
// x+3, the vertical one is r1 / r3, the two diagonals are mode t; a 2x2 block is mode s for its four corners, one
// rotation each.  Same pixels => same 4-D cell and same LSB order => the same simplex, walked through the table
// with its axes permuted.  So the unit of work is a WINDOW ("anchor"), not a (pixel, rotation):
//   * one 5-compare-exchange sort per window instead of two (four for the 2x2 block);
//   * one block per window and table holding BOTH orientations' five vertices in walk order.
// The block is keyed by (cell, order of the taps' LSBs): nothing has to be selected out of it -- word (o, ch) is
// [P1, P2, P3, P4] for orientation o and output channel ch, and one DP4A with [f1-f2, f2-f3, f3-f4, f4] blends it;
// the P0 bytes share a word per orientation.  5 vertices x 3 channels x 2 orientations = 30 bytes: one 32-byte
// sector per TWO lookups where the max-tap format (lut_mt.cuh) needs one sector per lookup.  The stage-2 kernel
// runs at the speed of the L2 -> L1 sector stream (profiles/r1e_l2gather.csv), so halving the sectors is the lever.
//
// Layout of one table ("plane-major"): byte offset = (code6 << 16 | cell) * B, B = 32 (oC = 3) or 16 (oC = 1);
// code6 = t1 << 4 | t2 << 2 | t3 (the taps with the largest, 2nd and 3rd largest LSB; 24 of the 64 codes occur), so
// every LSB order is one 2 MiB (1 MiB) plane and only 24 planes of a table are ever touched -- 12 with the folded
// lookup of production (prepare_t<true> below: t1 < 2, so the product library builds planes 0..31 only).
//   oC = 3: word o*3 + ch = [P1..P4](o, ch);  word 6 + o = [P0(o,0), P0(o,1), P0(o,2), 0]
//   oC = 1: word o = [P1..P4](o);  word 2 = [P0(0), P0(1), 0, 0];  word 3 = 0
// Canonical tap order of a window: 2x2 block A B / C D -> (A, B, C, D); segments and diagonals from the anchor
// outwards.  Orientation o of family f reads canonical pixel kPi[f][o][k] as table axis k.
//
// This header is also compiled by g++ for the CPU emulation test (tests/csrc/cell_emul.cpp).
#pragma once
#include <stdint.h>

#include "lut_cell.cuh"

namespace lerf {
namespace pw {

using cell::dp4a_ss;
using cell::prmt;

// Window families.  Family f of stage 2 reads table f of (s r0, s r1, c r0, c r1, t r0, t r1); stage 1 has one
// table per mode (f >> 1).
enum { kFS0 = 0, kFS1 = 1, kFCH = 2, kFCV = 3, kFTD = 4, kFTA = 5, kFamilies = 6 };

// Orientation 0 / 1 of each family = rotation (and destination pixel, as an offset from the anchor):
//   S0: r0 -> A (0,0),  r2 -> D (+1,+1)        S1: r1 -> B (+1,0),  r3 -> C (0,+1)       [dest as (dx, dy)]
//   CH: r0 -> (0,0),    r2 -> (+3,0)           CV: r1 -> (0,0),     r3 -> (0,+3)
//   TD: r0 -> (0,0),    r2 -> (+3,+3)          TA: r1 -> (0,0),     r3 -> (-3,+3)
// kPi[f][o][k]: canonical pixel read by tap k (eval_lut_sr.py:30-81 composed with SURVEY.md A.3).
LERF_HD int pi_of(int f, int o, int k) {
  if (f == kFS1) {
    const int fwd[4] = {1, 3, 0, 2}, rev[4] = {2, 0, 3, 1};
    return o ? rev[k] : fwd[k];
  }
  return o ? 3 - k : k;
}

struct Lookup {
  uint32_t block;  // code6 << 16 | cell: block index inside a table
  uint32_t wd;     // [f1-f2, f2-f3, f3-f4, f4]: byte weights of P1..P4
  uint32_t w0;     // 16 - f1: weight of P0
  uint32_t flip;   // folded tables only: 1 = the block read is the REVERSED window's, its two orientations are swapped
};

// Taps in canonical order as cell::split_px words (lsb << 24 | msb << 8).
//
// FOLD (r2): every family's two orientations read the window forwards and backwards (pi_of(f, 1, k) = pi_of(f, 0, 3 - k)),
// so the block of window (p0, p1, p2, p3) with LSB order (t1, t2, t3, t4) is the block of the reversed window
// (p3, p2, p1, p0) with order (3 - t1, ...) with its two orientations swapped.  Exactly one of the two has t1 < 2 (t1 = 3 - t1
// has no solution), so only the 12 order planes with t1 in {0, 1} are ever read: half the table -- 72 instead of 144
// 2-MiB pages for the six stage-2 tables, inside the 128-page TLB reach (profiles/r2a_tlbgather.csv), and 144 instead of
// 288 MiB against the 126 MB L2 on inputs that scatter over the whole table.  Cost: the reversed cell index (three more
// IMADs), two selects, and a swap of the two results.
template <bool FOLD>
LERF_HD Lookup prepare_t(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t wd) {
  Lookup L;
  const uint32_t acc = ((wa * 16u + wb) * 16u + wc) * 16u + wd;  // bits 8..23 = cell (the lsb bytes land above)
  // key = lsb << 24 | msb << 8 | tap id: sorting descending orders the taps by lsb; ties (broken by msb, then id) may
  // fall either way, the tied vertices have weight 0
  int k1 = (int)wa, k2 = (int)(wb | 1u), k3 = (int)(wc | 2u), k4 = (int)(wd | 3u);
  int t;
  t = cell::imax(k1, k2); k2 = cell::imin(k1, k2); k1 = t;
  t = cell::imax(k3, k4); k4 = cell::imin(k3, k4); k3 = t;
  t = cell::imax(k1, k3); k3 = cell::imin(k1, k3); k1 = t;
  t = cell::imax(k2, k4); k4 = cell::imin(k2, k4); k2 = t;
  t = cell::imax(k2, k3); k3 = cell::imin(k2, k3); k2 = t;
  // bits 2..7 of the keys are zero, so the low six bits of k1*16 + k2*4 + k3 are (t1, t2, t3)
  const uint32_t raw = (uint32_t)k1 * 16u + (uint32_t)k2 * 4u + (uint32_t)k3;
  if (FOLD) {
    const uint32_t accr = ((wd * 16u + wc) * 16u + wb) * 16u + wa;  // the reversed window's cell
    L.flip = ((uint32_t)k1 >> 1) & 1u;                               // t1 >= 2
    const uint32_t code = (raw ^ (0u - L.flip)) & 0x3Fu;             // 3 - t = t ^ 3 on every tap id
    L.block = (code << 16) | (((L.flip ? accr : acc) >> 8) & 0xFFFFu);
  } else {
    L.flip = 0;
    L.block = ((raw & 0x3Fu) << 16) | ((acc >> 8) & 0xFFFFu);
  }
  const uint32_t f12 = prmt((uint32_t)k1, (uint32_t)k2, 0x0073u);  // [f1, f2, x, x]
  const uint32_t f34 = prmt((uint32_t)k3, (uint32_t)k4, 0x0073u);  // [f3, f4, x, x]
  const uint32_t F = prmt(f12, f34, 0x5410u);                      // [f1, f2, f3, f4]
  L.wd = F - (F >> 8);                                             // no borrows: the f are sorted
  L.w0 = 16u - ((uint32_t)k1 >> 24);
  return L;
}
LERF_HD Lookup prepare(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t wd) { return prepare_t<false>(wa, wb, wc, wd); }

// oC = 3: q = the block's 8 words.  n[o][ch] += the lookup of orientation o.
LERF_HD void blend3(const uint32_t q[8], const Lookup& L, int n[2][3]) {
  const uint32_t w0b = L.w0 << 8, w0c = L.w0 << 16;
  n[0][0] = dp4a_ss(q[0], L.wd, dp4a_ss(q[6], L.w0, n[0][0]));
  n[0][1] = dp4a_ss(q[1], L.wd, dp4a_ss(q[6], w0b, n[0][1]));
  n[0][2] = dp4a_ss(q[2], L.wd, dp4a_ss(q[6], w0c, n[0][2]));
  n[1][0] = dp4a_ss(q[3], L.wd, dp4a_ss(q[7], L.w0, n[1][0]));
  n[1][1] = dp4a_ss(q[4], L.wd, dp4a_ss(q[7], w0b, n[1][1]));
  n[1][2] = dp4a_ss(q[5], L.wd, dp4a_ss(q[7], w0c, n[1][2]));
}

// oC = 1: q = words 0..2 of the block.
LERF_HD void blend1(const uint32_t q[3], const Lookup& L, int n[2]) {
  n[0] = dp4a_ss(q[0], L.wd, dp4a_ss(q[2], L.w0, n[0]));
  n[1] = dp4a_ss(q[1], L.wd, dp4a_ss(q[2], L.w0 << 8, n[1]));
}

#if defined(__CUDACC__)
#define LERF_HDC __host__ __device__ constexpr
#else
#define LERF_HDC constexpr
#endif
LERF_HDC int block_bytes(int oC) { return oC == 3 ? 32 : 16; }
// `planes` order planes of 65536 blocks: 64 address every code6; folded lookups (prepare_t<true>) only read codes < 32.
LERF_HDC size_t table_bytes(int oC, int planes = 64) { return (size_t)planes * 65536 * (size_t)(oC == 3 ? 32 : 16); }

// Fills block (code6, cell) of family f from the row-major table T[17^4][entry_stride >= oC]; returns false (block untouched) for
// the 40 codes that are not an order of four distinct taps.  Used by the repack kernel and by the CPU emulation.
LERF_HD bool fill_block(const int8_t* T, int oC, int entry_stride, int f, uint32_t cellidx, uint32_t code, uint8_t* blk) {
  const int t1 = (code >> 4) & 3, t2 = (code >> 2) & 3, t3 = code & 3;
  if (t1 == t2 || t1 == t3 || t2 == t3) return false;
  const int order[4] = {t1, t2, t3, 6 - t1 - t2 - t3};
  const int msb[4] = {(int)((cellidx >> 12) & 15u), (int)((cellidx >> 8) & 15u), (int)((cellidx >> 4) & 15u),
                      (int)(cellidx & 15u)};
  const int B = block_bytes(oC);
  for (int i = 0; i < B; ++i) blk[i] = 0;
  for (int o = 0; o < 2; ++o) {
    int bump[4] = {0, 0, 0, 0};  // per canonical pixel
    for (int j = 0; j <= 4; ++j) {
      if (j) bump[order[j - 1]] = 1;
      int row = 0;
      for (int k = 0; k < 4; ++k) row = row * 17 + msb[pi_of(f, o, k)] + bump[pi_of(f, o, k)];
      for (int ch = 0; ch < oC; ++ch) {
        const uint8_t v = (uint8_t)T[(size_t)row * entry_stride + ch];
        if (oC == 3) blk[j ? (o * 3 + ch) * 4 + (j - 1) : 24 + o * 4 + ch] = v;
        else blk[j ? o * 4 + (j - 1) : 8 + o] = v;
      }
    }
  }
  return true;
}

// Anchor -> destination offsets (dx, dy) of orientation 1 (orientation 0 always lands on the anchor, except S1).
//   S0: (+1,+1)   S1: o0 (+1,0), o1 (0,+1)   CH: (+3,0)   CV: (0,+3)   TD: (+3,+3)   TA: (-3,+3)
// Tap k of the canonical window, as an offset from the anchor:
LERF_HD void window_tap(int f, int k, int& dx, int& dy) {
  switch (f) {
    case kFS0: case kFS1: dx = k & 1; dy = k >> 1; break;
    case kFCH: dx = k; dy = 0; break;
    case kFCV: dx = 0; dy = k; break;
    case kFTD: dx = k; dy = k; break;
    default: dx = -k; dy = k; break;
  }
}

}  // namespace pw
}  // namespace lerf

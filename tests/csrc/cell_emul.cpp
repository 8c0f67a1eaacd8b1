// CPU emulation of the cell-packed LUT lookup (lerf_pytorch_b200/csrc/lut_cell.cuh) -- TEST INFRASTRUCTURE ONLY.
// It compiles the product's own header with g++ (every device intrinsic has a host twin there) and walks an
// image exactly like lut_stage_cell_kernel does, so the bit tricks (key layout, PRMT selectors, byte weights, table
// repack) are checked against the oracle on the CPU before any GPU time is spent.  Nothing in the product loads it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../lerf_pytorch_b200/csrc/lut_cell.cuh"
#include "../../lerf_pytorch_b200/csrc/lut_mt.cuh"

using namespace lerf::cell;

static void tap_offset(int mode, int r, int k, int& dy, int& dx) {  // mode 0='s',1='c',2='t' (eval_lut_sr.py:30-81)
  const int di = mode == 0 ? (k >> 1) : (mode == 1 ? 0 : k);
  const int dj = mode == 0 ? (k & 1) : k;
  dy = r == 0 ? di : (r == 1 ? dj : (r == 2 ? -di : -dj));
  dx = r == 0 ? dj : (r == 1 ? -di : (r == 2 ? -dj : di));
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int rhe_div(int num, int den) {
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;
  return q;
}

// paired != 0 (oC = 1 only): two lookups per sorting network (simplex_pair_of), paired like lut_stage_cell_body does.
extern "C" int emul_stage_cell(int stage, const int8_t* const* tables, int oC, const uint8_t* img, int P, int H, int W,
                               int ha, int hb, int hc, int paired, uint8_t* out) {
  const Hash h{(uint32_t)ha, (uint32_t)hb, (uint32_t)hc};
  const int ntab = stage == 1 ? 3 : 6;
  const size_t stride = oC == 3 ? 48 : 16;
  std::vector<std::vector<uint8_t>> packed(ntab);
  const int ident[4] = {0, 1, 2, 3};
  for (int i = 0; i < ntab; ++i) {
    packed[i].assign((size_t)65536 * stride, 0);
    repack_cells(tables[i], oC, ident, h, packed[i].data(), stride, 0);
  }
  for (int p = 0; p < P; ++p)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int n[3] = {0, 0, 0};
        for (int mode = 0; mode < 3; ++mode)
          for (int r = 0; r < 4; ++r) {
            uint32_t xw[2][4];
            for (int j = 0; j < 2; ++j)
              for (int k = 0; k < 4; ++k) {
                int dy, dx;
                tap_offset(mode, (r & ~1) + j, k, dy, dx);
                xw[j][k] = split_px(img[((size_t)p * H + clampi(y + dy, 0, H - 1)) * W + clampi(x + dx, 0, W - 1)]);
              }
            Simplex s;
            if (paired && oC == 1) {
              Simplex sp[2];
              simplex_pair_of(xw[0], xw[1], h, sp[0], sp[1]);
              s = sp[r & 1];
            } else {
              s = simplex_of(xw[r & 1][0], xw[r & 1][1], xw[r & 1][2], xw[r & 1][3], h);
            }
            const uint8_t* tab = packed[stage == 1 ? mode : 2 * mode + (r & 1)].data() + (size_t)s.cell * stride;
            for (int ch = 0; ch < oC; ++ch) {
              uint32_t q[4];
              memcpy(q, tab + 16 * ch, 16);
              n[ch] += blend(q[0], q[1], q[2], q[3], s);
            }
          }
        for (int ch = 0; ch < oC; ++ch) {
          int v;
          if (stage == 1) {
            v = n[ch] <= 0 ? 0 : (rhe_div(n[ch], 48) > 255 ? 255 : rhe_div(n[ch], 48));
          } else {
            const int t = n[ch] + 127 * 192;
            v = t <= 0 ? 0 : (rhe_div(t, 192) > 255 ? 255 : rhe_div(t, 192));
          }
          out[(((size_t)p * oC + ch) * H + y) * W + x] = (uint8_t)v;
        }
      }
  return 0;
}

// Stage 2 (oC = 3) on max-tap blocks (lut_mt.cuh), walked like lut_stage2_mt_kernel does.
extern "C" int emul_stage2_maxtap(const int8_t* const* tables, const uint8_t* img, int P, int H, int W, uint8_t* out) {
  std::vector<std::vector<uint8_t>> packed(6);
  for (int i = 0; i < 6; ++i) {
    packed[i].assign(lerf::mt::kTableBytes, 0);
    lerf::mt::repack_maxtap(tables[i], packed[i].data());
  }
  for (int p = 0; p < P; ++p)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int n[3] = {0, 0, 0};
        for (int mode = 0; mode < 3; ++mode)
          for (int r = 0; r < 4; ++r) {
            uint32_t k[4], m[4];
            for (int t = 0; t < 4; ++t) {
              int dy, dx;
              tap_offset(mode, r, t, dy, dx);
              lerf::mt::split_px2(img[((size_t)p * H + clampi(y + dy, 0, H - 1)) * W + clampi(x + dx, 0, W - 1)], k[t], m[t]);
            }
            const lerf::mt::Lookup L = lerf::mt::prepare(k[0], m[0], k[1], m[1], k[2], m[2], k[3], m[3]);
            uint32_t q[8];
            memcpy(q, packed[2 * mode + (r & 1)].data() + (size_t)L.block * lerf::mt::kBlockBytes, 32);
            lerf::mt::blend3(q, L, n[0], n[1], n[2]);
          }
        for (int ch = 0; ch < 3; ++ch) {
          const int t = n[ch] + 127 * 192;
          const int v = t <= 0 ? 0 : (rhe_div(t, 192) > 255 ? 255 : rhe_div(t, 192));
          out[(((size_t)p * 3 + ch) * H + y) * W + x] = (uint8_t)v;
        }
      }
  return 0;
}

// Same stage on the single-word tap form (lut_mt.cuh prepare1, production since r1e).
extern "C" int emul_stage2_maxtap1(const int8_t* const* tables, const uint8_t* img, int P, int H, int W, uint8_t* out) {
  std::vector<std::vector<uint8_t>> packed(6);
  for (int i = 0; i < 6; ++i) {
    packed[i].assign(lerf::mt::kTableBytes, 0);
    lerf::mt::repack_maxtap(tables[i], packed[i].data());
  }
  for (int p = 0; p < P; ++p)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int n[3] = {0, 0, 0};
        for (int mode = 0; mode < 3; ++mode)
          for (int r = 0; r < 4; ++r) {
            uint32_t w[4];
            for (int t = 0; t < 4; ++t) {
              int dy, dx;
              tap_offset(mode, r, t, dy, dx);
              w[t] = lerf::cell::split_px(img[((size_t)p * H + clampi(y + dy, 0, H - 1)) * W + clampi(x + dx, 0, W - 1)]);
            }
            const lerf::mt::Lookup L = lerf::mt::prepare1(w[0], w[1], w[2], w[3]);  // single-word form: L.block is a byte offset
            uint32_t q[8];
            memcpy(q, packed[2 * mode + (r & 1)].data() + (size_t)L.block, 32);
            lerf::mt::blend3(q, L, n[0], n[1], n[2]);
          }
        for (int ch = 0; ch < 3; ++ch) {
          const int t = n[ch] + 127 * 192;
          const int v = t <= 0 ? 0 : (rhe_div(t, 192) > 255 ? 255 : rhe_div(t, 192));
          out[(((size_t)p * 3 + ch) * H + y) * W + x] = (uint8_t)v;
        }
      }
  return 0;
}

// Paired-window format (lut_pw.cuh): every window is sorted once and one block per table serves both orientations;
// walked window by window over the whole (edge-replicated) image like lut_stage_pw_kernel does per tile.
#include "../../lerf_pytorch_b200/csrc/lut_pw.cuh"

// fold != 0: folded tables (prepare_t<true>: only order planes with t1 < 2 are read, results swapped on a flipped lookup).
extern "C" int emul_stage_pw(int stage, const int8_t* const* tables, int oC, const uint8_t* img, int P, int H, int W,
                             int fold, uint8_t* out) {
  namespace pw = lerf::pw;
  static const int dest[6][2][2] = {{{0, 0}, {1, 1}}, {{1, 0}, {0, 1}}, {{0, 0}, {3, 0}},
                                    {{0, 0}, {0, 3}}, {{0, 0}, {3, 3}}, {{0, 0}, {-3, 3}}};  // [family][orientation](dx, dy)
  std::vector<int> acc((size_t)P * oC * H * W, 0);
  for (int p = 0; p < P; ++p)
    for (int f = 0; f < 6; ++f) {
      const int8_t* T = tables[stage == 1 ? (f >> 1) : f];
      for (int ay = -3; ay < H; ++ay)
        for (int ax = -3; ax < W + 3; ++ax) {
          uint32_t w[4];
          for (int k = 0; k < 4; ++k) {
            int dx, dy;
            pw::window_tap(f, k, dx, dy);
            w[k] = split_px(img[((size_t)p * H + clampi(ay + dy, 0, H - 1)) * W + clampi(ax + dx, 0, W - 1)]);
          }
          const pw::Lookup L = fold ? pw::prepare_t<true>(w[0], w[1], w[2], w[3]) : pw::prepare(w[0], w[1], w[2], w[3]);
          if (fold && ((L.block >> 20) & 3u) >= 2u) return 2;  // a folded lookup must stay in the planes with t1 < 2
          uint8_t blk[32];
          if (!pw::fill_block(T, oC, oC, f, L.block & 0xFFFFu, L.block >> 16, blk)) return 1;  // an impossible order code
          uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          memcpy(q, blk, pw::block_bytes(oC));
          int n[2][3] = {{0, 0, 0}, {0, 0, 0}};
          if (oC == 3) {
            pw::blend3(q, L, n);
          } else {
            int m[2] = {0, 0};
            pw::blend1(q, L, m);
            n[0][0] = m[0];
            n[1][0] = m[1];
          }
          for (int o = 0; o < 2; ++o) {
            const int x = ax + dest[f][o][0], y = ay + dest[f][o][1];
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
            for (int ch = 0; ch < oC; ++ch) acc[(((size_t)p * oC + ch) * H + y) * W + x] += n[o ^ (int)L.flip][ch];
          }
        }
    }
  for (size_t i = 0; i < acc.size(); ++i) {
    int v;
    if (stage == 1) {
      v = acc[i] <= 0 ? 0 : (rhe_div(acc[i], 48) > 255 ? 255 : rhe_div(acc[i], 48));
    } else {
      const int t = acc[i] + 127 * 192;
      v = t <= 0 ? 0 : (rhe_div(t, 192) > 255 ? 255 : rhe_div(t, 192));
    }
    out[i] = (uint8_t)v;
  }
  return 0;
}

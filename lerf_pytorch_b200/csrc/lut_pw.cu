// Rotation-ensembled LUT stages on paired-window tables (lut_pw.cuh): table build (repack kernel) and the stage kernel.
// Reference being replaced (ddlee-cn/LeRF-PyTorch): stage-1 / stage-2 loops resample/eval_lut_sr.py:541-628
// (= eval_lut_warp.py:104-191) over FourSimplexInterpFaster (:24-470).
//
// Kernel: one CTA = one 32 x 32 pixel tile of one plane, 256 threads, a thread owns the pixels (tx, tq + 8j).  The
// unit of work is a WINDOW anchored at a pixel: a 2x2 block, a horizontal / vertical 4-segment, a diagonal and an
// anti-diagonal.  Each window is sorted once and fetches one block per table that holds both orientations, i.e. two
// of the twelve lookups of the ensemble (four for the 2x2 block, from two tables).  The lookup that belongs to
// the anchor pixel accumulates in the owner's registers; the one that belongs to the pixel at the far end of the
// window goes through seven shared-memory exchange arrays, one per (family, far end), each written exactly once per
// tile pixel -- plain stores, one barrier before the final sum.
// Windows whose anchor lies in the 3-pixel halo around the tile only feed the far end ("halo anchors": they ride
// along as a fifth element of each thread's batch of loads).  Pixels outside the image are edge-replicated when the
// tile is staged, which is exactly what the reference's rotate-then-pad does, so a window that hangs
// over the border still pairs up.
// Measured (B200, r2a, 8 frames 2040x1356): stage 2 (LeRF-G) 271 us per frame on natural-like input (max-tap kernel:
// 345), 337 on uniform input (372); 6.7 sectors per sample cross the L2 -> L1 link instead of 12.
#include <vector>

#include "common.cuh"
#include "lut_pw.cuh"

namespace lerf {
namespace pwk {

constexpr int kT = 32, kHalo = 3, kPitch = 40, kRows = kT + 2 * kHalo;

struct Tables {
  const uint8_t* t[6];
};

// ---------------------------------------------------------------------------------------------------------------
// repack: row-major T[17^4][oC] (device) -> paired-window table of family f
// ---------------------------------------------------------------------------------------------------------------
__global__ void pw_repack_kernel(const int8_t* __restrict__ T, int oC, int entry_stride, int f, uint8_t* __restrict__ dst) {
  const uint32_t cellidx = blockIdx.x * blockDim.x + threadIdx.x;  // 65536 cells
  const uint32_t code = blockIdx.y;                                // 64 codes
  __align__(16) uint8_t blk[32];
  if (!pw::fill_block(T, oC, entry_stride, f, cellidx, code, blk)) return;
  const size_t B = pw::block_bytes(oC);
  uint4* o = reinterpret_cast<uint4*>(dst + (((size_t)code << 16) | cellidx) * B);
  const uint4* s = reinterpret_cast<const uint4*>(blk);
  o[0] = s[0];
  if (oC == 3) o[1] = s[1];
}

#ifdef LERF_EXPERIMENTS
// cell-pair table of family f (oC = 1): block `cell` = [orientation 0: 16 corners][orientation 1: 16 corners], corner m of
// the CANONICAL window at byte cell::corner_pos(m) of its half
__global__ void cp_repack_kernel(const int8_t* __restrict__ T, int f, uint8_t* __restrict__ dst) {
  const uint32_t cellidx = blockIdx.x * blockDim.x + threadIdx.x;
  const int msb[4] = {(int)((cellidx >> 12) & 15u), (int)((cellidx >> 8) & 15u), (int)((cellidx >> 4) & 15u), (int)(cellidx & 15u)};
  __align__(16) uint8_t blk[32];
  for (int o = 0; o < 2; ++o)
    for (int m = 0; m < 16; ++m) {
      const int bump[4] = {(m >> 3) & 1, (m >> 2) & 1, (m >> 1) & 1, m & 1};
      int row = 0;
      for (int k = 0; k < 4; ++k) row = row * 17 + msb[pw::pi_of(f, o, k)] + bump[pw::pi_of(f, o, k)];
      blk[16 * o + cell::corner_pos(m)] = (uint8_t)T[row];
    }
  uint4* d = reinterpret_cast<uint4*>(dst + (size_t)cellidx * 32);
  d[0] = reinterpret_cast<const uint4*>(blk)[0];
  d[1] = reinterpret_cast<const uint4*>(blk)[1];
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// packed accumulators: oC = 3 keeps (n0 + n1 * 65536, n2) -- every partial sum of the twelve |N| <= 2048 fits 16 bits
// ---------------------------------------------------------------------------------------------------------------
template <int OC>
struct Pk;
template <>
struct Pk<3> {
  int a, b;
  __device__ __forceinline__ void zero() { a = b = 0; }
  __device__ __forceinline__ void add(const Pk& o) { a += o.a; b += o.b; }
  __device__ __forceinline__ void set(const int* n) { a = n[1] * 65536 + n[0]; b = n[2]; }
  __device__ __forceinline__ void get(int* n) const {
    n[0] = (int)(short)(a & 0xFFFF);
    n[1] = (a - n[0]) >> 16;
    n[2] = b;
  }
};
template <>
struct Pk<1> {
  int a;
  __device__ __forceinline__ void zero() { a = 0; }
  __device__ __forceinline__ void add(const Pk& o) { a += o.a; }
  __device__ __forceinline__ void set(const int* n) { a = n[0]; }
  __device__ __forceinline__ void get(int* n) const { n[0] = a; }
};

// Exchange-array entry of the window kernel.  Pk<3> is 8 bytes; the compact form splits it into a 4-byte array (n0, n1 as
// a packed int16 pair -- exactly Pk::a) and a 2-byte array (n2: one lookup's |N| <= 2048), 6 bytes per entry, so four
// instead of three 32 x 32 blocks fit an SM's shared memory.
template <int OC, bool X6>
struct Xch {
  Pk<OC>* p;
  __device__ __forceinline__ void init(unsigned char* base, int entries) { p = reinterpret_cast<Pk<OC>*>(base); (void)entries; }
  __device__ __forceinline__ void put(int i, const Pk<OC>& v) const { p[i] = v; }
  __device__ __forceinline__ Pk<OC> get(int i) const { return p[i]; }
  static constexpr int kBytes = sizeof(Pk<OC>);
};
template <>
struct Xch<3, true> {
  int* a;
  short* b;
  __device__ __forceinline__ void init(unsigned char* base, int entries) {
    a = reinterpret_cast<int*>(base);
    b = reinterpret_cast<short*>(base + (size_t)entries * 4);
  }
  __device__ __forceinline__ void put(int i, const Pk<3>& v) const { a[i] = v.a; b[i] = (short)v.b; }
  __device__ __forceinline__ Pk<3> get(int i) const {
    Pk<3> v;
    v.a = a[i];
    v.b = b[i];
    return v;
  }
  static constexpr int kBytes = 6;
};

// LD: 0 = ld.global.nc (allocates in L1), 1 = nc + L1::no_allocate
template <int LD>
__device__ __forceinline__ void ld32(const uint8_t* p, uint32_t* q) {
  if (LD == 1)
    asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
        : "l"(p));
  else
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
        : "l"(p));
}
template <int LD>
__device__ __forceinline__ void ld16(const uint8_t* p, uint32_t* q) {
  if (LD == 1)
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "l"(p));
  else
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "l"(p));
}

// Table formats of the window kernel.
//   FmtPW<OC>: paired-window blocks keyed by (cell, LSB order) (lut_pw.cuh) -- no reuse between neighbours, fetched
//              around L1.  Production for stage 2 of LeRF-G: the kernel is bound by sectors through the L2 -> L1 link.
//   FmtCP:     "cell pair", oC = 1: the 16 corners of the cell for both orientations (2 x 16 bytes, lut_cell.cuh corner
//              order) keyed by the cell alone -- 2 MiB per table, so neighbouring windows hit L1 like the per-pixel
//              cell kernel does, but one sort and one 256-bit load serve two lookups.
template <int OC, bool FOLD = false>
struct FmtPW {
  using Lookup = pw::Lookup;
  static constexpr int nq = OC == 3 ? 8 : 4;
  static constexpr int kDefaultLD = 1;
  __device__ static __forceinline__ Lookup prepare(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return pw::prepare_t<FOLD>(a, b, c, d); }
  template <int LD>
  __device__ static __forceinline__ void fetch(const uint8_t* __restrict__ tab, const Lookup& L, uint32_t* q) {
    if (OC == 3) ld32<LD>(tab + (size_t)L.block * 32, q);
    else ld16<LD>(tab + (size_t)L.block * 16, q);
  }
  __device__ static __forceinline__ void blend(const uint32_t* q, const Lookup& L, Pk<OC>& o0, Pk<OC>& o1) {
    if (OC == 3) {
      int n[2][3] = {{0, 0, 0}, {0, 0, 0}};
      pw::blend3(q, L, n);
      o0.set(n[0]);
      o1.set(n[1]);
    } else {
      int n[2] = {0, 0};
      pw::blend1(q, L, n);
      o0.set(&n[0]);
      o1.set(&n[1]);
    }
    if (FOLD && L.flip) {  // the reversed window's block: its orientations are ours, swapped
      const Pk<OC> t = o0;
      o0 = o1;
      o1 = t;
    }
  }
};

#ifdef LERF_EXPERIMENTS
struct FmtCP {
  using Lookup = cell::Simplex;
  static constexpr int nq = 8;
  static constexpr int kDefaultLD = 0;
  __device__ static __forceinline__ Lookup prepare(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return cell::simplex_of(a, b, c, d, cell::Hash{0u, 0u, 0u});
  }
  template <int LD>
  __device__ static __forceinline__ void fetch(const uint8_t* __restrict__ tab, const Lookup& L, uint32_t* q) {
    ld32<LD>(tab + (size_t)L.cell * 32, q);
  }
  __device__ static __forceinline__ void blend(const uint32_t* q, const Lookup& L, Pk<1>& o0, Pk<1>& o1) {
    o0.a = cell::blend(q[0], q[1], q[2], q[3], L);
    o1.a = cell::blend(q[4], q[5], q[6], q[7], L);
  }
};
#endif

template <int TY>
__device__ __forceinline__ bool in_tile(int x, int y) { return (unsigned)x < (unsigned)kT && (unsigned)y < (unsigned)TY; }

__device__ __forceinline__ int rhe_div(int num, int den) {  // round_half_even(num / den), num > 0, den even
  const int t = num + den / 2;
  int q = t / den;
  if (t - q * den == 0 && (q & 1)) --q;
  return q;
}

// Window group G (0 = 2x2 block, 1 = CH, 2 = CV, 3 = TD, 4 = TA): tap offsets in tile words, number of halo anchors and
// their coordinates, destination of the far-end lookup.
template <int G, int TY>
struct Grp {
  static constexpr int P = kPitch;
  static constexpr int o1 = G == 0 ? 1 : (G == 1 ? 1 : (G == 2 ? P : (G == 3 ? P + 1 : P - 1)));
  static constexpr int o2 = G == 0 ? P : 2 * o1, o3 = G == 0 ? P + 1 : 3 * o1;
  static constexpr int nhalo = G == 0 ? TY + 33 : (G == 1 ? 3 * TY : (G == 2 ? 96 : 105 + 3 * TY));
  static constexpr int ddx = G == 0 ? 1 : (G == 1 ? 3 : (G == 2 ? 0 : (G == 3 ? 3 : -3)));
  static constexpr int ddy = G == 0 ? 1 : (G == 1 ? 0 : 3);
  __device__ static __forceinline__ void halo(int i, int& ax, int& ay) {
    if (G == 0) { ax = i <= TY ? -1 : i - (TY + 1); ay = i <= TY ? i - 1 : -1; }      // column -1 (rows -1..TY-1), then row -1
    if (G == 1) { ax = -3 + i % 3; ay = i / 3; }                                     // columns -3..-1
    if (G == 2) { ax = i & 31; ay = -3 + (i >> 5); }                                 // rows -3..-1
    if (G == 3) { ax = i < 105 ? -3 + i % 35 : -3 + (i - 105) % 3; ay = i < 105 ? -3 + i / 35 : (i - 105) / 3; }
    if (G == 4) { ax = i < 105 ? i % 35 : 32 + (i - 105) % 3; ay = i < 105 ? -3 + i / 35 : (i - 105) / 3; }
  }
};

// One pass of group G over the tile: the thread's four anchors (tx, tq + 8j) plus at most one halo anchor.  All five
// table loads are issued before the first one is consumed (the kernel lives on load latency: ncu r2a, long_scoreboard).
// Exchange arrays (each written exactly once per tile pixel, so plain stores and a single barrier before the final
// sum):  X[0] S0.o1 (+1,+1)   X[1] S1.o0 (+1,0)   X[2] S1.o1 (0,+1)   X[3] CH (+3,0)   X[4] CV (0,+3)   X[5] TD (+3,+3)   X[6] TA (-3,+3)
template <typename Fmt, int OC, int LD, int G, int NJ, typename XT>
__device__ __forceinline__ void group_pass(const Tables& t, const uint32_t* __restrict__ tile, const XT& X, int tx,
                                           int tq, int tid, Pk<OC> own[NJ]) {
  constexpr int TY = 8 * NJ;
  using Gr = Grp<G, TY>;
  constexpr int nq = Fmt::nq;
  constexpr int kPx = kT * TY;
  typename Fmt::Lookup L[NJ + 1];
  uint32_t q[NJ + 1][nq];
  int hx = 0, hy = 0;
  const bool h = tid < Gr::nhalo;
  if (h) Gr::halo(tid, hx, hy);
#pragma unroll
  for (int j = 0; j <= NJ; ++j) {
    if (j == NJ && !h) break;
    const int ax = j < NJ ? tx : hx, ay = j < NJ ? tq + 8 * j : hy;
    const uint32_t* c = tile + (ay + kHalo) * kPitch + ax + kHalo;
    L[j] = Fmt::prepare(c[0], c[Gr::o1], c[Gr::o2], c[Gr::o3]);
    Fmt::template fetch<LD>(t.t[G == 0 ? 0 : G + 1], L[j], q[j]);
  }
#pragma unroll
  for (int j = 0; j <= NJ; ++j) {
    if (j == NJ && !h) break;
    const int ax = j < NJ ? tx : hx, ay = j < NJ ? tq + 8 * j : hy;
    Pk<OC> f, r;
    Fmt::blend(q[j], L[j], f, r);
    if (j < NJ) own[j].add(f);
    const int dx = ax + Gr::ddx, dy = ay + Gr::ddy;
    if (in_tile<TY>(dx, dy)) X.put((G == 0 ? 0 : G + 2) * kPx + dy * kT + dx, r);
  }
  if (G == 0) {  // the second table of the 2x2 block: rotation 1 lands on B (+1,0), rotation 3 on C (0,+1)
#pragma unroll
    for (int j = 0; j <= NJ; ++j) {
      if (j == NJ && !h) break;
      Fmt::template fetch<LD>(t.t[1], L[j], q[j]);
    }
#pragma unroll
    for (int j = 0; j <= NJ; ++j) {
      if (j == NJ && !h) break;
      const int ax = j < NJ ? tx : hx, ay = j < NJ ? tq + 8 * j : hy;
      Pk<OC> b0, b1;
      Fmt::blend(q[j], L[j], b0, b1);
      if (in_tile<TY>(ax + 1, ay)) X.put(1 * kPx + ay * kT + ax + 1, b0);
      if (in_tile<TY>(ax, ay + 1)) X.put(2 * kPx + (ay + 1) * kT + ax, b1);
    }
  }
}

// Software-pipelined flavour of group_pass (r2): the loads of stage k+1 are ISSUED before the blocks of stage k are
// consumed, so a warp always has a batch of table loads in flight while it blends the previous one.  Six stages: the 2x2
// block on table 0, the same windows on table 1 (lookups reused), CH, CV, TD, TA.  Two register buffers of NJ + 1 blocks.
template <typename Fmt, int NJ>
struct StageBuf {
  typename Fmt::Lookup L[NJ + 1];
  uint32_t q[NJ + 1][Fmt::nq];
  int hx, hy;
  bool h;
};

// ST: 0 = S on table 0, 1 = S on table 1 (reuses `prev`'s lookups), 2..5 = groups 1..4
template <typename Fmt, int LD, int ST, int NJ>
__device__ __forceinline__ void stage_issue(const Tables& t, const uint32_t* __restrict__ tile, int tx, int tq, int tid,
                                            StageBuf<Fmt, NJ>& b, const StageBuf<Fmt, NJ>* prev) {
  constexpr int G = ST <= 1 ? 0 : ST - 1;
  constexpr int TY = 8 * NJ;
  using Gr = Grp<G, TY>;
  b.h = tid < Gr::nhalo;
  b.hx = b.hy = 0;
  if (b.h) Gr::halo(tid, b.hx, b.hy);
#pragma unroll
  for (int j = 0; j <= NJ; ++j) {
    if (j == NJ && !b.h) break;
    if (ST == 1) {
      b.L[j] = prev->L[j];
    } else {
      const int ax = j < NJ ? tx : b.hx, ay = j < NJ ? tq + 8 * j : b.hy;
      const uint32_t* c = tile + (ay + kHalo) * kPitch + ax + kHalo;
      b.L[j] = Fmt::prepare(c[0], c[Gr::o1], c[Gr::o2], c[Gr::o3]);
    }
    Fmt::template fetch<LD>(t.t[ST], b.L[j], b.q[j]);
  }
}

template <typename Fmt, int OC, int ST, int NJ, typename XT>
__device__ __forceinline__ void stage_consume(const StageBuf<Fmt, NJ>& b, const XT& X, int tx, int tq, Pk<OC> own[NJ]) {
  constexpr int G = ST <= 1 ? 0 : ST - 1;
  constexpr int TY = 8 * NJ;
  using Gr = Grp<G, TY>;
  constexpr int kPx = kT * TY;
#pragma unroll
  for (int j = 0; j <= NJ; ++j) {
    if (j == NJ && !b.h) break;
    const int ax = j < NJ ? tx : b.hx, ay = j < NJ ? tq + 8 * j : b.hy;
    Pk<OC> f, r;
    Fmt::blend(b.q[j], b.L[j], f, r);
    if (ST == 1) {  // rotation 1 lands on B (+1,0), rotation 3 on C (0,+1)
      if (in_tile<TY>(ax + 1, ay)) X.put(1 * kPx + ay * kT + ax + 1, f);
      if (in_tile<TY>(ax, ay + 1)) X.put(2 * kPx + (ay + 1) * kT + ax, r);
    } else {
      if (j < NJ) own[j].add(f);
      const int dx = ax + Gr::ddx, dy = ay + Gr::ddy;
      if (in_tile<TY>(dx, dy)) X.put((G == 0 ? 0 : G + 2) * kPx + dy * kT + dx, r);
    }
  }
}

template <typename Fmt, int OC, int LD, int NJ, typename XT>
__device__ __forceinline__ void pipelined_passes(const Tables& t, const uint32_t* __restrict__ tile, const XT& X, int tx,
                                                 int tq, int tid, Pk<OC> own[NJ]) {
  StageBuf<Fmt, NJ> A, B;
  stage_issue<Fmt, LD, 0, NJ>(t, tile, tx, tq, tid, A, nullptr);
  stage_issue<Fmt, LD, 1, NJ>(t, tile, tx, tq, tid, B, &A);
  stage_consume<Fmt, OC, 0, NJ>(A, X, tx, tq, own);
  stage_issue<Fmt, LD, 2, NJ>(t, tile, tx, tq, tid, A, nullptr);
  stage_consume<Fmt, OC, 1, NJ>(B, X, tx, tq, own);
  stage_issue<Fmt, LD, 3, NJ>(t, tile, tx, tq, tid, B, nullptr);
  stage_consume<Fmt, OC, 2, NJ>(A, X, tx, tq, own);
  stage_issue<Fmt, LD, 4, NJ>(t, tile, tx, tq, tid, A, nullptr);
  stage_consume<Fmt, OC, 3, NJ>(B, X, tx, tq, own);
  stage_issue<Fmt, LD, 5, NJ>(t, tile, tx, tq, tid, B, nullptr);
  stage_consume<Fmt, OC, 4, NJ>(A, X, tx, tq, own);
  stage_consume<Fmt, OC, 5, NJ>(B, X, tx, tq, own);
}

template <typename Fmt, int STAGE, int OC, int LD, int MINB, int NJ = 4, bool PIPE = false, bool X6 = false>
__global__ void __launch_bounds__(256, MINB)
    lut_stage_pw_kernel(Tables t, const uint8_t* __restrict__ in, InAddr ia, int H, int W, int y0, int y1,
                        uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TY = 8 * NJ, kTileRows = TY + 2 * kHalo;
  uint32_t* tile = reinterpret_cast<uint32_t*>(smem_raw);
  Xch<OC, X6> X;
  X.init(smem_raw + kTileRows * kPitch * 4, 7 * kT * TY);
  const int bx = blockIdx.x * kT, by = y0 + blockIdx.y * TY, p = blockIdx.z;
  const uint8_t* src = in + (long long)(p / ia.channels) * ia.batch_stride + (long long)(p % ia.channels) * ia.chan_stride;
  const int tid = threadIdx.x;
  for (int i = tid; i < kTileRows * kRows; i += 256) {
    const int r = i / kRows, c = i - r * kRows;
    const int gy = min(max(by + r - kHalo, 0), H - 1), gx = min(max(bx + c - kHalo, 0), W - 1);
    tile[r * kPitch + c] = cell::split_px(__ldcg(src + (long long)gy * ia.row_stride + (long long)gx * ia.pix_stride));
  }
  __syncthreads();
  const int tx = tid & 31, tq = tid >> 5;
  Pk<OC> own[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) own[j].zero();
  if (PIPE) {
    pipelined_passes<Fmt, OC, LD, NJ>(t, tile, X, tx, tq, tid, own);
  } else {
    group_pass<Fmt, OC, LD, 0, NJ>(t, tile, X, tx, tq, tid, own);
    group_pass<Fmt, OC, LD, 1, NJ>(t, tile, X, tx, tq, tid, own);
    group_pass<Fmt, OC, LD, 2, NJ>(t, tile, X, tx, tq, tid, own);
    group_pass<Fmt, OC, LD, 3, NJ>(t, tile, X, tx, tq, tid, own);
    group_pass<Fmt, OC, LD, 4, NJ>(t, tile, X, tx, tq, tid, own);
  }
  __syncthreads();

  const int x = bx + tx;
  if (x >= W) return;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int ty = tq + 8 * j, y = by + ty;
    if (y >= y1) continue;
    Pk<OC> s = own[j];
#pragma unroll
    for (int a = 0; a < 7; ++a) s.add(X.get(a * kT * TY + ty * kT + tx));
    int n[3];
    s.get(n);
#pragma unroll
    for (int k = 0; k < OC; ++k) {
      int v;
      if (STAGE == 1) {
        v = n[k] <= 0 ? 0 : min(rhe_div(n[k], 48), 255);
      } else {
        const int u = n[k] + 127 * 192;
        v = u <= 0 ? 0 : min(rhe_div(u, 192), 255);
      }
      __stcg(out + ((long long)(p * OC + k) * H + y) * W + x, (uint8_t)v);
    }
  }
}

template <int OC, int NJ = 4, bool X6 = false>
constexpr size_t smem_bytes() { return (size_t)(8 * NJ + 2 * kHalo) * kPitch * 4 + (size_t)7 * kT * 8 * NJ * Xch<OC, X6>::kBytes; }

}  // namespace pwk

// Builds the paired-window copies of the stage-2 tables (6 window families).  An experiments build adds the stage-1
// copies (family f reads table f >> 1) and the cell-pair tables.
int build_pw_tables(lerf_luts_impl* L) {
  const int oC = L->oC2;
#ifdef LERF_EXPERIMENTS
  constexpr int kPlanes = 64;  // the unfolded variants read every order plane
#else
  constexpr int kPlanes = 32;  // production reads folded: codes with t1 < 2 only (lut_pw.cuh prepare_t<true>)
#endif
  const size_t b2 = pw::table_bytes(oC, kPlanes);
#ifdef LERF_EXPERIMENTS
  const size_t b1 = pw::table_bytes(1);
  const size_t off2 = 6 * b1;
#else
  const size_t off2 = 0;
#endif
  const size_t total = off2 + 6 * b2;
  cudaError_t e = cudaMalloc(&L->pw_block, total);
  if (e != cudaSuccess) return fail(LERF_ENOMEM, "cudaMalloc(%zu) for the paired-window LUT block failed: %s", total, cudaGetErrorString(e));
  L->pw_block_bytes = total;
  dim3 grid(65536 / 256, kPlanes);
  for (int f = 0; f < 6; ++f) {
    uint8_t* d2 = (uint8_t*)L->pw_block + off2 + f * b2;
    // the row-major device copy of an oC = 3 table is padded to 4 bytes per entry
    pwk::pw_repack_kernel<<<grid, 256>>>((const int8_t*)L->s2[f], oC, oC == 3 ? 4 : 1, f, d2);
    L->pw2[f] = d2;
#ifdef LERF_EXPERIMENTS
    uint8_t* d1 = (uint8_t*)L->pw_block + f * b1;
    pwk::pw_repack_kernel<<<grid, 256>>>(L->s1[f >> 1], 1, 1, f, d1);
    L->pw1[f] = d1;
#endif
  }
#ifdef LERF_EXPERIMENTS
  // cell-pair tables (oC = 1 only): stage 1 always, stage 2 for LeRF-L
  const size_t cpb = (size_t)65536 * 32;
  e = cudaMalloc(&L->cp_block, (oC == 1 ? 12 : 6) * cpb);
  if (e != cudaSuccess) return fail(LERF_ENOMEM, "cudaMalloc for the cell-pair LUT block failed: %s", cudaGetErrorString(e));
  for (int f = 0; f < 6; ++f) {
    uint8_t* d1 = (uint8_t*)L->cp_block + f * cpb;
    pwk::cp_repack_kernel<<<65536 / 256, 256>>>(L->s1[f >> 1], f, d1);
    L->cp1[f] = d1;
    if (oC == 1) {
      uint8_t* d2 = (uint8_t*)L->cp_block + (6 + f) * cpb;
      pwk::cp_repack_kernel<<<65536 / 256, 256>>>((const int8_t*)L->s2[f], f, d2);
      L->cp2[f] = d2;
    }
  }
#endif
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(LERF_ECUDA, "paired-window LUT repack failed: %s", cudaGetErrorString(e));
  return LERF_OK;
}

int launch_stage_pw(const lerf_luts_impl* L, int stage, const uint8_t* in, const InAddr& ia, int planes, int H, int W,
                    int y0, int y1, uint8_t* out, int variant, cudaStream_t st) {
  if (!L->pw_block) return fail(LERF_EUNSUPPORTED, "paired-window tables were not built for this LUT set");
  dim3 grid((W + pwk::kT - 1) / pwk::kT, (y1 - y0 + pwk::kT - 1) / pwk::kT, planes);
#define LERF_GO(F, S, O, LD, B)                                                                                   \
  {                                                                                                               \
    static bool attr_set = false; /* per instantiation: > 48 KB of dynamic shared memory needs the opt-in */       \
    if (!attr_set) {                                                                                              \
      LERF_CUDA(cudaFuncSetAttribute(pwk::lut_stage_pw_kernel<F, S, O, LD, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)pwk::smem_bytes<O>()));                                                 \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    pwk::lut_stage_pw_kernel<F, S, O, LD, B><<<grid, 256, pwk::smem_bytes<O>(), st>>>(t, in, ia, H, W, y0, y1, out); \
  }
  using pwk::FmtPW;
  using PWF3 = pwk::FmtPW<3, true>;  // folded lookups (lut_pw.cuh prepare_t<true>): half the order planes are read
  using PWF1 = pwk::FmtPW<1, true>;
  pwk::Tables t;
#define LERF_GO_PIPE(B)                                                                                                       \
  {                                                                                                                            \
    static bool attr_set = false;                                                                                              \
    if (!attr_set) {                                                                                                           \
      LERF_CUDA(cudaFuncSetAttribute(pwk::lut_stage_pw_kernel<FmtPW<3>, 2, 3, 1, B, 4, true>,                                   \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pwk::smem_bytes<3>()));                   \
      attr_set = true;                                                                                                         \
    }                                                                                                                          \
    pwk::lut_stage_pw_kernel<FmtPW<3>, 2, 3, 1, B, 4, true><<<grid, 256, pwk::smem_bytes<3>(), st>>>(t, in, ia, H, W, y0, y1, out); \
  }
#define LERF_GO_X6(B)                                                                                                         \
  {                                                                                                                            \
    static bool attr_set = false;                                                                                              \
    if (!attr_set) {                                                                                                           \
      LERF_CUDA(cudaFuncSetAttribute(pwk::lut_stage_pw_kernel<FmtPW<3>, 2, 3, 1, B, 4, false, true>,                            \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pwk::smem_bytes<3, 4, true>()));          \
      attr_set = true;                                                                                                         \
    }                                                                                                                          \
    pwk::lut_stage_pw_kernel<FmtPW<3>, 2, 3, 1, B, 4, false, true><<<grid, 256, pwk::smem_bytes<3, 4, true>(), st>>>(t, in, ia, H, W, y0, y1, out); \
  }
#ifdef LERF_EXPERIMENTS
  // (x) 32 x 8 tiles, one pixel per thread, for launches that do not fill the GPU (one 256 x 256 image is 192 blocks of
  // 32 x 32): four times the blocks, but half again as many halo windows -- measured SLOWER on cfg-1 (21.7 vs 17.8 us).
  dim3 grid_small((W + pwk::kT - 1) / pwk::kT, (y1 - y0 + 7) / 8, planes);
#define LERF_GO_SMALL(F, O)                                                                                                  \
  pwk::lut_stage_pw_kernel<F, 2, O, 1, 3, 1><<<grid_small, 256, pwk::smem_bytes<O, 1>(), st>>>(t, in, ia, H, W, y0, y1, out);
  using pwk::FmtCP;
  const bool cp = variant >= 10;  // variants 10+: cell-pair format (oC = 1)
  if (cp && stage == 2 && L->oC2 != 1) return fail(LERF_EUNSUPPORTED, "cell-pair tables exist for oC = 1 only");
  for (int i = 0; i < 6; ++i) t.t[i] = cp ? (stage == 1 ? L->cp1[i] : L->cp2[i]) : (stage == 1 ? L->pw1[i] : L->pw2[i]);
  if (cp) {
    if (stage == 1) {
      switch (variant) {
        case 11: LERF_GO(FmtCP, 1, 1, 0, 3) break;
        case 12: LERF_GO(FmtCP, 1, 1, 0, 5) break;
        case 13: LERF_GO(FmtCP, 1, 1, 1, 4) break;
        default: LERF_GO(FmtCP, 1, 1, 0, 4)
      }
    } else {
      switch (variant) {
        case 11: LERF_GO(FmtCP, 2, 1, 0, 3) break;
        default: LERF_GO(FmtCP, 2, 1, 0, 4)
      }
    }
  } else if (stage == 1) {
    switch (variant) {
      case 1: LERF_GO(FmtPW<1>, 1, 1, 0, 3) break;
      case 2: LERF_GO(FmtPW<1>, 1, 1, 1, 4) break;
      case 3: LERF_GO(FmtPW<1>, 1, 1, 0, 4) break;
      default: LERF_GO(FmtPW<1>, 1, 1, 1, 3)
    }
  } else if (L->oC2 == 3) {
    switch (variant) {
      case 1: LERF_GO(FmtPW<3>, 2, 3, 0, 3) break;
      case 2: LERF_GO(FmtPW<3>, 2, 3, 1, 2) break;
      case 3: LERF_GO_SMALL(FmtPW<3>, 3) break;
      case 4: LERF_GO_PIPE(2) break;
      case 5: LERF_GO_PIPE(3) break;
      case 6: LERF_GO_PIPE(1) break;
      case 7: LERF_GO_X6(4) break;
      case 8: LERF_GO_X6(3) break;
      case 9: LERF_GO(FmtPW<3>, 2, 3, 1, 3) break;  // unfolded lookups (r2a .. r2e production)
      default: LERF_GO(PWF3, 2, 3, 1, 3)
    }
  } else {
    switch (variant) {
      case 1: LERF_GO(FmtPW<1>, 2, 1, 0, 3) break;
      case 9: LERF_GO(FmtPW<1>, 2, 1, 1, 3) break;
      default: LERF_GO(PWF1, 2, 1, 1, 3)
    }
  }
#else
  (void)variant;
  if (stage != 2) return fail(LERF_EUNSUPPORTED, "the window kernel serves stage 2 in this build");
  for (int i = 0; i < 6; ++i) t.t[i] = L->pw2[i];
  if (L->oC2 == 3) LERF_GO(PWF3, 2, 3, 1, 3)
  else LERF_GO(PWF1, 2, 1, 1, 3)
#endif
#undef LERF_GO
#undef LERF_GO_PIPE
#undef LERF_GO_X6
#ifdef LERF_EXPERIMENTS
#undef LERF_GO_SMALL
#endif
  LERF_LAUNCHED();
  return LERF_OK;
}

}  // namespace lerf

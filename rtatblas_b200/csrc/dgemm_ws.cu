// Persistent, warp-specialised DGEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C
//
// Replaces cublasDgemm behind MatrixMult / MatrixMultAlloc (reference src/matrix_ops/matrixop.h:15-46,
// :347-440).  FP64 has no tcgen05 kind, so the math is DMMA.8x8x4 issued by eight consumer warps; what
// is Blackwell/Hopper-style here is the data path:
//
//   producer   TMA (cp.async.bulk.tensor, 128 B swizzle) issued by ONE elected thread into a
//              6-stage shared-memory ring guarded by full/empty mbarriers.  The operand's stored
//              orientation is consumed in place: a K-contiguous operand (op T for A, op N for B) is
//              one [16 k] x [128 rows] box per stage, an M/N-contiguous operand eight [16 mn] x [16 k]
//              boxes.  This is the "transpose folded into the TMA descriptor" plan variant.
//              Operands TMA cannot address (base not 16 B aligned or odd ld, e.g. 4097/3001 in
//              config 3) are fed by four producer warps with 8-byte cp.async that complete on the same
//              mbarriers and write the same swizzled layout ("pad folded into the load").
//   consumers  8 warps, 64x32 accumulator tile each (64 FP64 registers), fragments read with
//              conflict-free LDS.64: rows/columns of each 8x8 DMMA block are assigned to lanes
//              through a permutation chosen so that the 16 lanes of a half-warp hit 16 different
//              8-byte bank pairs under the 128 B swizzle (see frag_*).  No __syncthreads in the main
//              loop; the only per-k-tile synchronisation is one mbarrier wait and one arrive per warp.
//   schedule   persistent, hybrid stream-K: the last (#tiles mod #CTAs) + #CTAs tiles are cut into equal
//              runs of k-tiles, one run per CTA, so every SM finishes at the same time instead of
//              idling through a ragged last wave (and a 1024^3 problem, 64 tiles, still occupies all
//              148 SMs).  A CTA whose run starts inside a tile stores its accumulators, fragment-major,
//              to a per-CTA slot in the handle's arena and raises a flag; the CTA that owns the tile's
//              first k-tile finishes last, adds the slots in CTA order (deterministic) and runs the
//              epilogue.  The remaining tiles are data-parallel: CTA c walks tiles c, c+grid, ... in a
//              grouped raster so resident CTAs share operand strips in L2.  The producer runs ahead
//              across segment boundaries, so the ring is full when the consumers leave an epilogue.
#include <cuda.h>
#include "common.cuh"

namespace rtat_b200 {

namespace ws {

constexpr int BK = 16;
constexpr int CPASYNC_PRODUCER_WARPS = 4;

// CTA tile configuration.  Cfg128: 128x128 tiles, eight 64x32 consumer warps, one CTA per SM — the
// steady-state workhorse (98 % DMMA pipe utilisation on 8192^3).  Cfg64: 64x64 tiles, four 32x32 consumer
// warps, two CTAs per SM — four times smaller stream-K slots and finer edge granularity, for problems with
// few or ragged tiles (1024^3; 4097 x 3001).
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int CTAS_PER_SM_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, CTAS_PER_SM = CTAS_PER_SM_;
  static constexpr int CONSUMER_WARPS = (BM / WM) * (BN / WN);
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLOT_DOUBLES = CONSUMER_WARPS * 32 * (WM / 8) * (WN / 8) * 2;
  static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  // consumers of a ragged tile skip single DMMA blocks (64x64 tiles) or only whole warps (128x128: at the register cap)
  static constexpr bool FINE_EDGES = BM_ == 64;
  // 64x32 warp tiles (128 accumulator registers): the consumers need more registers than an even split of the file gives
  // them (setmaxnreg, see the kernel); 64x64 tiles hold 64 accumulator registers and do not (measured with the split: 1-3 %
  // slower)
  static constexpr bool SPLIT_REGS = BM_ == 128;
};
using Cfg128 = TileCfg<128, 128, 64, 32, 6, 1>;
using Cfg64 = TileCfg<64, 64, 32, 32, 6, 2>;
// 128x64 tiles, four 64x32 consumer warps, two CTAs per SM (4 stages of 24 KB each): the warp tile — and with it the
// fragment loads and ring hand-offs per DMMA — of the 128x128 configuration at half the tile area: twice the tiles to
// balance, half the stream-K slot, N quantised in 64s.
using Cfg128x64 = TileCfg<128, 64, 64, 32, 4, 2>;

struct Params {
  int m, n, k;
  const double *A;
  int lda;
  const double *B;
  int ldb;
  double *C;
  int ldc;
  double alpha, beta;
  int tiles_m, tiles_n, num_tiles, kt;
  // hybrid stream-K schedule (host-computed): tiles [0, dp_tiles) are data-parallel, tiles
  // [dp_tiles, num_tiles) are covered by sk_units = sk_tiles * kt k-tile units split evenly
  int dp_tiles;
  long long sk_units;
  double *sk_slots;        // gridDim.x slots of CONSUMER_WARPS*32*64 doubles
  unsigned *sk_flags;      // gridDim.x flags, compared against sk_generation
  unsigned sk_generation;
  // SYRK (TRI kernels): 1 = only the lower triangle of C is computed and written, 2 = upper.  Tiles are
  // then the tiles_m (tiles_m + 1) / 2 tiles that touch the triangle, numbered row by row.
  int tri;
  // Tile order (GEMM): the tm_full x tn_full full tiles first (grouped raster), then the ragged last column of tiles,
  // then the ragged last row, then the corner.  The consumers of a partial tile skip the DMMA blocks that lie wholly
  // outside C, so those tiles are cheap; last in the order they shorten the tail instead of opening a hole in a wave.
  // (Charging them less in the stream-K split — CTAs given equal weighted lengths — measured 5-20 % slower on config 3
  // than equal k-tile counts: the ring protocol and operand delivery of a partial tile cost what a full one's do.)
  int tm_full, tn_full, n_int, e_col, e_row;
  // L2 eviction priority of the TMA loads (option dgemm.l2_hints, off by default).  Idea: in the grouped raster the 8 row
  // panels of A are re-read by every wave of the sweep across n while a column panel of B is used by 8 tiles that run at
  // the same time, so keep A (evict_last) and stream B (evict_first).  Measured on config 2 (ncu): DRAM reads 9.7 -> 11.6
  // GB, L2 hit rate 78.9 -> 75.7 %, time unchanged (DRAM is at 4 % of its bandwidth either way): the CTAs that share a B
  // panel drift apart in k, and an evict_first line is gone before the laggard asks for it.
  unsigned long long l2_a, l2_b;
  // rows of op(A) / columns of op(B) the operand's 3-D map covers (see tma_load_3d; 0 = no 3-D map): a tile wholly below
  // the limit is ONE box per stage, the ragged last tile row / column takes the 2-D boxes
  int a_3d, b_3d;
  int group;       // grouped raster: tile rows per sweep across n
  int tail_steps;  // DMMA steps of the last k-tile that see a valid k (1..4; 4 when k is a multiple of 16): see mma_ktile
  int prefetch_c;  // beta != 0: consumers prefetch their part of the C tile into L2 when the tile starts (option dgemm.prefetch_c)
};

// Tile t of a triangular enumeration -> (i, j) with i >= j: t = i (i + 1) / 2 + j.
__device__ __forceinline__ void tri_decode(int t, int &i, int &j) {
  i = int((__fsqrt_rn(8.0f * float(t) + 1.0f) - 1.0f) * 0.5f);
  while ((long long)i * (i + 1) / 2 > t) i--;
  while ((long long)(i + 1) * (i + 2) / 2 <= t) i++;
  j = t - int((long long)i * (i + 1) / 2);
}

// One contiguous piece of work: k-tiles [kb, ke) of `tile`.
struct Segment {
  int tile, kb, ke;
};

// Walks the segments of CTA `cta`: first its stream-K run (cut at tile boundaries), then its
// data-parallel tiles.  Producer and consumers iterate with identical state.
struct SegmentIter {
  long long u, u1;
  int next_dp, P, dp_tiles, kt;
  __device__ SegmentIter(const Params &p, int cta, int nctas)
      : u(p.sk_units * cta / nctas), u1(p.sk_units * (cta + 1) / nctas), next_dp(cta), P(nctas), dp_tiles(p.dp_tiles), kt(p.kt) {}
  __device__ bool next(Segment &s) {
    if (u < u1) {
      int t = int(u / kt);
      s.kb = int(u - (long long)t * kt);
      long long rem = u1 - u;
      s.ke = (rem < kt - s.kb) ? s.kb + int(rem) : kt;
      s.tile = dp_tiles + t;
      u += s.ke - s.kb;
      return true;
    }
    if (next_dp < dp_tiles) {
      s.tile = next_dp;
      s.kb = 0;
      s.ke = kt;
      next_dp += P;
      return true;
    }
    return false;
  }
};

// ---- mbarrier / TMA primitives ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// A consumer warp hands a ring stage back with release_stage(): a CTA-scope fence, THEN the arrive on the stage's "empty"
// barrier.  The fence is load-bearing.  ptxas schedules the arrive (SASS SYNCS.ARRIVE) directly behind the last LDS of the
// k-tile — the DMMAs that consume that load are moved after it — and gives it no scoreboard wait on the load, and the
// hardware does not order a shared-memory read against a later arrive either: when this warp's arrive is the last one the
// stage is waiting for, the producer's refill can reach shared memory before the read does.  Found with 16-byte fragment
// loads on 64x64 tiles, two CTAs per SM (4-14 % of 4096 x 1024 x 1024 runs took a fragment from the NEXT fill of the stage:
// with A(i, k) = k / 16 the errors are +6 per k-value, always in the load issued last before the arrive); 0 failures in
// 36 000 runs with the fence, any of membar.cta / fence.acq_rel.cta / fence.proxy.async, none with a fence after the
// full-barrier wait instead.  The 8-byte scheme and the 128x128 configuration never showed it (0 in 80 000 runs) but have
// the same instruction pattern, so every release goes through here.  Cost: not measurable (30.01 vs 30.01 ms on 8192^3).
// (profiles/r02_stage_release_race.txt.)
__device__ __forceinline__ void release_stage(uint32_t empty_bar, int lane) {
  asm volatile("fence.acq_rel.cta;\n" ::: "memory");
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(empty_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// `policy`: an L2 eviction-priority descriptor (createpolicy encoding; the three constants below)
constexpr unsigned long long L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull,
                             L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar, unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}
// 3-D form: an M/N-contiguous operand whose contiguous extent is a multiple of 16 is viewed as {16, k, extent / 16}
// (strides 8 B, ld, 128 B), so ONE box {16, BK, BMN / 16} lands as the same BMN / 16 swizzled [BK][128 B] atoms that
// BMN / 16 two-dimensional boxes would produce.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// ---- lane <-> element maps ---------------------------------------------------------------------------
// Fragments are fetched with 16-byte shared loads, two DMMA operands per instruction, and the k indices of a k-tile are
// dealt to the DMMA steps so that both kinds of operand can do that: step ks, lane q multiplies k = kmap(q, ks).  (Any
// partition of the 16 k's into four steps of four is a valid DGEMM as long as A and B agree on it.)
__host__ __device__ constexpr int kmap(int q, int ks) { return 2 * q + (ks & 1) + 8 * (ks >> 1); }
// K-contiguous operand: rows (m or n) are 128 B smem rows swizzled by (row & 7); lane index x (0..7) of a DMMA block
// owns physical row rho(x) of the block.  One 16-byte load = k = kmap(q, 2h), kmap(q, 2h + 1): steps 2h and 2h + 1 of one
// block.  For a fixed q the eight rows hit eight different 16 B chunks: 4 wavefronts per warp load, the minimum.
__host__ __device__ constexpr int rho(int x) { return 2 * (x & 3) + (x >> 2); }
// M/N-contiguous operand: 16 consecutive m (or n) form one 128 B smem row per k, swizzled by (k & 7).  Lane index x owns
// elements 2x and 2x + 1 of the atom: blocks 2p and 2p + 1 of atom p.  One 16-byte load = both blocks of one step; the
// eight lanes of a k-row cover its eight chunks: again 4 wavefronts.

// physical index (within the warp tile) of DMMA block `blk`, lane index x
template <bool KC>
__host__ __device__ constexpr int phys_index(int blk, int x) {
  return KC ? 8 * blk + rho(x) : 16 * (blk / 2) + 2 * x + (blk & 1);
}

// swizzled byte offset of element (mn, k) of an operand tile (what TMA produces; the cp.async
// producers write the same layout)
template <bool KC>
__device__ __forceinline__ uint32_t tile_offset(int mn, int k) {
  if (KC) return mn * 128 + ((((k >> 1) ^ (mn & 7)) & 7) << 4) + ((k & 1) << 3);
  return (mn >> 4) * (BK * 128) + k * 128 + (((((mn & 15) >> 1) ^ (k & 7)) & 7) << 4) + ((mn & 1) << 3);
}

// Per-thread fragment addressing for one operand: a base, two lane-dependent swizzle terms, immediates for the rest.
//   KC : pair(blk, h)  -> 16 bytes = steps 2h, 2h + 1 of block blk
//   !KC: pair(p, ks)   -> 16 bytes = blocks 2p, 2p + 1 of step ks
template <bool KC>
struct FragAddr {
  uint32_t base;
  uint32_t off[2];
  __device__ __forceinline__ void init(uint32_t tile_base_warp /* byte offset of the warp's first row */, int q, int x) {
    if (KC) {
      const int r = rho(x);
      base = tile_base_warp + r * 128;
#pragma unroll
      for (int h = 0; h < 2; h++) off[h] = uint32_t(((q + 4 * h) ^ r) & 7) << 4;  // chunk k >> 1 = q + 4h
    } else {
      base = tile_base_warp + 2 * q * 128;                                          // k-row 2q (+ (ks & 1) + 8 (ks >> 1))
#pragma unroll
      for (int e = 0; e < 2; e++) off[e] = uint32_t((x ^ (2 * q + e)) & 7) << 4;   // chunk x, k & 7 = 2q + (ks & 1)
    }
  }
  __device__ __forceinline__ uint32_t pair(int a, int b) const {
    if (KC) return base + a * 1024 + off[b];
    return base + a * (BK * 128) + (b & 1) * 128 + (b >> 1) * 1024 + off[b & 1];
  }
};

// Fragments of steps 2h, 2h + 1 for the first `lim` blocks of an operand: f[e][blk], e = step parity
template <bool KC, int NBLK, bool ALL>
__device__ __forceinline__ void load_half(double (&f)[2][NBLK], const FragAddr<KC> &fa, uint32_t sbase, int h, int lim) {
  if (KC) {
#pragma unroll
    for (int blk = 0; blk < NBLK; blk++)
      if (ALL || blk < lim) {
        const double2 v = lds128(sbase + fa.pair(blk, h));
        f[0][blk] = v.x;
        f[1][blk] = v.y;
      }
  } else {
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int p = 0; p < NBLK / 2; p++)
        if (ALL || 2 * p < lim) {
          const double2 v = lds128(sbase + fa.pair(p, 2 * h + e));
          f[e][2 * p] = v.x;
          f[e][2 * p + 1] = v.y;
        }
  }
}

// DMMAs of one k-tile on one warp tile.  ALL: every block (interior tile); otherwise the first mi_act x ni_act blocks.
// FULL: all four DMMA steps; otherwise only the first `steps` of them — the last k-tile of a K that is not a multiple of 16
// holds k_rem valid k's and step ks multiplies k = 2q + (ks & 1) + 8 (ks >> 1) (kmap), so steps whose smallest k is beyond
// k_rem would only add the producers' zero fill (K = 2049, config 3: three of the 129th k-tile's four steps).
template <bool KCA, bool KCB, int MI, int NI, bool ALL, bool FULL>
__device__ __forceinline__ void mma_ktile(double (&acc)[MI][NI][2], const FragAddr<KCA> &fa_addr, const FragAddr<KCB> &fb_addr, uint32_t sbase,
                                          int mi_act, int ni_act, int steps) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (!FULL && 2 * h >= steps) break;
    double fa[2][MI], fb[2][NI];
    load_half<KCA, MI, ALL>(fa, fa_addr, sbase, h, mi_act);
    load_half<KCB, NI, ALL>(fb, fb_addr, sbase, h, ni_act);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      if (!FULL && 2 * h + e >= steps) break;
#pragma unroll
      for (int i = 0; i < MI; i++)
        if (ALL || i < mi_act) {
#pragma unroll
          for (int j = 0; j < NI; j++)
            if (ALL || j < ni_act) dmma(acc[i][j], fa[e][i], fb[e][j]);
        }
    }
  }
}

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}

// cp.async producer: one BMN x BK operand tile per call, 128 threads.  Element ownership is chosen so that everything
// but a pointer bump per element is loop-invariant: K-contiguous -> thread owns column kk = pt % 16 of rows
// pt/16 + 8i (row & 7 is constant, so the swizzle term is too); M/N-contiguous -> thread owns element pt % BMN of
// k-rows pt/BMN + (128/BMN) i (the swizzle term is a compile-time XOR per i).  Interior tiles — every k-tile but
// possibly the last, every tile but the last row / column of tiles — take a path without bounds checks: the producer
// warps share issue slots with the DMMA warps, and ~28 instructions per copy on the checked path cost the consumers
// 12 % of the pipe (ncu, config 3).
template <bool KC, int BMN>
__device__ __forceinline__ void issue_operand_cpasync(uint32_t s, const double *__restrict__ X, int ld, int mn0, int k0,
                                                      int MN, int K, int pt) {
  constexpr int PT = CPASYNC_PRODUCER_WARPS * 32;
  static_assert(PT == 128 && PT % BMN == 0, "ownership pattern below assumes 128 producer threads");
  const bool interior = mn0 + BMN <= MN && k0 + BK <= K;
  if (KC) {
    constexpr int PER = BMN * BK / PT;
    const int kk = pt % BK, r0 = pt / BK;
    const double *src = X + (size_t)(mn0 + r0) * ld + (k0 + kk);
    const size_t step = (size_t)(PT / BK) * ld;
    const uint32_t dst0 = s + tile_offset<true>(r0, kk);
    if (interior) {
#pragma unroll
      for (int i = 0; i < PER; i++) {
        cp_async8(dst0 + i * (PT / BK) * 128, src);
        src += step;
      }
    } else {
      const bool kok = (k0 + kk) < K;
#pragma unroll
      for (int i = 0; i < PER; i++) {
        bool ok = kok && (mn0 + r0 + (PT / BK) * i) < MN;
        cp_async8_zfill(dst0 + i * (PT / BK) * 128, ok ? src : X, ok ? 8 : 0);
        src += step;
      }
    }
  } else {
    constexpr int KSTEP = PT / BMN, PER = BK / KSTEP;
    const int mn = pt % BMN, kr0 = pt / BMN;  // kr0 < KSTEP <= 2, so (kr0 + KSTEP i) & 7 == ((KSTEP i) & 7) | kr0
    const double *src = X + (size_t)(k0 + kr0) * ld + (mn0 + mn);
    const size_t step = (size_t)KSTEP * ld;
    const uint32_t dst0 = s + (mn >> 4) * (BK * 128) + ((mn & 1) << 3) + kr0 * 128;
    const uint32_t c4 = (((mn & 15) >> 1) ^ kr0) << 4;
    if (interior) {
#pragma unroll
      for (int i = 0; i < PER; i++) {
        cp_async8(dst0 + i * KSTEP * 128 + (c4 ^ (((KSTEP * i) & 7) << 4)), src);
        src += step;
      }
    } else {
      const bool mok = (mn0 + mn) < MN;
#pragma unroll
      for (int i = 0; i < PER; i++) {
        bool ok = mok && (k0 + kr0 + KSTEP * i) < K;
        cp_async8_zfill(dst0 + i * KSTEP * 128 + (c4 ^ (((KSTEP * i) & 7) << 4)), ok ? src : X, ok ? 8 : 0);
        src += step;
      }
    }
  }
}

template <class Cfg, bool TA, bool TB, bool USE_TMA, bool TRI = false>
__global__ void __launch_bounds__((Cfg::CONSUMER_WARPS + ((USE_TMA && !Cfg::SPLIT_REGS) ? 1 : CPASYNC_PRODUCER_WARPS)) * 32, Cfg::CTAS_PER_SM)
dgemm_ws_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapA3,
                const __grid_constant__ CUtensorMap mapB3, const Params p) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
  constexpr int CONSUMER_WARPS = Cfg::CONSUMER_WARPS, A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr bool KCA = TA;    // op(A)(m,k): stored k x m when transposed => k contiguous
  constexpr bool KCB = !TB;   // op(B)(k,n): stored k x n when not transposed => k contiguous
  constexpr int PRODUCER_THREADS = USE_TMA ? 1 : CPASYNC_PRODUCER_WARPS * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[s] at +8s, empty[s] at +8(STAGES+s)
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (USE_TMA && threadIdx.x == CONSUMER_WARPS * 32) {
    // descriptors into the TMA unit's cache while the barriers are being set up, before the first load asks for them
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(KCA || p.a_3d == 0 ? &mapA : &mapA3) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(KCB || p.b_3d == 0 ? &mapB : &mapB3) : "memory");
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_base + 8 * s, PRODUCER_THREADS);
      mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  auto tile_coords = [&](int tile, int &m0, int &n0) {
    if (TRI) {
      int i, j;
      tri_decode(tile, i, j);
      m0 = (p.tri == 1 ? i : j) * BM;
      n0 = (p.tri == 1 ? j : i) * BN;
      return;
    }
    if (tile >= p.n_int) {  // ragged edge tiles, after all full ones
      tile -= p.n_int;
      if (tile < p.e_col) { m0 = tile * BM; n0 = p.tn_full * BN; return; }
      tile -= p.e_col;
      m0 = p.tm_full * BM;
      n0 = (tile < p.e_row ? tile : p.tn_full) * BN;
      return;
    }
    const int GROUP = p.group;  // tile rows swept together across the columns (option dgemm.group, default 8)
    int group_size = GROUP * p.tn_full;
    int first_m = (tile / group_size) * GROUP;
    int gm = min(p.tm_full - first_m, GROUP);
    m0 = (first_m + (tile % group_size) % gm) * BM;
    n0 = ((tile % group_size) / gm) * BN;
  };

  // Register file split (setmaxnreg, per warpgroup): every SM sub-partition hosts one producer warp and two consumer warps
  // (128x128 tiles; SPLIT_REGS), so what the producers give back — a TMA issuer needs next to nothing — lets the
  // consumers hold 128 accumulator registers plus a whole half k-tile of fragments without spilling.  The producer
  // warpgroup is always four warps (three of them idle with TMA) because the instruction is warpgroup-wide.
  // Budget: the kernel launches with LAUNCH_REGS per thread (what __launch_bounds__ leaves each of the CTAS_PER_SM
  // resident CTAs); the four producer warps shrink to PRODUCER_REGS and the consumers share what that frees.
  constexpr int PRODUCER_REGS = USE_TMA ? 40 : 56;
  constexpr int LAUNCH_REGS = 65536 / ((CONSUMER_WARPS + CPASYNC_PRODUCER_WARPS) * 32 * Cfg::CTAS_PER_SM) / 8 * 8;
  constexpr int CONSUMER_REGS = (LAUNCH_REGS + (LAUNCH_REGS - PRODUCER_REGS) * CPASYNC_PRODUCER_WARPS / CONSUMER_WARPS) / 8 * 8;
  static_assert(!Cfg::SPLIT_REGS || BN != 128 || CONSUMER_REGS == 232 || CONSUMER_REGS == 224, "128x128 split as measured");
  static_assert(!Cfg::SPLIT_REGS || BN != 64 || CONSUMER_REGS == 216 || CONSUMER_REGS == 200, "128x64 split: half a register file per CTA");
  if (warp >= CONSUMER_WARPS) {
    if (Cfg::SPLIT_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(PRODUCER_REGS));
    // ======================= producer =======================
    int stage = 0;
    uint32_t phase = 0;
    if (USE_TMA) {
      if (warp == CONSUMER_WARPS && lane == 0) {
        SegmentIter it(p, blockIdx.x, gridDim.x);
        Segment seg;
        while (it.next(seg)) {
          int m0, n0;
          tile_coords(seg.tile, m0, n0);
          for (int kt = seg.kb; kt < seg.ke; kt++) {
            mbar_wait(bar_base + 8 * (STAGES + stage), phase ^ 1);
            uint32_t full = bar_base + 8 * stage;
            uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
            mbar_arrive_expect_tx(full, STAGE_BYTES);
            int k0 = kt * BK;
            if (KCA) {
              tma_load_2d(sA, &mapA, k0, m0, full, p.l2_a);
            } else if (m0 + BM <= p.a_3d) {
              tma_load_3d(sA, &mapA3, 0, k0, m0 / 16, full, p.l2_a);
            } else {
#pragma unroll
              for (int b = 0; b < BM / 16; b++) tma_load_2d(sA + b * (BK * 128), &mapA, m0 + 16 * b, k0, full, p.l2_a);
            }
            if (KCB) {
              tma_load_2d(sB, &mapB, k0, n0, full, p.l2_b);
            } else if (n0 + BN <= p.b_3d) {
              tma_load_3d(sB, &mapB3, 0, k0, n0 / 16, full, p.l2_b);
            } else {
#pragma unroll
              for (int b = 0; b < BN / 16; b++) tma_load_2d(sB + b * (BK * 128), &mapB, n0 + 16 * b, k0, full, p.l2_b);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else {
      const int pt = threadIdx.x - CONSUMER_WARPS * 32;  // 0..127
      SegmentIter it(p, blockIdx.x, gridDim.x);
      Segment seg;
      while (it.next(seg)) {
        int m0, n0;
        tile_coords(seg.tile, m0, n0);
        for (int kt = seg.kb; kt < seg.ke; kt++) {
          mbar_wait(bar_base + 8 * (STAGES + stage), phase ^ 1);
          uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
          issue_operand_cpasync<KCA, BM>(sA, p.A, p.lda, m0, kt * BK, p.m, p.k, pt);
          issue_operand_cpasync<KCB, BN>(sB, p.B, p.ldb, n0, kt * BK, p.n, p.k, pt);
          cp_async_mbar_arrive_noinc(bar_base + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      asm volatile("cp.async.wait_all;\n" ::: "memory");
    }
    return;
  }

  // ======================= consumers =======================
  if (Cfg::SPLIT_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(CONSUMER_REGS));
  constexpr int MI = WM / 8, NI = WN / 8, KS = BK / 4;
  const int q = lane % 4, g = lane / 4;
  const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
  FragAddr<KCA> fa_addr;
  FragAddr<KCB> fb_addr;
  fa_addr.init(KCA ? wm0 * 128 : (wm0 / 16) * (BK * 128), q, g);
  fb_addr.init(A_BYTES + (KCB ? wn0 * 128 : (wn0 / 16) * (BK * 128)), q, g);
  // physical index (within the warp tile) of DMMA block blk, lane index x
  auto row_of = [](int blk, int x) { return phys_index<KCA>(blk, x); };
  auto col_of = [](int blk, int x) { return phys_index<KCB>(blk, x); };

  int stage = 0;
  uint32_t phase = 0;
  const double alpha = p.alpha, beta = p.beta;

  SegmentIter it(p, blockIdx.x, gridDim.x);
  Segment seg;
  const bool opt_prefetch_c = p.prefetch_c != 0;
  const int ctid = threadIdx.x;  // 0..255 (consumer threads come first)
  constexpr int SLOT_DOUBLES = Cfg::SLOT_DOUBLES;
  while (it.next(seg)) {
    int m0, n0;
    tile_coords(seg.tile, m0, n0);
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // beta != 0: the epilogue reads C.  Its loads come in batches of MI per column, each batch paying one memory latency
    // (C aliases itself, so loads cannot be hoisted over the stores); from DRAM that is ~1 us x 2 NI batches per tile, 5-7 %
    // of a 1024-deep product.  Ask for the warp's part of the C tile now, so that the epilogue finds it in L2.
    if (beta != 0.0 && seg.kb == 0 && opt_prefetch_c) {
      const int col = n0 + wn0 + lane;
      if (lane < WN && col < p.n) {
        const double *c0 = p.C + (size_t)col * p.ldc + m0 + wm0;
#pragma unroll
        for (int r = 0; r < WM; r += 16)
          if (m0 + wm0 + r < p.m) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(c0 + r));
      }
    }

    // DMMA blocks of this warp that touch C: a ragged tile's consumers skip the rest (blocks are 8 rows / columns of a
    // K-contiguous operand, pairs of blocks interleave within 16 of an M/N-contiguous one)
    const int rv = min(max(p.m - m0 - wm0, 0), WM), cv = min(max(p.n - n0 - wn0, 0), WN);
    const int mi_act = KCA ? (rv + 7) / 8 : 2 * ((rv + 15) / 16), ni_act = KCB ? (cv + 7) / 8 : 2 * ((cv + 15) / 16);
    const bool whole = mi_act == MI && ni_act == NI;
    static_assert(KS == 4, "two halves of two DMMA steps per k-tile");
    // the last k-tile of the tile runs apart when only some of its DMMA steps see a valid k (mma_ktile)
    const int ke_main = (seg.ke == p.kt && p.tail_steps < KS) ? seg.ke - 1 : seg.ke;
    if (whole || !Cfg::FINE_EDGES) {
      // 64-row warp tiles (no registers to spare for a second loop body) skip at warp granularity only
      const bool active = whole || (mi_act > 0 && ni_act > 0);
      for (int kt = seg.kb; kt < ke_main; kt++) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (active) mma_ktile<KCA, KCB, MI, NI, true, true>(acc, fa_addr, fb_addr, smem_base + stage * STAGE_BYTES, MI, NI, KS);
        release_stage(bar_base + 8 * (STAGES + stage), lane);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (ke_main < seg.ke) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (active) mma_ktile<KCA, KCB, MI, NI, true, false>(acc, fa_addr, fb_addr, smem_base + stage * STAGE_BYTES, MI, NI, p.tail_steps);
        release_stage(bar_base + 8 * (STAGES + stage), lane);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else {
      // ragged tile (or a warp with nothing to do): same ring protocol, only the blocks that touch C are multiplied
      const bool active = mi_act > 0 && ni_act > 0;
      for (int kt = seg.kb; kt < seg.ke; kt++) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (active)
          mma_ktile<KCA, KCB, MI, NI, false, false>(acc, fa_addr, fb_addr, smem_base + stage * STAGE_BYTES, mi_act, ni_act,
                                                    kt < ke_main ? KS : p.tail_steps);
        release_stage(bar_base + 8 * (STAGES + stage), lane);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }


    if (seg.kb > 0) {
      // stream-K contributor: this run started inside the tile.  Park the accumulators
      // (fragment-major, coalesced) in this CTA's slot and publish them.
      double *slot = p.sk_slots + (size_t)blockIdx.x * SLOT_DOUBLES + ctid;
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) {
          __stcg(slot + ((i * NI + j) * 2 + 0) * (CONSUMER_WARPS * 32), acc[i][j][0]);
          __stcg(slot + ((i * NI + j) * 2 + 1) * (CONSUMER_WARPS * 32), acc[i][j][1]);
        }
      // publication: the consumers' barrier orders every thread's stores before thread 0's release at GPU scope (releases
      // are cumulative over what happens-before them), so no thread needs a fence of its own — a MEMBAR.SC.GPU per thread
      // in round 1
      asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_WARPS * 32) : "memory");
      if (ctid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p.sk_flags + blockIdx.x), "r"(p.sk_generation) : "memory");
      continue;
    }
    if (seg.ke < p.kt) {
      // stream-K finisher: owns the tile's first k-tile; the CTAs after this one cover the rest of
      // the tile and parked their parts early in their runs.  Add them in CTA order.
      const long long tile_end = (long long)(seg.tile - p.dp_tiles + 1) * p.kt;
      for (int c = blockIdx.x + 1; c < (int)gridDim.x && p.sk_units * c / (long long)gridDim.x < tile_end; c++) {
        if (ctid == 0) {
          // The contributor is a CTA of this same launch; the grid never exceeds what the occupancy calculator says is
          // co-resident on an otherwise idle device, so it is running or done.  Should other work hold SMs (a second
          // stream-K launch on another stream), a contributor may not be scheduled while finishers spin: the wait is
          // bounded and the kernel then traps, so the failure surfaces as a CUDA error instead of a hung GPU.
          unsigned v;
          unsigned long long t0 = 0;
          for (unsigned polls = 0;; polls++) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.sk_flags + c) : "memory");
            if (v == p.sk_generation) break;
            if ((polls & 1023u) == 1023u) {
              unsigned long long now;
              asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
              if (t0 == 0) t0 = now;
              else if (now - t0 > 20000000000ull) __trap();  // 20 s
            }
          }
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_WARPS * 32) : "memory");
        const double *slot = p.sk_slots + (size_t)c * SLOT_DOUBLES + ctid;
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
          for (int j = 0; j < NI; j++) {
            acc[i][j][0] += __ldcg(slot + ((i * NI + j) * 2 + 0) * (CONSUMER_WARPS * 32));
            acc[i][j][1] += __ldcg(slot + ((i * NI + j) * 2 + 1) * (CONSUMER_WARPS * 32));
          }
      }
    }

    // epilogue: lane (q,g) of block (i,j) holds C[row(i,g)][col(j,2q+e)].  With beta != 0 the loads of EPI_COLS columns
    // (MI each) are issued together before the first store: a load-modify-store loop would serialise on one memory
    // latency per element (C aliases itself, so the compiler cannot hoist the loads), and nothing else runs on this SM
    // while its consumers are in the epilogue.  The fragment registers are dead here, which is what pays for the batch.
#ifndef RTAT_DGEMM_EPI_COLS
#define RTAT_DGEMM_EPI_COLS 2
#endif
    constexpr int EPI_COLS = RTAT_DGEMM_EPI_COLS;
    static_assert((2 * NI) % EPI_COLS == 0, "columns per thread divide into batches");
#pragma unroll
    for (int c0 = 0; c0 < 2 * NI; c0 += EPI_COLS) {
      bool ok[EPI_COLS][MI];
      double old[EPI_COLS][MI];
      double *ccol[EPI_COLS];
#pragma unroll
      for (int c = 0; c < EPI_COLS; c++) {
        const int j = (c0 + c) >> 1, e = (c0 + c) & 1;
        const int col = n0 + wn0 + col_of(j, 2 * q + e);
        ccol[c] = p.C + (size_t)col * p.ldc;
#pragma unroll
        for (int i = 0; i < MI; i++) {
          const int row = m0 + wm0 + row_of(i, g);
          ok[c][i] = col < p.n && row < p.m && !(TRI && (p.tri == 1 ? row < col : row > col));
          old[c][i] = 0.0;
        }
      }
      if (beta != 0.0) {
#pragma unroll
        for (int c = 0; c < EPI_COLS; c++)
#pragma unroll
          for (int i = 0; i < MI; i++)
            if (ok[c][i]) old[c][i] = ccol[c][m0 + wm0 + row_of(i, g)];
      }
#pragma unroll
      for (int c = 0; c < EPI_COLS; c++) {
        const int j = (c0 + c) >> 1, e = (c0 + c) & 1;
#pragma unroll
        for (int i = 0; i < MI; i++) {
          if (ok[c][i]) {
            double v = alpha * acc[i][j][e];
            if (beta != 0.0) v += beta * old[c][i];
            ccol[c][m0 + wm0 + row_of(i, g)] = v;
          }
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_t tensor_map_encoder() { return reinterpret_cast<encode_tiled_t>(tensor_map_encoder_raw()); }

static bool make_map(CUtensorMap *map, const double *base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems,
                     uint32_t box0, uint32_t box1) {
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld_elems * sizeof(double)};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  auto encode = tensor_map_encoder();
  if (!encode) return false;
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// {16, cols, rows / 16} view of a column-major rows x cols matrix (rows % 16 == 0), box {16, box_cols, box_rows / 16}
static bool make_map_3d(CUtensorMap *map, const double *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_cols,
                        uint32_t box_rows) {
  cuuint64_t dims[3] = {16, cols, rows / 16};
  cuuint64_t strides[2] = {ld_elems * sizeof(double), 16 * sizeof(double)};
  cuuint32_t box[3] = {16, box_cols, box_rows / 16};
  cuuint32_t estr[3] = {1, 1, 1};
  auto encode = tensor_map_encoder();
  if (!encode) return false;
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class Cfg, bool TA, bool TB, bool USE_TMA, bool TRI = false>
static int launch(rtat_b200_context *h, const Params &p, int grid, const CUtensorMap (&maps)[4], const char *name) {
  auto kern = dgemm_ws_kernel<Cfg, TA, TB, USE_TMA, TRI>;
  constexpr size_t SMEM_BYTES = Cfg::SMEM_BYTES;
  constexpr int CONSUMER_WARPS = Cfg::CONSUMER_WARPS;
  static std::atomic<uint64_t> configured{0};  // per template instantiation
  if (int st = configure_dynamic_smem(kern, SMEM_BYTES, h->device, configured)) return st;
  // with setmaxnreg the producer warpgroup is always complete
  int threads = (CONSUMER_WARPS + ((USE_TMA && !Cfg::SPLIT_REGS) ? 1 : CPASYNC_PRODUCER_WARPS)) * 32;
  kern<<<grid, threads, SMEM_BYTES, h->stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  return check_launch(h, name);
}

}  // namespace ws

int dgemm_skinny(rtat_b200_context *h, int r, int N, int K, double alpha, const double *S, long long ss_i, long long ss_k, const double *X,
                 long long xs_k, long long xs_j, double beta, double *Y, long long ys_i, long long ys_j);

bool dgemm_ws_tma_ok(const double *A, int lda, const double *B, int ldb) {
  auto ok = [](const void *ptr, int ld) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && ld % 2 == 0; };
  return ok(A, lda) && ok(B, ldb);
}

// Stream-K finishers spin on flags raised by other CTAs, so every CTA of the grid must be resident at
// once: never launch more CTAs than the occupancy calculator says fit.
template <class Cfg, bool TA, bool TB, bool USE_TMA, bool TRI = false>
static int resident_ctas_per_sm() {
  static int cached = -1;
  if (cached < 0) {
    auto kern = ws::dgemm_ws_kernel<Cfg, TA, TB, USE_TMA, TRI>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    int occ = 0;
    int threads = (Cfg::CONSUMER_WARPS + ((USE_TMA && !Cfg::SPLIT_REGS) ? 1 : ws::CPASYNC_PRODUCER_WARPS)) * 32;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, Cfg::SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      occ = 1;
    }
    cached = occ < 1 ? 1 : (occ > Cfg::CTAS_PER_SM ? Cfg::CTAS_PER_SM : occ);
  }
  return cached;
}
template <class Cfg>
static int resident_ctas_per_sm(bool ta, bool tb, bool tma, bool tri) {
  if (tri) {
    if (ta) return tma ? resident_ctas_per_sm<Cfg, true, false, true, true>() : resident_ctas_per_sm<Cfg, true, false, false, true>();
    return tma ? resident_ctas_per_sm<Cfg, false, true, true, true>() : resident_ctas_per_sm<Cfg, false, true, false, true>();
  }
#define RTAT_OCC(TA_, TB_) (tma ? resident_ctas_per_sm<Cfg, TA_, TB_, true>() : resident_ctas_per_sm<Cfg, TA_, TB_, false>())
  if (!ta && !tb) return RTAT_OCC(false, false);
  if (ta && !tb) return RTAT_OCC(true, false);
  if (!ta && tb) return RTAT_OCC(false, true);
  return RTAT_OCC(true, true);
#undef RTAT_OCC
}

// Cost of one k-tile of a tile with `rows` x `cols` valid elements (0 = the full extent), in sixteenths of a full tile's:
// the DMMA blocks of the busiest SM sub-partition.  Warp w sits on sub-partition w % 4, owns M-slice w % (BM / WM) and
// N-slice w / (BM / WM); a warp multiplies only the 8-row / 8-column blocks that touch C (pairs of blocks for an
// M/N-contiguous operand).  Floor of 3/16: operand delivery and the ring protocol do not shrink with the tile.
template <class Cfg>
static int tile_cost(bool a_mn_contiguous, int rows, int cols) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN;
  if (rows <= 0) rows = BM;
  if (cols <= 0) cols = BN;
  int load[4] = {0, 0, 0, 0};
  for (int w = 0; w < Cfg::CONSUMER_WARPS; w++) {
    int rv = rows - (w % (BM / WM)) * WM, cv = cols - (w / (BM / WM)) * WN;
    rv = rv < 0 ? 0 : (rv > WM ? WM : rv);
    cv = cv < 0 ? 0 : (cv > WN ? WN : cv);
    // op(A) is M-contiguous when A is not transposed; op(B) N-contiguous when B is transposed.  Rounding pairs of blocks
    // for both operands is the conservative choice for B.
    int mi = a_mn_contiguous ? 2 * ((rv + 15) / 16) : (rv + 7) / 8, ni = 2 * ((cv + 15) / 16);
    if (!Cfg::FINE_EDGES && mi > 0 && ni > 0) { mi = WM / 8; ni = WN / 8; }
    load[w % 4] += mi * ni;
  }
  int worst = 0, full = (WM / 8) * (WN / 8) * (Cfg::CONSUMER_WARPS > 4 ? Cfg::CONSUMER_WARPS / 4 : 1);
  for (int x : load) worst = x > worst ? x : worst;
  int c = (16 * worst + full - 1) / full;
  return c < 3 ? 3 : (c > 16 ? 16 : c);
}

template <class Cfg>
static int run_cfg(rtat_b200_context *h, bool ta, bool tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B,
                   int ldb, double beta, double *C, int ldc, bool allow_tma, const char *tile_name, int tri) {
  using namespace ws;
  constexpr int BM = Cfg::BM, BN = Cfg::BN;
  Params p{};
  p.m = m; p.n = n; p.k = k;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.alpha = alpha; p.beta = beta;
  p.tiles_m = (m + BM - 1) / BM;
  p.tiles_n = (n + BN - 1) / BN;
  long long tiles = (long long)p.tiles_m * p.tiles_n;
  p.tri = tri;
  if (tri) {
    if (m != n || ta == tb) return RTAT_B200_STATUS_INVALID_VALUE;
    tiles = (long long)p.tiles_m * (p.tiles_m + 1) / 2;
  }
  if (tiles > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
  p.num_tiles = (int)tiles;
  p.kt = (k + BK - 1) / BK;
  {
    const int k_rem = k - (p.kt - 1) * BK;  // valid k's of the last k-tile, 1..16 (k >= 1 here)
    p.tail_steps = (k_rem >= 10 || opt(h, "dgemm.ktail", 1) == 0) ? 4 : k_rem == 9 ? 3 : k_rem >= 2 ? 2 : 1;
  }
  // ---- schedule: hybrid stream-K ---------------------------------------------------------------
  // grid: CTAS_PER_SM resident CTAs per SM, fewer when there are not even MIN_UNITS k-tiles of work per CTA.
  constexpr int MIN_UNITS = 4;
  long long units = tiles * p.kt;
  bool use_tma = allow_tma && dgemm_ws_tma_ok(A, lda, B, ldb);
  const int max_grid = h->sm_count * resident_ctas_per_sm<Cfg>(ta, tb, use_tma, tri != 0);
  int grid = max_grid;
  if (units < (long long)grid * MIN_UNITS) grid = (int)((units + MIN_UNITS - 1) / MIN_UNITS);
  if (grid < 1) grid = 1;
  int sk_mode = opt(h, "dgemm.streamk", 1);  // 0 = data-parallel only (for A/B measurements)
  int rem = p.num_tiles % grid;
  int sk_tiles = 0;
  if (sk_mode && (rem != 0 || p.num_tiles < grid) && p.kt >= 8) {
    // cover the ragged wave plus one full wave, so every CTA's run is 1-2 tiles long
    sk_tiles = p.num_tiles <= grid ? p.num_tiles : rem + grid;
    if (sk_tiles > p.num_tiles) sk_tiles = p.num_tiles;
  }
  if (!sk_tiles && p.num_tiles < grid) grid = p.num_tiles;
  p.dp_tiles = p.num_tiles - sk_tiles;
  p.sk_units = (long long)sk_tiles * p.kt;
  // tile order and the cost of the ragged tiles (see Params)
  if (tri) {
    p.tm_full = p.tiles_m; p.tn_full = p.tiles_n; p.n_int = p.num_tiles; p.e_col = p.e_row = 0;
  } else {
    p.tm_full = m / BM; p.tn_full = n / BN;
    p.n_int = p.tm_full * p.tn_full;
    p.e_col = (n % BN) ? p.tm_full : 0;
    p.e_row = (m % BM) ? p.tn_full : 0;
  }
  if (sk_tiles) {
    // sized for the largest grid of either tile configuration up front, so no call ever pays a
    // reallocation (and its sync) later
    size_t arena = sizeof(double) * (size_t)Cfg128::SLOT_DOUBLES * h->sm_count * Cfg128::CTAS_PER_SM;
    size_t arena64 = sizeof(double) * (size_t)Cfg64::SLOT_DOUBLES * h->sm_count * Cfg64::CTAS_PER_SM;
    const size_t arena12864 = sizeof(double) * (size_t)Cfg128x64::SLOT_DOUBLES * h->sm_count * Cfg128x64::CTAS_PER_SM;
    if (arena64 < arena12864) arena64 = arena12864;
    int st = ensure_arena(h, arena > arena64 ? arena : arena64);
    if (st) return st;
    if (!h->sk_flags) {
      if (cudaMalloc(&h->sk_flags, 4096) != cudaSuccess || cudaMemsetAsync(h->sk_flags, 0, 4096, h->stream) != cudaSuccess) {
        (void)cudaGetLastError();
        return RTAT_B200_STATUS_ALLOC_FAILED;
      }
      h->sk_generation = 0;
    }
    p.sk_slots = static_cast<double *>(h->arena);
    p.sk_flags = static_cast<unsigned *>(h->sk_flags);
    p.sk_generation = ++h->sk_generation;
  }
  p.prefetch_c = opt(h, "dgemm.prefetch_c", 1);
  p.group = opt(h, "dgemm.group", 8);
  if (p.group < 1) p.group = 1;
  const bool hints = !tri && opt(h, "dgemm.l2_hints", 0) != 0;
  p.l2_a = hints ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  p.l2_b = hints ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  CUtensorMap maps[4];  // A, B (2-D boxes) and their 3-D views
  memset(maps, 0, sizeof maps);
  if (use_tma) {
    // M/N-contiguous operands: one 3-D box per stage instead of BMN / 16 two-dimensional ones for every tile that lies
    // wholly inside the multiple of 16 the view can cover; only the ragged last tile row / column uses the 2-D boxes
    const bool box3 = opt(h, "dgemm.tma3d", 1) != 0;
    if (!ta && box3 && m / 16 * 16 >= BM && make_map_3d(&maps[2], A, m / 16 * 16, k, lda, BK, BM)) p.a_3d = m / 16 * 16;
    if (tb && box3 && n / 16 * 16 >= BN && make_map_3d(&maps[3], B, n / 16 * 16, k, ldb, BK, BN)) p.b_3d = n / 16 * 16;
    bool ok = ta ? make_map(&maps[0], A, k, m, lda, BK, BM) : make_map(&maps[0], A, m, k, lda, 16, BK);
    ok = ok && (!tb ? make_map(&maps[1], B, k, n, ldb, BK, BN) : make_map(&maps[1], B, n, k, ldb, 16, BK));
    if (!ok) return RTAT_B200_STATUS_NOT_SUPPORTED;  // descriptor out of range (dims >= 2^32): caller falls back
  }
  // kernel name reported through rtat_b200_last_kernel: dgemm_ws_dmma_<tile>_{tma|cpasync8}[_streamk]
  static thread_local char name[96];
  snprintf(name, sizeof name, "%s_ws_dmma_%s_%s%s", tri ? "dsyrk" : "dgemm", tile_name, use_tma ? "tma" : "cpasync8", sk_tiles ? "_streamk" : "");
  if (tri) {
    if (ta)
      return use_tma ? launch<Cfg, true, false, true, true>(h, p, grid, maps, name) : launch<Cfg, true, false, false, true>(h, p, grid, maps, name);
    return use_tma ? launch<Cfg, false, true, true, true>(h, p, grid, maps, name) : launch<Cfg, false, true, false, true>(h, p, grid, maps, name);
  }
#define RTAT_WS(TA_, TB_) \
  (use_tma ? launch<Cfg, TA_, TB_, true>(h, p, grid, maps, name) : launch<Cfg, TA_, TB_, false>(h, p, grid, maps, name))
  if (!ta && !tb) return RTAT_WS(false, false);
  if (ta && !tb) return RTAT_WS(true, false);
  if (!ta && tb) return RTAT_WS(false, true);
  return RTAT_WS(true, true);
#undef RTAT_WS
}

// tile_cfg: 0 = choose, 128 / 64 = forced, 192 = 128x64 tiles.  tri: 0 = GEMM; 1 / 2 = SYRK on the lower / upper triangle
// (m == n, B is A with the opposite op; only tiles touching the triangle are scheduled and the diagonal tiles are masked).
int dgemm_ws(rtat_b200_context *h, bool ta, bool tb, int m, int n, int k, double alpha, const double *A, int lda,
             const double *B, int ldb, double beta, double *C, int ldc, bool allow_tma, int tile_cfg, int tri) {
  // Thin remainders are peeled off.  A 128-row tile with only a few valid rows (or columns) costs a whole tile: its
  // consumers skip work at warp granularity only (they sit at the register cap), and with m % 128 <= PEEL_MAX the rows
  // all belong to the same M-slice, i.e. to half the warps on two of the four sub-partitions.  Config 3
  // (m = 4097) spent 24 of its 792 tile times on ONE row of C.  Such a strip goes to a second, small launch on 64x64
  // tiles (whose consumers skip single 8-row blocks) and the main launch sees a multiple of its tile size.
  constexpr int PEEL_MAX = 16;
  auto peeled = [&](int x, int b) { return (!tri && opt(h, "dgemm.peel", 1) != 0 && x % b > 0 && x % b <= PEEL_MAX && x >= 512) ? x - x % b : x; };
  const bool tma = allow_tma && dgemm_ws_tma_ok(A, lda, B, ldb);
  if (tile_cfg == 0) {
    // Estimated cost = padded tile area x waves.  128x128 tiles are the more efficient steady state
    // (98 % vs ~94 % of the DMMA pipe), 64x64 tiles waste less on ragged edges and keep stream-K fix-ups
    // four times smaller, which decides small problems.
    // ragged tiles are charged what their busiest sub-partition multiplies (tile_cost)
    auto padded = [&](int bm, int bn, int mm, int nn) {
      double tm = (mm + bm - 1) / bm, tn = (nn + bn - 1) / bn;
      if (tri) return tm * (tm + 1) / 2 * bm * bn;
      const bool amn = !ta;
      auto cost = [&](int rows, int cols) {
        return bn == 128 ? tile_cost<ws::Cfg128>(amn, rows, cols) : bm == 128 ? tile_cost<ws::Cfg128x64>(amn, rows, cols) : tile_cost<ws::Cfg64>(amn, rows, cols);
      };
      double tmf = mm / bm, tnf = nn / bn, sixteenths = 16.0 * tmf * tnf;
      if (nn % bn) sixteenths += tmf * cost(0, nn % bn);
      if (mm % bm) sixteenths += tnf * cost(mm % bm, 0);
      if (mm % bm && nn % bn) sixteenths += cost(mm % bm, nn % bn);
      return sixteenths / 16.0 * bm * bn;
    };
    // measured steady-state efficiencies: TMA producers 0.985 / 0.96 (128 / 64 tiles); cp.async producers 0.92 / 0.85
    // the peeled strips ride along at the 64x64 rate plus ~8 us of fixed cost per strip (as C area: t x rate / 2k)
    const double launch_area = 8e-6 * 36e12 / (2.0 * (k > 0 ? k : 1));
    auto with_strips = [&](int bm, int bn, double eff) {
      const int mm = peeled(m, bm), nn = peeled(n, bn);
      double strips = (mm != m || nn != n) ? (padded(64, 64, m - mm, n) + padded(64, 64, mm, n - nn)) / (tma ? 0.96 : 0.85) + launch_area * ((mm != m) + (nn != n)) : 0.0;
      return padded(bm, bn, mm, nn) / eff + strips;
    };
    double t128 = with_strips(128, 128, tma ? 0.985 : 0.92), t64 = padded(64, 64, m, n) / (tma ? 0.96 : 0.85);
    long long units128 = (long long)((m + 127) / 128) * ((n + 127) / 128) * ((k + 15) / 16) / (tri ? 2 : 1);
    bool small = units128 < (long long)h->sm_count * 64;  // under ~64 k-tiles per SM the fix-up chain is a visible tail
    // 128x64 tiles: the warp tile, and the steady state, of the 128x128 configuration (8192^3: 36.16 against 36.47 TF: one
    // CTA's ring feeds half the DMMAs) with twice the tiles: from a dozen waves of 128x128 tiles down they balance better
    // than they cost (4096^3 36.19 / 36.13, 3000^3 34.86 / 34.13, config 3's pad plan 1.450 / 1.477 ms; 64x64: 1.500).
    // SYRK keeps the square tiles its triangle schedule assumes.
    const double waves128 = (double)((m + 127) / 128) * ((n + 127) / 128) / h->sm_count;
    double tmid = tri ? 1e300 : with_strips(128, 64, tma ? (waves128 >= 12.0 ? 0.977 : 0.987) : 0.91);
    if (opt(h, "dgemm.mid", 1) == 0) tmid = 1e300;
    // 64x64 tiles pull twice the operand bytes through L2; when A and B fit in L2 that costs nothing and the
    // estimates are closer than their error, so a peeled 128x128 schedule has to win by 3 % (config 3 inside the planner,
    // operands hot in L2 from the pad pass: 64x64 1.453 ms, 128x128 + strip 1.470; cold: 1.527 against 1.496)
    const bool l2_resident = 8.0 * ((double)m * k + (double)k * n) < 0.95 * (double)h->l2_bytes;
    if ((peeled(m, 128) != m || peeled(n, 128) != n) && l2_resident) t128 *= 1.03;
    // few tiles and a long k (512 x 512 x 16384: 16 tiles of 128x128): the finer cut gives every CTA a shorter fix-up
    const bool few = (long long)((m + 127) / 128) * ((n + 127) / 128) <= h->sm_count / 4;
    // ... unless the 64x64 tiles would be ragged: their block-granular edge loop is the slower code (1000^3: 82 us against
    // 76 on 128x64 tiles; 1024^3: 71 against 77)
    const bool ragged64 = (m % 64 != 0 || n % 64 != 0) && tmid < 1e300 && m >= 256 && n >= 256;
    // (a peeled strip is one or two more launches: the estimates carry their fixed cost, which decides the K-panel products
    // of the host pipeline — 4097 x 384 x 512 stays on 64x64 tiles)
    if (small || few) tile_cfg = (ragged64 && tmid < 1.05 * t64) ? 192 : 64;
    else tile_cfg = (t64 < t128 && t64 < tmid) ? 64 : (tmid < t128 ? 192 : 128);
  }
  if (tile_cfg == 64)
    return run_cfg<ws::Cfg64>(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, allow_tma, "64x64x16", tri);
  const bool mid = tile_cfg == 192 && !tri;
  const int m_main = peeled(m, 128), n_main = peeled(n, mid ? 64 : 128);
  const bool peel = m_main != m || n_main != n;
  auto main_launch = [&](int mm, int nn) {
    return mid ? run_cfg<ws::Cfg128x64>(h, ta, tb, mm, nn, k, alpha, A, lda, B, ldb, beta, C, ldc, allow_tma, "128x64x16", 0)
               : run_cfg<ws::Cfg128>(h, ta, tb, mm, nn, k, alpha, A, lda, B, ldb, beta, C, ldc, allow_tma, "128x128x16", tri);
  };
  if (!peel) return main_launch(m, n);
  int st = main_launch(m_main, n_main);
  if (st) return st;
  // The strips: at most 16 rows / columns of C each.  Up to 4 wide they are a handful of GEMVs and go to the skinny kernel
  // (dgemm_skinny.cu), which reads the large operand once; wider ones stay on a 64x64-tile launch (one tile per 64
  // columns, stream-K over every CTA: 27 us for config 3's single row against 16; option dgemm.skinny = 0: always).
  const bool skinny = opt(h, "dgemm.skinny", 1) != 0;
  static thread_local char name[112];
  const bool any_skinny = skinny && ((m_main != m && m - m_main <= 4) || (n_main != n && n - n_main <= 4));
  snprintf(name, sizeof name, "%s+strips%s", h->last_kernel, any_skinny ? "_skinny" : "64");
  if (m_main != m) {  // rows [m_main, m), all columns
    const double *As = ta ? A + (size_t)m_main * lda : A + m_main;
    st = RTAT_B200_STATUS_NOT_SUPPORTED;
    if (skinny)  // S = op(A) rows, X = op(B) (k x n), Y = rows of C
      st = dgemm_skinny(h, m - m_main, n, k, alpha, As, ta ? lda : 1, ta ? 1 : lda, B, tb ? ldb : 1, tb ? 1 : ldb, beta, C + m_main, 1, ldc);
    if (st == RTAT_B200_STATUS_NOT_SUPPORTED)
      st = run_cfg<ws::Cfg64>(h, ta, tb, m - m_main, n, k, alpha, As, lda, B, ldb, beta, C + m_main, ldc, allow_tma, "64x64x16", 0);
    if (st) return st;
  }
  if (n_main != n) {  // columns [n_main, n) of the rows the main launch covered
    const double *Bs = tb ? B + n_main : B + (size_t)n_main * ldb;
    st = RTAT_B200_STATUS_NOT_SUPPORTED;
    if (skinny)  // transposed: S = op(B)^T rows (the strip's columns), X = op(A)^T (k x m_main), Y = columns of C read as rows
      st = dgemm_skinny(h, n - n_main, m_main, k, alpha, Bs, tb ? 1 : ldb, tb ? ldb : 1, A, ta ? 1 : lda, ta ? lda : 1, beta,
                        C + (size_t)n_main * ldc, ldc, 1);
    if (st == RTAT_B200_STATUS_NOT_SUPPORTED)
      st = run_cfg<ws::Cfg64>(h, ta, tb, m_main, n - n_main, k, alpha, A, lda, Bs, ldb, beta, C + (size_t)n_main * ldc, ldc, allow_tma, "64x64x16", 0);
    if (st) return st;
  }
  h->last_kernel = name;
  return RTAT_B200_STATUS_SUCCESS;
}

}  // namespace rtat_b200

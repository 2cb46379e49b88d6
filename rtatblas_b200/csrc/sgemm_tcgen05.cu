// SGEMM on 5th-generation tensor cores with FP32-faithful accuracy (3xTF32):
//     C = alpha * op(A) * op(B) + beta * C      (column-major)
//
// Replaces cublasSgemm (default math mode = true FP32) behind MatrixMult / MatrixMultAlloc for
// T = float (reference src/matrix_ops/matrixop.h:33-42).  Each FP32 operand x is split into
// big = tf32(x) and small = tf32(x - big); the product is accumulated in FP32 in tensor memory as
// small*big + big*small + big*big, which recovers FP32-level accuracy (the dropped small*small
// term and the rounding of the smalls are O(2^-22) relative).
//
// The tensor core adds into its FP32 accumulator with truncation, so a long accumulation chain at
// full magnitude drifts by ~(k/8) half-ulps (measured: 3e-3 absolute at k = 4096, 12x an FP32 FMA
// chain).  The accumulator is therefore promoted every CHUNK_KB k-blocks (256 k): the epilogue warps
// drain it from TMEM and add it, with round-to-nearest, into FP32 registers while the MMAs continue
// in the twin accumulator; inside a chunk the partial sums are small, and so are their ulps.
//
// Structure (one persistent CTA per SM, warp-specialised, everything synchronised by mbarriers):
//   warp 0        TMA producer: raw FP32 tiles of op(A) (128 x 32) and op(B) (128 x 32) into a 4-stage
//                 smem ring.  Either stored orientation is loaded in place — the transpose is folded into
//                 the descriptors, never materialised.
//   warps 4-7     split warps.  A never goes back to shared memory: thread r reads row r of the A tile
//                 (conflict-free either way: an M-contiguous tile is read one 128 B k-row per warp, a
//                 K-contiguous tile through the 128 B swizzle), splits it in registers and tcgen05.st's
//                 big and small straight into TENSOR MEMORY (lane r, 32 + 32 columns per stage), from
//                 where the MMA takes its A operand.  B is split in shared memory: big in place, small
//                 to a twin tile at the same (swizzled) offsets; fence to the async proxy.
//   warp 1        MMA issuer: one elected thread issues 12 tcgen05.mma.kind::tf32 (M128 N128 K8, A from
//                 TMEM, B from smem: K-major SWIZZLE_128B or MN-major 128B_BASE32B) per k-block into one
//                 of two ping-pong 128-column TMEM accumulators (switched every chunk); tcgen05.commit
//                 releases the stage (smem + TMEM A slots) / publishes an accumulator.
//   warps 8-11    epilogue: per chunk tcgen05.ld the finished accumulator (each warp its own 32-lane
//                 quarter) and add it into 128 FP32 registers per thread; at the end of the tile apply
//                 alpha/beta, coalesced stores to C.
// Keeping A out of the MMA's shared-memory reads matters because this kernel is shared-memory-bandwidth
// bound: per 32-k block the SM moves TMA 32 KB + split reads 32 KB + B split writes 32 KB + MMA B reads
// 48 KB = 144 KB (the all-smem version moved 224 KB and reached 43 % tensor-pipe utilisation).
// TMEM map (512 columns): [0,128) [128,256) accumulators; [256 + 64 s, +32) A_big, (+32, +64) A_small of stage s.
#include <cuda.h>
#include <algorithm>
#include "common.cuh"

namespace rtat_b200 {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
#ifndef RTAT_SGEMM_STAGES
#define RTAT_SGEMM_STAGES 4  // probe builds override this (probe/build_variants.sh)
#endif
constexpr int STAGES = RTAT_SGEMM_STAGES;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB: one operand tile
constexpr int STAGE_BYTES = 3 * TILE_BYTES;        // A raw, B raw -> big, B small
constexpr int TMEM_A0 = 256;                       // first TMEM column of the A staging slots (64 per stage)
// Warp roles: 0-3 split, 4-7 epilogue, 8 TMA producer, 9 MMA issuer.  tcgen05.ld / .st reach tensor-memory lanes
// 32 * (warp % 4) .. + 31, so both four-warp groups start at a multiple of four.  The two issuing warps come LAST on
// purpose: an SM sub-partition hands its issue slot to the eligible warp with the highest index, and the MMA issuer's
// loop — a dozen tcgen05.mma with their descriptor arithmetic per k-block, one thread's worth of instructions — is what
// paces the tensor core.  With the issuer as warp 1 (sharing a scheduler with split warp 5 and epilogue warp 9, and
// losing to both) ncu's sampling showed it busy in that loop all the time while the tensor pipe sat idle 30 % of the
// kernel: 1 130 cycles per k-block against the 768 the twelve MMAs need.
constexpr int NUM_THREADS = 320;
constexpr int SPLIT_WARP0 = 0, EPI_WARP0 = 4, TMA_WARP = 8, MMA_WARP = 9;
constexpr int TMEM_COLS = 512;   // two accumulators + four A staging slots
// Promote the accumulator every CHUNK_KB k-blocks (8 = 256 k).  The drain is not free: tcgen05.ld of 64 KB of
// accumulators per chunk competes with the MMAs for tensor memory (config 4, CTA pairs: 0.599 ms at 4, 0.553 at 8, 0.570
// at 16; 8192^3: 203 / 217 / 229 TFLOP/s), while the truncation drift inside a chunk grows only with the square root
// of its length; 8 keeps the error within 4x an FP32 FMA chain's (tests/test_gpu_parity.py, accuracy class test).
#ifndef RTAT_SGEMM_CHUNK_KB
#define RTAT_SGEMM_CHUNK_KB 8
#endif
constexpr int CHUNK_KB = RTAT_SGEMM_CHUNK_KB;
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 + 256;

struct Params {
  int m, n, k;
  float *C;
  int ldc;
  float alpha, beta;
  int tiles_m, tiles_n, num_tiles, kb;  // kb = ceil(k / BK)
  int tri;  // SSYRK: 1 = lower / 2 = upper triangle only (tiles enumerate the triangle, diagonal tiles are masked)
};

constexpr int RASTER_GROUP = 8;
// Origin of tile `tile`: grouped raster over the tile grid for GEMM; for SYRK tile t = i (i + 1) / 2 + j, i >= j.
__device__ __forceinline__ void tile_origin(const Params &p, int tile, int &m0, int &n0) {
  if (p.tri) {
    int i = int((__fsqrt_rn(8.0f * float(tile) + 1.0f) - 1.0f) * 0.5f);
    while ((long long)i * (i + 1) / 2 > tile) i--;
    while ((long long)(i + 1) * (i + 2) / 2 <= tile) i++;
    const int j = tile - int((long long)i * (i + 1) / 2);
    m0 = (p.tri == 1 ? i : j) * BM;
    n0 = (p.tri == 1 ? j : i) * BN;
  } else {
    // grouped raster: RASTER_GROUP tile rows at a time sweep all tile columns, so the CTAs that run together cover a
    // GROUP x (grid / GROUP) patch of C and share both their A row panels and their B column panels in L2 (walking
    // straight down the columns re-read all of A once per tile column: 16384^3 moved 137 GB for 1 GB of A)
    const int group_size = RASTER_GROUP * p.tiles_n, first = (tile / group_size) * RASTER_GROUP;
    const int gm = min(p.tiles_m - first, RASTER_GROUP), r = tile % group_size;
    m0 = (first + r % gm) * BM;
    n0 = (r / gm) * BN;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

// Split of one op(A) element for the tensor core: big = tf32(x) (round to nearest, ties away), small = x - big.
// RTAT_SGEMM_FAST_SPLIT = 0: both through cvt.rna.tf32.f32, which ptxas expands to VIADD + LOP3 + FSETP + SEL each (the
// NaN-preserving form) — 8 ALU instructions per element besides the FADD, ~350 issue slots per k-block for a split warp.
// = 1: big = (bits + 0x1000) & ~0x1fff (the same value for every finite x, Inf included; only NaNs whose payload has
// bits 12-22 all set come out wrong, and those are caught by small), small = x - big handed over UNROUNDED: the
// tensor core ignores the 13 low mantissa bits of a tf32 operand, i.e. truncates small (relative error 2^-10 of a value
// that is itself <= 2^-11 |x|: 2^-21 |x|, zero-mean, the size of the small * small term 3xTF32 drops anyway), and a NaN or
// Inf in x makes small = NaN, which the truncation keeps a NaN.  3 instructions per element.
// Measured (config 4 / 4096^3 TN / 8192^3 NT, CTA pairs): 0.598 -> 0.554 ms / 220 -> 243 / 215 -> 243 TFLOP/s, error
// against the exact product unchanged to three digits (probe/sgemm_pair.py); the split warps were the latency between a
// stage landing and its MMAs, and their ALU work counts against the power cap the kernel runs at.
#ifndef RTAT_SGEMM_FAST_SPLIT
#define RTAT_SGEMM_FAST_SPLIT 1
#endif
__device__ __forceinline__ void split_a(float x, uint32_t &big, uint32_t &small) {
#if RTAT_SGEMM_FAST_SPLIT
  big = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  small = __float_as_uint(x - __uint_as_float(big));
#else
  big = to_tf32(x);
  small = to_tf32(x - __uint_as_float(big));
#endif
}

// One lane of the (converged) warp, chosen by the hardware.  The issuing warps run their loops with all 32 lanes and
// only issue under this predicate: every operand is then warp-uniform for ptxas (uniform registers, no per-instruction
// election loops), which a `lane == 0` region does not give it.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- UMMA descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128 B swizzle: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
// layout_type: 2 = SWIZZLE_128B (16 B chunks; K-major tiles), 1 = SWIZZLE_128B_BASE32B (32 B chunks; the only
// swizzled layout the tensor core accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}
// instruction descriptor for kind::tf32, FP32 accumulate: c_format=F32 [4,6) | a,b format TF32=2 [7,10),[10,13) |
// a_major [15] | b_major [16] (1 = MN-major) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_instr_desc(bool a_mn_major, bool b_mn_major, int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row, one column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

// One operand's descriptor geometry.  KC: K contiguous in global => K-major tile [128 rows][32 k];
// the four K=8 steps of a k-block advance the start address by 32 B inside the swizzle atom.
// !KC: M/N contiguous => MN-major tile, four boxes [32 k][32 mn] 4 KB apart (LBO), loaded with TMA's
// 128B_ATOM_32B swizzle: rows of 128 B per k, 32 B chunks XORed with (k & 3); the swizzle atom is 4 k-rows
// (512 B = SBO), so a K=8 step is two atoms (1 KB).
template <bool KC>
__device__ __forceinline__ uint64_t operand_desc(uint32_t tile_addr, int kstep) {
  if (KC) return make_smem_desc(tile_addr + kstep * 32, 16, 1024, 2);
  return make_smem_desc(tile_addr + kstep * 1024, BK * 128, 512, 1);
}


// 32 columns of this warp's 32 tensor-memory lanes <- / -> one register per column and lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// One thread's row of a finished tile: C[row, n0 .. n0 + 128) = alpha * sum + beta * C.  8 columns at a time: with
// beta != 0 the 8 loads go out together before the first store (a load-modify-store loop would pay one DRAM latency per
// element, C aliasing itself).
__device__ __forceinline__ void store_row(const Params &p, const float (&sum)[BN], int row, int n0, float alpha, float beta) {
  if (row >= p.m) return;
#pragma unroll
  for (int j0 = 0; j0 < BN; j0 += 8) {
    float old[8];
    bool ok[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int col = n0 + j0 + j;
      ok[j] = col < p.n && !(p.tri == 1 && row < col) && !(p.tri == 2 && row > col);
      old[j] = 0.f;
    }
    if (beta != 0.f) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (ok[j]) old[j] = p.C[(size_t)(n0 + j0 + j) * p.ldc + row];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (ok[j]) {
        float v = alpha * sum[j0 + j];
        if (beta != 0.f) v += beta * old[j];
        p.C[(size_t)(n0 + j0 + j) * p.ldc + row] = v;
      }
    }
  }
}

// PRESPLIT: op(B) was split into big / small arrays in global memory by split_tf32_kernel (it is the
// operand every M-tile re-reads, so splitting it once instead of once per tile takes the B work off
// the split warps and 32 KB per k-block off the shared-memory pipe); mapB / mapBs address those arrays.
template <bool TA, bool TB, bool PRESPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapBs, const Params p) {
  constexpr bool KCA = TA, KCB = !TB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barrier map (8 B each): full[s] | conv[s] | empty[s] | tmem_full[2] | tmem_empty[2] | tmem ptr slot
  const uint32_t full0 = bar_base, conv0 = bar_base + 8 * STAGES, empty0 = bar_base + 16 * STAGES;
  const uint32_t tfull0 = bar_base + 24 * STAGES, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(conv0 + 8 * s, 4);   // one arrive per split warp
      mbar_init(empty0 + 8 * s, 1);  // tcgen05.commit
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == TMA_WARP) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m0, n0;
      tile_origin(p, tile, m0, n0);
      for (int kb = 0; kb < p.kb; kb++) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t full = full0 + 8 * stage;
          const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + TILE_BYTES;
          mbar_arrive_expect_tx(full, (PRESPLIT ? 3 : 2) * TILE_BYTES);
          const int k0 = kb * BK;
          if (KCA) tma_load_2d(sA, &mapA, k0, m0, full);
          else
            tma_load_2d(sA, &mapA, m0, k0, full);  // one box: 32 k-rows of 128 consecutive m (512 contiguous bytes each)
          if (KCB) tma_load_2d(sB, &mapB, k0, n0, full);
          else
            for (int b = 0; b < BN / 32; b++) tma_load_2d(sB + b * (BK * 128), &mapB, n0 + 32 * b, k0, full);
          if (PRESPLIT) {
            const uint32_t sBs = sB + TILE_BYTES;
            if (KCB) tma_load_2d(sBs, &mapBs, k0, n0, full);
            else
              for (int b = 0; b < BN / 32; b++) tma_load_2d(sBs + b * (BK * 128), &mapBs, n0 + 32 * b, k0, full);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    constexpr uint32_t idesc = make_instr_desc(false, !KCB, BM, BN);  // A comes from TMEM: always K-major
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.kb; kb++) {
        const bool chunk_start = (kb % CHUNK_KB) == 0, chunk_end = ((kb + 1) % CHUNK_KB) == 0 || kb + 1 == p.kb;
        if (chunk_start) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
        mbar_wait(conv0 + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (elect_one()) {
          const uint32_t d = tmem_base + acc * BN;
          const uint32_t a_big_t = tmem_base + TMEM_A0 + 64 * stage, a_small_t = a_big_t + 32;
          const uint32_t sB = smem_base + stage * STAGE_BYTES + TILE_BYTES, sBs = sB + TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ks++) {
            const uint64_t b_big = operand_desc<KCB>(sB, ks), b_small = operand_desc<KCB>(sBs, ks);
            umma_tf32_ts(d, a_small_t + 8 * ks, b_big, idesc, !(chunk_start && ks == 0));
            umma_tf32_ts(d, a_big_t + 8 * ks, b_small, idesc, 1);
            umma_tf32_ts(d, a_big_t + 8 * ks, b_big, idesc, 1);
          }
          umma_commit(empty0 + 8 * stage);  // frees the smem stage and its TMEM A slots once these MMAs have read them
          if (chunk_end) umma_commit(tfull0 + 8 * acc);  // this chunk's accumulator is complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + 4) {
    // ===================== split warps =====================
    const int t = threadIdx.x - SPLIT_WARP0 * 32;  // 0..127 == row of the A tile == TMEM lane
    const int quarter = warp - SPLIT_WARP0;        // == warp % 4: the TMEM lanes this warp may write
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.kb; kb++) {
        mbar_wait(full0 + 8 * stage, phase);
        const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + TILE_BYTES, sBs = sB + TILE_BYTES;
        // ---- A: row t of the raw tile -> registers -> big / small -> tensor memory
        uint32_t big[32], small[32];
        if (KCA) {
          // K-major tile [128 rows][32 k], 16 B chunks XORed with (row & 7): 8 vector loads, conflict-free
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sA + t * 128 + ((c ^ (t & 7)) << 4)));
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              split_a(x[e], big[4 * c + e], small[4 * c + e]);
            }
          }
        } else {
          // M-major tile: [32 k][128 m] unswizzled (one TMA box); a warp reads 128 contiguous bytes of one k-row per load
#pragma unroll
          for (int kk = 0; kk < 32; kk++) {
            float x;
            asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(x) : "r"(sA + kk * (BM * 4) + t * 4));
            split_a(x, big[kk], small[kk]);
          }
        }
        const uint32_t ta = lane_base + TMEM_A0 + 64 * stage;
        tmem_st32(ta, big);
        tmem_st32(ta + 32, small);
        // ---- B: raw -> big in place, small to the twin tile (position-preserving, so layout-agnostic)
#pragma unroll
        for (int c = 0; c < (PRESPLIT ? 0 : TILE_BYTES / 16 / 128); c++) {
          const uint32_t off = (c * 128 + t) * 16;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sB + off));
          uint32_t b0 = to_tf32(v.x), b1 = to_tf32(v.y), b2 = to_tf32(v.z), b3 = to_tf32(v.w);
          uint32_t s0 = to_tf32(v.x - __uint_as_float(b0)), s1 = to_tf32(v.y - __uint_as_float(b1));
          uint32_t s2 = to_tf32(v.z - __uint_as_float(b2)), s3 = to_tf32(v.w - __uint_as_float(b3));
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(sB + off), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(sBs + off), "r"(s0), "r"(s1), "r"(s2), "r"(s3) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");          // TMEM stores complete
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        if (!PRESPLIT) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy smem writes -> visible to UMMA
        __syncwarp();
        if (lane == 0) mbar_arrive(conv0 + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ===================== epilogue =====================
    const int quarter = warp - EPI_WARP0;  // == warp % 4: the TMEM lanes this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const float alpha = p.alpha, beta = p.beta;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m0, n0;
      tile_origin(p, tile, m0, n0);
      float sum[BN];  // this thread's row of the tile: FP32 registers, round-to-nearest adds
#pragma unroll
      for (int j = 0; j < BN; j++) sum[j] = 0.f;
      const int chunks = (p.kb + CHUNK_KB - 1) / CHUNK_KB;
      for (int ch = 0; ch < chunks; ch++) {
        mbar_wait(tfull0 + 8 * acc, acc_phase);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_base + acc * BN + c0, r);
#pragma unroll
          for (int j = 0; j < 32; j++) sum[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      store_row(p, sum, m0 + quarter * 32 + lane, n0, alpha, beta);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// ======================================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of one cluster — the two SMs of a TPC — work on one 256 x 128 tile.
// Each CTA owns 128 rows of op(A) (streamed once, split in registers, staged in ITS tensor memory exactly as above) and
// only HALF of the tile's op(B) columns (64 of 128): the pair's MMA (M256 N128 K8, issued by the leader CTA alone) reads
// each half from the shared memory of the CTA that loaded it, so per k-block a CTA takes 16 KB of pre-split B from L2
// instead of 32 KB and its tensor core reads 24 KB of B from shared memory instead of 48 KB.  That is the traffic that
// bounds the single-CTA kernel (shared memory moves 112 KB per k-block at 128 B/clk = 875 clk, against 768 clk of MMA
// issue); the pair moves 72 KB (562 clk).  Config 4 (131072 x 128 x 4096) is the shape this is for: n = 128 is ONE tile
// column, so all 1024 M-tiles re-read the same op(B).
//
// Synchronisation (barriers sit at the same shared-memory offsets in both CTAs):
//   full[s]    local   TMA bytes of this CTA's stage s (A raw 16 KB + B big / small halves 8 KB each)
//   conv[s]    LEADER  8 arrivals: the 4 split warps of BOTH CTAs (remote arrive from the peer) — A is in tensor memory
//                      and, because a split warp waits on its full[s] first, that CTA's half of B has landed
//   empty[s]   both    tcgen05.commit.cta_group::2 multicast: the pair's MMAs have read stage s in both CTAs
//   tfull[a]   both    same multicast commit: accumulator a holds a finished chunk (each CTA drains its own 128 rows)
//   tempty[a]  LEADER  8 arrivals: the 4 epilogue warps of both CTAs have drained accumulator a
namespace pair {
constexpr int BN_HALF = BN / 2;
constexpr int B_HALF_BYTES = BN_HALF * BK * 4;             // 8 KB
constexpr int STAGE_BYTES = TILE_BYTES + 2 * B_HALF_BYTES;  // A raw + B big half + B small half = 32 KB
#ifndef RTAT_SGEMM_PAIR_STAGES
#define RTAT_SGEMM_PAIR_STAGES 6
#endif
// The shared-memory ring (what TMA fills) is deeper than the tensor-memory ring of split A tiles: the first has to cover
// the L2 latency of a machine whose L2 -> SM path runs at its limit (config 4 moves 42 B/clk into every SM; four 32 KB
// stages covered barely 3000 cycles, and the pair ran latency-bound at 56 % tensor-pipe activity), the second only the
// split -> MMA hand-off.  k-block i of a CTA uses shared-memory stage i % STAGES and tensor-memory slot i % TSLOTS.
constexpr int STAGES = RTAT_SGEMM_PAIR_STAGES;
constexpr int TSLOTS = 4;  // 64 tensor-memory columns each (32 big + 32 small), next to the two 128-column accumulators
static_assert(STAGES >= TSLOTS && STAGES <= 7, "7 x 32 KB is what fits in 227 KB of shared memory");
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier in the leader's shared memory.  Default semantics (release at CTA scope), as CUTLASS' ClusterBarrier
// does: what the arrival publishes lives in tensor memory / TMA-written shared memory and is ordered by the tcgen05 fences
// around it; asking for .release.cluster here makes ptxas put MEMBAR.ALL.GPU in front of every arrival and CCTL.IVALL behind
// every cluster-scope wait, which halved the kernel's speed (1.23 ms against 0.67 ms on config 4).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// A (256 x 8, 128 rows in each CTA's tensor memory) x B (8 x 128, 64 columns in each CTA's shared memory)
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrives (once) on the barrier at this shared-memory offset in BOTH CTAs when all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <bool TA, bool TB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
sgemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                          const __grid_constant__ CUtensorMap mapBs, const Params p) {
  constexpr bool KCA = TA, KCB = !TB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full0 = bar_base, conv0 = bar_base + 8 * STAGES, empty0 = bar_base + 16 * STAGES;
  const uint32_t tfull0 = bar_base + 24 * STAGES, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x / 2, num_pairs = gridDim.x / 2;
  // pair tiles: 256 rows x 128 columns, numbered down the columns of the tile grid; this CTA's half starts at row 128 * rank.
  // SSYRK (p.tri): only the pair tiles that touch the triangle, pair row by pair row — columns 0 .. 2 I + 1 of pair row I
  // for the lower triangle, 2 I .. tiles_n - 1 for the upper one.  The half tile on the far side of the diagonal (one per
  // pair row) is multiplied and masked away by the epilogue: 1 / (tiles_m + 1) of the work.
  const int pairs_m = (p.tiles_m + 1) / 2;
  auto tri_count = [&](int I) { return p.tri == 1 ? min(2 * I + 2, p.tiles_n) : p.tiles_n - 2 * I; };
  int num_pair_tiles = pairs_m * p.tiles_n;
  if (p.tri) {
    num_pair_tiles = 0;
    for (int I = 0; I < pairs_m; I++) num_pair_tiles += tri_count(I);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(conv0 + 8 * s, 8);   // four split warps in each CTA of the pair (only the leader's copy is used)
      mbar_init(empty0 + 8 * s, 1);  // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull0 + 8 * a, 1);   // multicast tcgen05.commit
      mbar_init(tempty0 + 8 * a, 8);  // four epilogue warps in each CTA (leader's copy)
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == MMA_WARP) {  // the same warp in both CTAs: the allocation is made in both tensor memories at the same columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync_all();  // both CTAs' barriers are initialised before anyone arrives on a peer's
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));

  auto origin = [&](int pt, int &m0, int &n0) {
    if (p.tri) {
      int I = 0;
      for (int c = tri_count(0); pt >= c; c = tri_count(I)) { pt -= c; I++; }  // <= pairs_m steps, once per 256 x 128 tile
      m0 = I * (2 * BM) + (int)rank * BM;
      n0 = (p.tri == 1 ? pt : 2 * I + pt) * BN;
      return;
    }
    // grouped raster over pair rows (see tile_origin)
    const int group_size = RASTER_GROUP * p.tiles_n, first = (pt / group_size) * RASTER_GROUP;
    const int gm = min(pairs_m - first, RASTER_GROUP), r = pt % group_size;
    m0 = (first + r % gm) * (2 * BM) + (int)rank * BM;
    n0 = (r / gm) * BN;
  };

  if (warp == TMA_WARP) {
    // ===================== TMA producer (each CTA loads its own rows of A and its own half of B) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
      int m0, n0;
      origin(pt, m0, n0);
      const int nb0 = n0 + (int)rank * BN_HALF;
      for (int kb = 0; kb < p.kb; kb++) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t full = full0 + 8 * stage;
          const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + TILE_BYTES, sBs = sB + B_HALF_BYTES;
          mbar_arrive_expect_tx(full, STAGE_BYTES);
          const int k0 = kb * BK;
          if (KCA) tma_load_2d(sA, &mapA, k0, m0, full);
          else
            tma_load_2d(sA, &mapA, m0, k0, full);  // one box: 32 k-rows of 128 consecutive m (512 contiguous bytes each)
          if (KCB) {
            tma_load_2d(sB, &mapB, k0, nb0, full);
            tma_load_2d(sBs, &mapBs, k0, nb0, full);
          } else {
            for (int b = 0; b < BN_HALF / 32; b++) {
              tma_load_2d(sB + b * (BK * 128), &mapB, nb0 + 32 * b, k0, full);
              tma_load_2d(sBs + b * (BK * 128), &mapBs, nb0 + 32 * b, k0, full);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer: the leader CTA's elected lane drives both tensor cores =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_instr_desc(false, !KCB, 2 * BM, BN);  // M = 256 across the pair; A from TMEM: K-major
      int stage = 0, acc = 0, tslot = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
        for (int kb = 0; kb < p.kb; kb++) {
          const bool chunk_start = (kb % CHUNK_KB) == 0, chunk_end = ((kb + 1) % CHUNK_KB) == 0 || kb + 1 == p.kb;
          if (chunk_start) mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
          mbar_wait(conv0 + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (elect_one()) {
            const uint32_t d = tmem_base + acc * BN;
            const uint32_t a_big_t = tmem_base + TMEM_A0 + 64 * tslot, a_small_t = a_big_t + 32;
            const uint32_t sB = smem_base + stage * STAGE_BYTES + TILE_BYTES, sBs = sB + B_HALF_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ks++) {
              const uint64_t b_big = operand_desc<KCB>(sB, ks), b_small = operand_desc<KCB>(sBs, ks);
              umma_tf32_ts_pair(d, a_small_t + 8 * ks, b_big, idesc, !(chunk_start && ks == 0));
              umma_tf32_ts_pair(d, a_big_t + 8 * ks, b_small, idesc, 1);
              umma_tf32_ts_pair(d, a_big_t + 8 * ks, b_big, idesc, 1);
            }
            umma_commit_pair(empty0 + 8 * stage);  // frees stage s (shared memory + tensor-memory A slot) in BOTH CTAs
            if (chunk_end) umma_commit_pair(tfull0 + 8 * acc);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++tslot == TSLOTS) tslot = 0;
          if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + 4) {
    // ===================== split warps (per CTA, as in the single-CTA kernel; B is pre-split) =====================
    const int t = threadIdx.x - SPLIT_WARP0 * 32;
    const int quarter = warp - SPLIT_WARP0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t conv_leader = map_to_cta(conv0, 0);
    int stage = 0, tslot = 0;
    uint32_t phase = 0;
    // tensor-memory slot i % TSLOTS may be rewritten once the MMAs of k-block i - TSLOTS are complete, which is what the
    // commit on that k-block's `empty` barrier reports: (old_stage, old_phase) trail (stage, phase) by TSLOTS k-blocks.
    // The barrier cannot run a whole phase ahead of this wait: its next completion needs k-block i - TSLOTS + STAGES > i.
    int it = 0, old_stage = 0;
    uint32_t old_phase = 0;
    for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
      for (int kb = 0; kb < p.kb; kb++, it++) {
        mbar_wait(full0 + 8 * stage, phase);
        const uint32_t sA = smem_base + stage * STAGE_BYTES;
        uint32_t big[32], small[32];
        if (KCA) {
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sA + t * 128 + ((c ^ (t & 7)) << 4)));
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              split_a(x[e], big[4 * c + e], small[4 * c + e]);
            }
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < 32; kk++) {
            float x;
            asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(x) : "r"(sA + kk * (BM * 4) + t * 4));
            split_a(x, big[kk], small[kk]);
          }
        }
        if (STAGES > TSLOTS && it >= TSLOTS) {
          mbar_wait(empty0 + 8 * old_stage, old_phase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (++old_stage == STAGES) { old_stage = 0; old_phase ^= 1; }
        }
        const uint32_t ta = lane_base + TMEM_A0 + 64 * tslot;
        tmem_st32(ta, big);
        tmem_st32(ta + 32, small);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(conv_leader + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++tslot == TSLOTS) tslot = 0;
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ===================== epilogue: each CTA drains and stores its own 128 rows =====================
    const int quarter = warp - EPI_WARP0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float alpha = p.alpha, beta = p.beta;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tempty_leader = map_to_cta(tempty0, 0);
    for (int pt = pair_id; pt < num_pair_tiles; pt += num_pairs) {
      int m0, n0;
      origin(pt, m0, n0);
      float sum[BN];
#pragma unroll
      for (int j = 0; j < BN; j++) sum[j] = 0.f;
      const int chunks = (p.kb + CHUNK_KB - 1) / CHUNK_KB;
      for (int ch = 0; ch < chunks; ch++) {
        mbar_wait(tfull0 + 8 * acc, acc_phase);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_base + acc * BN + c0, r);
#pragma unroll
          for (int j = 0; j < 32; j++) sum[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      store_row(p, sum, m0 + quarter * 32 + lane, n0, alpha, beta);
    }
  }

  // neither CTA may leave (or give its tensor memory back) while the pair's MMAs / multicast commits can still touch it
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync_all();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}
}  // namespace pair

typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap *map, const float *base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems, uint32_t box0, uint32_t box1,
                     CUtensorMapSwizzle swizzle) {
  auto encode = reinterpret_cast<encode_tiled_t>(tensor_map_encoder_raw());
  if (!encode) return false;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld_elems * sizeof(float)};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// B (rows x cols, leading dimension ldb) -> big = tf32(b), small = tf32(b - big), both with leading dimension ldo
__global__ void split_tf32_kernel(const float *__restrict__ B, int ldb, int rows, int cols, float *__restrict__ big, float *__restrict__ small,
                                  int ldo) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = int(i % rows), c = int(i / rows);
    const float x = B[(size_t)c * ldb + r];
    const uint32_t b = to_tf32(x);
    big[(size_t)c * ldo + r] = __uint_as_float(b);
    small[(size_t)c * ldo + r] = __uint_as_float(to_tf32(x - __uint_as_float(b)));
  }
}

template <bool TA, bool TB, bool PRESPLIT>
static int launch(rtat_b200_context *h, const Params &p, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mbs) {
  auto kern = sgemm_tcgen05_kernel<TA, TB, PRESPLIT>;
  static std::atomic<uint64_t> configured{0};  // per template instantiation
  if (int st = configure_dynamic_smem(kern, SMEM_BYTES, h->device, configured)) return st;
  int grid = p.num_tiles < h->sm_count ? p.num_tiles : h->sm_count;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, h->stream>>>(ma, mb, mbs, p);
  if (p.tri) return check_launch(h, PRESPLIT ? "ssyrk_tcgen05_3xtf32_128x128x32_presplitB" : "ssyrk_tcgen05_3xtf32_128x128x32", PRESPLIT ? 2 : 1);
  return check_launch(h, PRESPLIT ? "sgemm_tcgen05_3xtf32_128x128x32_presplitB" : "sgemm_tcgen05_3xtf32_128x128x32", PRESPLIT ? 2 : 1);
}

// CTA-pair launch: clusters of two CTAs (the kernel carries __cluster_dims__(2, 1, 1)), one pair per two SMs
template <bool TA, bool TB>
static int launch_pair(rtat_b200_context *h, const Params &p, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mbs) {
  auto kern = pair::sgemm_tcgen05_pair_kernel<TA, TB>;
  static std::atomic<uint64_t> configured{0};  // per template instantiation
  if (int st = configure_dynamic_smem(kern, pair::SMEM_BYTES, h->device, configured)) return st;
  const int pairs_m = (p.tiles_m + 1) / 2, max_pairs = h->sm_count / 2;
  long long pair_tiles = (long long)pairs_m * p.tiles_n;
  if (p.tri) {  // the pair tiles that touch the triangle (same count the kernel derives)
    pair_tiles = 0;
    for (int I = 0; I < pairs_m; I++) pair_tiles += p.tri == 1 ? std::min(2 * I + 2, p.tiles_n) : p.tiles_n - 2 * I;
  }
  const int grid = 2 * (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
  kern<<<grid, NUM_THREADS, pair::SMEM_BYTES, h->stream>>>(ma, mb, mbs, p);
  return check_launch(h, p.tri ? "ssyrk_tcgen05_3xtf32_2cta_256x128x32_presplitB" : "sgemm_tcgen05_3xtf32_2cta_256x128x32_presplitB", 2);
}

}  // namespace tc

// op(B) is pre-split into the handle's arena when both halves fit in PRESPLIT_LIMIT bytes and B is re-read
// by at least two M-tiles; B then needs no alignment at all (the pre-pass reads any ldb).
constexpr size_t PRESPLIT_LIMIT = size_t(4) << 30;  // 16384^2 floats of B -> 2 GB of arena; beyond that B is split per tile in shared memory

// b_addressable: TMA can read B as stored, so the pre-split is a choice.  It is not made for small products, where the extra
// launch costs more than splitting op(B) again in every M-tile (256^3: 18.5 -> 14.5 us, 512^3: 22.6 -> 20.1, 768^3: 26.7 ->
// 24.7; from 1024^3 or k = 4096 on the pre-split is level or ahead: 512 x 512 x 4096 73.9 against 89.9 us; probe/perf_sgemm_small2.py)
static bool presplit_wanted(const rtat_b200_context *h, int m, int n, int k, bool b_addressable) {
  if (opt(h, "sgemm.presplit", 1) == 0) return false;
  if (b_addressable && opt(h, "sgemm.presplit", 1) != 2 && (double)m * n * k <= 768.0 * 768.0 * 768.0) return false;
  // either stored orientation of B: rows rounded up to 4 floats
  const size_t bytes_n = 2 * sizeof(float) * ((size_t)(k + 3) / 4 * 4) * (size_t)n;  // stored k x n
  const size_t bytes_t = 2 * sizeof(float) * ((size_t)(n + 3) / 4 * 4) * (size_t)k;  // stored n x k
  return m > tc::BM && bytes_n <= PRESPLIT_LIMIT && bytes_t <= PRESPLIT_LIMIT;
}

// An op(A) TMA cannot address is first copied into a 16 B-aligned buffer owned by the handle (one streaming pass over A,
// geam.cu: ~7 TB/s) — for anything but a very thin product that is a few per cent of the multiply, where the FP32 SIMT
// fallback runs at a fifth of this kernel's rate (4097^3: 3.4 ms against ~0.65).  Not for more than PAD_A_LIMIT bytes of A
// (the copy doubles A's footprint) and not when the product does too little work per element of A to pay for the pass.
constexpr size_t PAD_A_LIMIT = size_t(1) << 30;
static bool pad_a_wanted(const rtat_b200_context *h, bool ta, int m, int n, int k) {
  if (opt(h, "sgemm.pad_a", 1) == 0) return false;
  const size_t rows = ta ? k : m, cols = ta ? m : k;
  return n >= 64 && sizeof(float) * ((rows + 3) / 4 * 4) * cols <= PAD_A_LIMIT;
}

bool sgemm_tcgen05_supported(const rtat_b200_context *h, int transa, int, int m, int n, int k, const float *A, int lda, const float *B, int ldb,
                             const float *, int) {
  auto ok = [](const void *ptr, int ld) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && ld % 4 == 0; };
  if (!(m > 0 && n > 0 && k > 0)) return false;
  if (!ok(A, lda) && !pad_a_wanted(h, transa == RTAT_B200_OP_T, m, n, k)) return false;
  return ok(B, ldb) || presplit_wanted(h, m, n, k, false);
}

int sgemm_tcgen05(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha, const float *A, int lda, const float *B,
                  int ldb, float beta, float *C, int ldc, int tri) {
  using namespace tc;
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  Params p{m, n, k, C, ldc, alpha, beta, 0, 0, 0, 0, tri};
  p.tiles_m = (m + BM - 1) / BM;
  p.tiles_n = (n + BN - 1) / BN;
  long long tiles = (long long)p.tiles_m * p.tiles_n;
  if (tri) {
    if (m != n) return RTAT_B200_STATUS_INVALID_VALUE;
    tiles = (long long)p.tiles_m * (p.tiles_m + 1) / 2;
  }
  if (tiles > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
  p.num_tiles = (int)tiles;
  p.kb = (k + BK - 1) / BK;
  bool padded_a = false;
  if (reinterpret_cast<uintptr_t>(A) % 16 != 0 || lda % 4 != 0) {
    if (!pad_a_wanted(h, ta, m, n, k)) return RTAT_B200_STATUS_NOT_SUPPORTED;
    const int rows = ta ? k : m, cols = ta ? m : k, ldp = (rows + 3) / 4 * 4;
    const size_t need = sizeof(float) * (size_t)ldp * cols;
    if (need > h->sgemm_a_pad_bytes) {
      if (h->sgemm_a_pad) {
        cudaStreamSynchronize(h->stream);  // an earlier product may still be reading the old copy
        cudaFree(h->sgemm_a_pad);
        h->sgemm_a_pad = nullptr;
        h->sgemm_a_pad_bytes = 0;
      }
      const size_t want = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
      if (cudaMalloc(&h->sgemm_a_pad, want) != cudaSuccess) {
        (void)cudaGetLastError();
        return RTAT_B200_STATUS_ALLOC_FAILED;
      }
      h->sgemm_a_pad_bytes = want;
    }
    h->scratch_in_flight = true;
    float *Ap = static_cast<float *>(h->sgemm_a_pad);
    if (int st = geam<float>(h, RTAT_B200_OP_N, RTAT_B200_OP_N, rows, cols, 1.f, A, lda, 0.f, Ap, ldp, Ap, ldp)) return st;
    A = Ap;
    lda = ldp;
    padded_a = true;
  }
  // (SYRK keeps the pre-split, and with it the CTA pairs, at every size: its timings were taken that way)
  const bool presplit = presplit_wanted(h, m, n, k, !tri && reinterpret_cast<uintptr_t>(B) % 16 == 0 && ldb % 4 == 0);
  const float *Bbig = B, *Bsmall = B;
  int ldb_eff = ldb;
  if (presplit) {
    // stored B is (tb ? n x k : k x n); split copies keep that shape with a TMA-friendly leading dimension
    const int rows = tb ? n : k, cols = tb ? k : n, ldo = (rows + 3) / 4 * 4;
    const size_t half = sizeof(float) * (size_t)ldo * cols;
    const size_t half_al = (half + 1023) & ~size_t(1023);
    int st = ensure_arena(h, 2 * half_al);
    if (st) return st;
    float *big = static_cast<float *>(h->arena), *small = reinterpret_cast<float *>(static_cast<char *>(h->arena) + half_al);
    long long total = (long long)rows * cols;
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 8);
    split_tf32_kernel<<<blocks, 256, 0, h->stream>>>(B, ldb, rows, cols, big, small, ldo);
    Bbig = big;
    Bsmall = small;
    ldb_eff = ldo;
  }
  CUtensorMap ma, mb, mbs;
  // K-contiguous operand: dims {K, MN}, box {32, 128}; M/N-contiguous: dims {MN, K}, box {32, 32}
  const CUtensorMapSwizzle kmaj = CU_TENSOR_MAP_SWIZZLE_128B, mnmaj = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  // CTA pairs (cta_group::2) whenever there are at least two M-tiles to pair and op(B) is pre-split: each CTA then loads
  // only its half of the tile's B columns (box {32, 64})
  const bool use_pair = presplit && p.tiles_m >= 2 && opt(h, "sgemm.pair", 1) != 0;
  const int bn_box = use_pair ? BN / 2 : BN;
  // A is only ever read by the split warps (it reaches the MMA through TMEM): an M-contiguous tile needs no swizzle
  bool ok = ta ? make_map(&ma, A, k, m, lda, BK, BM, kmaj) : make_map(&ma, A, m, k, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE);
  ok = ok && (!tb ? make_map(&mb, Bbig, k, n, ldb_eff, BK, bn_box, kmaj) : make_map(&mb, Bbig, n, k, ldb_eff, 32, BK, mnmaj));
  ok = ok && (!tb ? make_map(&mbs, Bsmall, k, n, ldb_eff, BK, bn_box, kmaj) : make_map(&mbs, Bsmall, n, k, ldb_eff, 32, BK, mnmaj));
  if (!ok) return RTAT_B200_STATUS_NOT_SUPPORTED;
  int st;
#define RTAT_TC(TA_, TB_) (use_pair ? launch_pair<TA_, TB_>(h, p, ma, mb, mbs) : presplit ? launch<TA_, TB_, true>(h, p, ma, mb, mbs) : launch<TA_, TB_, false>(h, p, ma, mb, mbs))
  if (!ta && !tb) st = RTAT_TC(false, false);
  else if (ta && !tb) st = RTAT_TC(true, false);
  else if (!ta && tb) st = RTAT_TC(false, true);
  else st = RTAT_TC(true, true);
  if (st == RTAT_B200_STATUS_SUCCESS && padded_a) {  // reported name: <kernel>_padA
    static thread_local char name[96];
    snprintf(name, sizeof name, "%s_padA", h->last_kernel);
    h->last_kernel = name;
  }
  return st;
#undef RTAT_TC
}

}  // namespace rtat_b200

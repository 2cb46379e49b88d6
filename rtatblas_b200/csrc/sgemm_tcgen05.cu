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
// chain).  The accumulator is therefore promoted every CHUNK_KB k-blocks (128 k): the epilogue warps
// drain it from TMEM and add it, with round-to-nearest, into FP32 registers while the MMAs continue
// in the twin accumulator; inside a chunk the partial sums are small, and so are their ulps.
//
// Where the operands live
//   op(B)  is the operand every M-tile re-reads, so it is split ONCE per call by split_tf32_kernel into
//          big / small arrays in the handle's arena (N is processed in slabs when B is too large for
//          that); the GEMM kernel TMA-loads both halves straight into UMMA layout (K-major:
//          SWIZZLE_128B; N-major: TMA 128B_ATOM_32B + UMMA SWIZZLE_128B_BASE32B, the only MN-major
//          layout tf32 accepts).  The pre-pass reads any ldb, so B needs no alignment.
//   op(A)  is streamed once.  Its raw FP32 tiles land in a 4 x 16 KB shared-memory ring; thread r of
//          the split warps reads row r of the tile (conflict-free either way: an M-contiguous tile is
//          read one 128 B k-row per warp, a K-contiguous tile through the 128 B swizzle), splits it in
//          registers and tcgen05.st's big and small into TENSOR MEMORY (lane r, 32 + 32 columns per
//          slot).  The smem slot is released as soon as the row is in registers — not when the MMA is
//          done — which is what keeps enough HBM requests in flight for the tall-skinny config 4.
//          The MMA takes A from TMEM (TS mode), so A never costs shared-memory read bandwidth again.
//
// One persistent CTA per SM, warp-specialised, everything synchronised by mbarriers:
//   warp 0 / 2    TMA producers (one thread each) for the A ring and the B ring
//   warp 1        MMA issuer: 12 tcgen05.mma.kind::tf32 (M128 N128 K8) per 32-k block into one of two
//                 ping-pong 128-column TMEM accumulators; tcgen05.commit frees the TMEM A slot and the B stage
//   warps 4-7     split warps (A: smem -> registers -> TMEM)
//   warps 8-11    epilogue: per chunk tcgen05.ld the finished accumulator (each warp its own 32-lane
//                 quarter) into 128 FP32 registers per thread; at tile end alpha/beta and coalesced stores
// TMEM map (512 columns): [0,128) [128,256) accumulators; [256 + 64 t, +32) A_big, (+32, +64) A_small of slot t.
#include <cuda.h>
#include <algorithm>
#include "common.cuh"

namespace rtat_b200 {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int A_STAGES = 4, B_STAGES = 5, T_SLOTS = 4;  // B is the latency-critical ring: a stage is only refilled after its MMAs complete
constexpr int TILE_BYTES = BM * BK * 4;   // 16 KB: one operand tile
constexpr int B_STAGE_BYTES = 2 * TILE_BYTES;  // big, small
constexpr int TMEM_A0 = 256;              // first TMEM column of the A slots (64 per slot)
constexpr int NUM_THREADS = 384;
constexpr int SPLIT_WARP0 = 4, EPI_WARP0 = 8;
constexpr int TMEM_COLS = 512;            // two accumulators + four A slots
constexpr int CHUNK_KB = 4;               // promote the accumulator every 4 k-blocks = 128 k
constexpr size_t SMEM_BYTES = size_t(A_STAGES) * TILE_BYTES + size_t(B_STAGES) * B_STAGE_BYTES + 1024 + 256;

struct Params {
  int m, n, k;
  float *C;
  int ldc;
  float alpha, beta;
  int tiles_m, tiles_n, num_tiles, kb;  // kb = ceil(k / BK)
  int debug;  // profiling experiments only (sgemm.debug): 1 = issue big*big only, 4 = epilogue skips the chunk drains
};

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

// ---- UMMA descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout [61,64): 2 = SWIZZLE_128B (16 B chunks; K-major tiles), 1 = SWIZZLE_128B_BASE32B (32 B chunks; the only
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

// B tile descriptor for K=8 step `kstep` of a 32-k block.  KC: K contiguous => K-major tile [128 rows][32 k]; a
// step advances the start address by 32 B inside the swizzle atom.  !KC: N contiguous => four boxes [32 k][32 n]
// 4 KB apart (LBO), rows of 128 B per k with 32 B chunks XORed by (k & 3); the swizzle atom is 4 k-rows
// (512 B = SBO), so a K=8 step is two atoms (1 KB).
template <bool KC>
__device__ __forceinline__ uint64_t operand_desc(uint32_t tile_addr, int kstep) {
  if (KC) return make_smem_desc(tile_addr + kstep * 32, 16, 1024, 2);
  return make_smem_desc(tile_addr + kstep * 1024, BK * 128, 512, 1);
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapBs, const Params p) {
  constexpr bool KCA = TA, KCB = !TB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = smem_base, b_ring = smem_base + A_STAGES * TILE_BYTES, bar_base = b_ring + B_STAGES * B_STAGE_BYTES;
  // barriers (8 B each)
  const uint32_t a_full = bar_base, a_empty = a_full + 8 * A_STAGES;                // A smem ring: TMA -> split warps -> TMA
  const uint32_t b_full = a_empty + 8 * A_STAGES, b_empty = b_full + 8 * B_STAGES;  // B smem ring: TMA -> MMA -> TMA
  const uint32_t t_ready = b_empty + 8 * B_STAGES, t_free = t_ready + 8 * T_SLOTS;  // TMEM A slots: split -> MMA -> split
  const uint32_t acc_full = t_free + 8 * T_SLOTS, acc_empty = acc_full + 16, tmem_slot = acc_empty + 16;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_STAGES; s++) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 4); }
    for (int s = 0; s < B_STAGES; s++) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
    for (int s = 0; s < T_SLOTS; s++) { mbar_init(t_ready + 8 * s, 4); mbar_init(t_free + 8 * s, 1); }
    for (int a = 0; a < 2; a++) { mbar_init(acc_full + 8 * a, 1); mbar_init(acc_empty + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer: A ring =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile % p.tiles_m) * BM;
        for (int kb = 0; kb < p.kb; kb++) {
          mbar_wait(a_empty + 8 * stage, phase ^ 1);
          const uint32_t full = a_full + 8 * stage, sA = a_ring + stage * TILE_BYTES;
          mbar_arrive_expect_tx(full, TILE_BYTES);
          if (KCA) tma_load_2d(sA, &mapA, kb * BK, m0, full);
          else
            for (int b = 0; b < BM / 32; b++) tma_load_2d(sA + b * (BK * 128), &mapA, m0 + 32 * b, kb * BK, full);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer: B ring (pre-split big / small) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n0 = (tile / p.tiles_m) * BN;
        for (int kb = 0; kb < p.kb; kb++) {
          mbar_wait(b_empty + 8 * stage, phase ^ 1);
          const uint32_t full = b_full + 8 * stage, sB = b_ring + stage * B_STAGE_BYTES, sBs = sB + TILE_BYTES;
          mbar_arrive_expect_tx(full, B_STAGE_BYTES);
          if (KCB) {
            tma_load_2d(sB, &mapB, kb * BK, n0, full);
            tma_load_2d(sBs, &mapBs, kb * BK, n0, full);
          } else {
            for (int b = 0; b < BN / 32; b++) {
              tma_load_2d(sB + b * (BK * 128), &mapB, n0 + 32 * b, kb * BK, full);
              tma_load_2d(sBs + b * (BK * 128), &mapBs, n0 + 32 * b, kb * BK, full);
            }
          }
          if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_instr_desc(false, !KCB, BM, BN);  // A comes from TMEM: always K-major
      int bs = 0, ts = 0, acc = 0;
      uint32_t bphase = 0, tphase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.kb; kb++) {
          const bool chunk_start = (kb % CHUNK_KB) == 0, chunk_end = ((kb + 1) % CHUNK_KB) == 0 || kb + 1 == p.kb;
          if (chunk_start) mbar_wait(acc_empty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
          mbar_wait(t_ready + 8 * ts, tphase);
          mbar_wait(b_full + 8 * bs, bphase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t d = tmem_base + acc * BN;
          const uint32_t a_big_t = tmem_base + TMEM_A0 + 64 * ts, a_small_t = a_big_t + 32;
          const uint32_t sB = b_ring + bs * B_STAGE_BYTES, sBs = sB + TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ks++) {
            const uint64_t b_big = operand_desc<KCB>(sB, ks), b_small = operand_desc<KCB>(sBs, ks);
            if (!(p.debug & 1)) {
              umma_tf32_ts(d, a_small_t + 8 * ks, b_big, idesc, !(chunk_start && ks == 0));
              umma_tf32_ts(d, a_big_t + 8 * ks, b_small, idesc, 1);
              umma_tf32_ts(d, a_big_t + 8 * ks, b_big, idesc, 1);
            } else {
              umma_tf32_ts(d, a_big_t + 8 * ks, b_big, idesc, !(chunk_start && ks == 0));
            }
          }
          umma_commit(t_free + 8 * ts);   // TMEM A slot reusable once these MMAs have read it
          umma_commit(b_empty + 8 * bs);  // B stage reusable
          if (++ts == T_SLOTS) { ts = 0; tphase ^= 1; }
          if (++bs == B_STAGES) { bs = 0; bphase ^= 1; }
          if (chunk_end) {
            umma_commit(acc_full + 8 * acc);  // this chunk's accumulator is complete
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + 4) {
    // ===================== split warps: A smem -> registers -> big / small -> tensor memory =====================
    const int t = threadIdx.x - SPLIT_WARP0 * 32;  // 0..127 == row of the A tile == TMEM lane
    const int quarter = warp - SPLIT_WARP0;        // == warp % 4: the TMEM lanes this warp may write
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int as = 0, ts = 0;
    uint32_t aphase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.kb; kb++) {
        mbar_wait(a_full + 8 * as, aphase);
        const uint32_t sA = a_ring + as * TILE_BYTES;
        float x[32];
        if (KCA) {
          // K-major tile [128 rows][32 k], 16 B chunks XORed with (row & 7): 8 vector loads, conflict-free
#pragma unroll
          for (int c = 0; c < 8; c++)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                         : "=f"(x[4 * c]), "=f"(x[4 * c + 1]), "=f"(x[4 * c + 2]), "=f"(x[4 * c + 3])
                         : "r"(sA + t * 128 + ((c ^ (t & 7)) << 4)));
        } else {
          // M-major tile: box (t / 32) holds [32 k][32 m] unswizzled; a warp reads one 128 B k-row per load
#pragma unroll
          for (int kk = 0; kk < 32; kk++)
            asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(x[kk]) : "r"(sA + (t >> 5) * (BK * 128) + kk * 128 + (t & 31) * 4));
        }
        uint32_t big[32], small[32];
#pragma unroll
        for (int kk = 0; kk < 32; kk++) {
          big[kk] = to_tf32(x[kk]);
          small[kk] = to_tf32(x[kk] - __uint_as_float(big[kk]));
        }
        // the row is in registers: hand the smem slot back to the TMA producer right away
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty + 8 * as);
        if (++as == A_STAGES) { as = 0; aphase ^= 1; }

        mbar_wait(t_free + 8 * ts, tphase ^ 1);  // the MMAs that read this TMEM slot last time have completed
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t ta = lane_base + TMEM_A0 + 64 * ts;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(ta),
            "r"(big[0]), "r"(big[1]), "r"(big[2]), "r"(big[3]), "r"(big[4]), "r"(big[5]), "r"(big[6]), "r"(big[7]), "r"(big[8]), "r"(big[9]),
            "r"(big[10]), "r"(big[11]), "r"(big[12]), "r"(big[13]), "r"(big[14]), "r"(big[15]), "r"(big[16]), "r"(big[17]), "r"(big[18]),
            "r"(big[19]), "r"(big[20]), "r"(big[21]), "r"(big[22]), "r"(big[23]), "r"(big[24]), "r"(big[25]), "r"(big[26]), "r"(big[27]),
            "r"(big[28]), "r"(big[29]), "r"(big[30]), "r"(big[31])
            : "memory");
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(ta + 32),
            "r"(small[0]), "r"(small[1]), "r"(small[2]), "r"(small[3]), "r"(small[4]), "r"(small[5]), "r"(small[6]), "r"(small[7]), "r"(small[8]),
            "r"(small[9]), "r"(small[10]), "r"(small[11]), "r"(small[12]), "r"(small[13]), "r"(small[14]), "r"(small[15]), "r"(small[16]),
            "r"(small[17]), "r"(small[18]), "r"(small[19]), "r"(small[20]), "r"(small[21]), "r"(small[22]), "r"(small[23]), "r"(small[24]),
            "r"(small[25]), "r"(small[26]), "r"(small[27]), "r"(small[28]), "r"(small[29]), "r"(small[30]), "r"(small[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");  // TMEM stores complete
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_ready + 8 * ts);
        if (++ts == T_SLOTS) { ts = 0; tphase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quarter = warp - EPI_WARP0;  // == warp % 4: the TMEM lanes this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const float alpha = p.alpha, beta = p.beta;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto tmem_ld32 = [](uint32_t taddr, uint32_t (&r)[32]) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m0 = (tile % p.tiles_m) * BM, n0 = (tile / p.tiles_m) * BN;
      float sum[BN];  // this thread's row of the tile: FP32 registers, round-to-nearest adds
#pragma unroll
      for (int j = 0; j < BN; j++) sum[j] = 0.f;
      const int chunks = (p.kb + CHUNK_KB - 1) / CHUNK_KB;
      for (int ch = 0; ch < chunks; ch++) {
        mbar_wait(acc_full + 8 * acc, acc_phase);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (!(p.debug & 4) || ch + 1 == chunks) {
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + acc * BN + c0, r);
#pragma unroll
            for (int j = 0; j < 32; j++) sum[c0 + j] += __uint_as_float(r[j]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int row = m0 + quarter * 32 + lane;
      if (row < p.m) {
#pragma unroll
        for (int j = 0; j < BN; j++) {
          const int col = n0 + j;
          if (col < p.n) {
            float *dst = p.C + (size_t)col * p.ldc + row;
            float v = alpha * sum[j];
            if (beta != 0.f) v += beta * *dst;
            *dst = v;
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
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

template <bool TA, bool TB>
static int launch(rtat_b200_context *h, const Params &p, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mbs) {
  auto kern = sgemm_tcgen05_kernel<TA, TB>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_INTERNAL_ERROR;
    }
    configured = true;
  }
  int grid = p.num_tiles < h->sm_count ? p.num_tiles : h->sm_count;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, h->stream>>>(ma, mb, mbs, p);
  return check_launch(h, "sgemm_tcgen05_3xtf32_128x128x32", 2);
}

}  // namespace tc

// A must be TMA-addressable (16 B aligned base, ld % 4 == 0); B may be anything (the pre-pass reads it).
bool sgemm_tcgen05_supported(const rtat_b200_context *, int, int, int m, int n, int k, const float *A, int lda, const float *, int,
                             const float *, int) {
  return m > 0 && n > 0 && k > 0 && reinterpret_cast<uintptr_t>(A) % 16 == 0 && lda % 4 == 0;
}

int sgemm_tcgen05(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha, const float *A, int lda, const float *B,
                  int ldb, float beta, float *C, int ldc) {
  using namespace tc;
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  // N is processed in slabs so that the split copy of op(B)'s slab (2 x k x slab floats) fits the arena budget
  const size_t budget = size_t(opt(h, "sgemm.presplit_mb", 512)) << 20;
  const size_t k4 = (size_t)(k + 3) / 4 * 4;
  long long slab = (long long)(budget / (2 * sizeof(float) * k4));
  slab = slab / BN * BN;
  if (slab < BN) slab = BN;
  if (slab > n) slab = n;
  for (int j0 = 0; j0 < n; j0 += (int)slab) {
    const int ns = (int)std::min<long long>(slab, n - j0);
    // stored shape of the slab of B: not transposed -> k x ns at B + j0*ldb; transposed -> ns x k at B + j0
    const float *Bs = tb ? B + j0 : B + (size_t)j0 * ldb;
    const int rows = tb ? ns : k, cols = tb ? k : ns, ldo = (rows + 3) / 4 * 4;
    const size_t half = ((sizeof(float) * (size_t)ldo * cols) + 1023) & ~size_t(1023);
    int st = ensure_arena(h, 2 * half);
    if (st) return st;
    float *big = static_cast<float *>(h->arena), *small = reinterpret_cast<float *>(static_cast<char *>(h->arena) + half);
    const long long total = (long long)rows * cols;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 8);
    split_tf32_kernel<<<blocks, 256, 0, h->stream>>>(Bs, ldb, rows, cols, big, small, ldo);

    Params p{m, ns, k, C + (size_t)j0 * ldc, ldc, alpha, beta, 0, 0, 0, 0, opt(h, "sgemm.debug")};
    p.tiles_m = (m + BM - 1) / BM;
    p.tiles_n = (ns + BN - 1) / BN;
    long long tiles = (long long)p.tiles_m * p.tiles_n;
    if (tiles > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
    p.num_tiles = (int)tiles;
    p.kb = (k + BK - 1) / BK;
    CUtensorMap ma, mb, mbs;
    // K-contiguous operand: dims {K, MN}, box {32, 128}; M/N-contiguous: dims {MN, K}, box {32, 32}.
    // A is only ever read by the split warps (it reaches the MMA through TMEM): an M-contiguous tile needs no swizzle.
    const CUtensorMapSwizzle kmaj = CU_TENSOR_MAP_SWIZZLE_128B, mnmaj = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    bool ok = ta ? make_map(&ma, A, k, m, lda, BK, BM, kmaj) : make_map(&ma, A, m, k, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    ok = ok && (!tb ? make_map(&mb, big, k, ns, ldo, BK, BN, kmaj) : make_map(&mb, big, ns, k, ldo, 32, BK, mnmaj));
    ok = ok && (!tb ? make_map(&mbs, small, k, ns, ldo, BK, BN, kmaj) : make_map(&mbs, small, ns, k, ldo, 32, BK, mnmaj));
    if (!ok) return RTAT_B200_STATUS_NOT_SUPPORTED;
    if (!ta && !tb) st = launch<false, false>(h, p, ma, mb, mbs);
    else if (ta && !tb) st = launch<true, false>(h, p, ma, mb, mbs);
    else if (!ta && tb) st = launch<false, true>(h, p, ma, mb, mbs);
    else st = launch<true, true>(h, p, ma, mb, mbs);
    if (st) return st;
  }
  return RTAT_B200_STATUS_SUCCESS;
}

}  // namespace rtat_b200

// Transposing GEAM through TMA:  C = alpha * A^T + beta * B   (B not transposed; C may alias B)
//
// This is MatrixMove with transpose (reference src/matrix_ops/matrixop.h:307-344) and the transposed
// MatrixAccumulate (:256-305) for operands TMA can address (16 B aligned base, ld * sizeof(T) a
// multiple of 16).  HBM-bound: algorithmic traffic is one read of A (and of B when beta != 0) and
// one write of C.
//
// One persistent CTA per SM walks 64x64-element tiles:
//   * a load thread TMA-loads the tile of A (boxes of 128 B x 64 rows, 128 B swizzle) into a 3-stage
//     ring, and — when beta != 0 — the tile of B straight into the output staging buffer;
//   * 8 warps transpose shared -> shared: LDS.128 along A's contiguous dimension (8 lanes = 8 rows =
//     8 different swizzled 16 B chunks: conflict-free), STS along C's contiguous dimension (a warp
//     writes whole 128 B rows: conflict-free), applying alpha/beta on the way;
//   * a store thread TMA-stores the staged tile (bulk tensor store clips ragged edges, loads
//     zero-fill them, so no edge code exists).
// All global traffic is full-line TMA; the SM's LSU only ever touches shared memory.
//
// A_CPASYNC variant: A has an odd leading dimension or an 8 B / 4 B aligned base (config 3: ld 4097, 3001), which TMA cannot
// address, but C (a padded scratch matrix in the pad plans) and B can.  Four producer warps then fill the same swizzled
// ring with element-sized cp.async copies (coalesced along A's contiguous dimension, zero-filled outside A) that complete
// on the same mbarriers; transposers and the TMA store path are unchanged.
#include <cuda.h>
#include "common.cuh"

namespace rtat_b200 {
namespace gt {

constexpr int TILE = 64;
constexpr int IN_STAGES = 3, OUT_STAGES = 3;
constexpr int THREADS = 320;  // 8 transpose warps + a TMA load warp + a TMA store warp
constexpr int PRODUCER_THREADS = 128, THREADS_CPASYNC = THREADS + PRODUCER_THREADS;  // + 4 cp.async producer warps

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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(dst), "l"(src), "n"(BYTES), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}

template <typename T>
__device__ __forceinline__ T combine(T alpha, T a, T beta, T b, bool use_b);
template <>
__device__ __forceinline__ double combine<double>(double alpha, double a, double beta, double b, bool use_b) {
  double pa = __dmul_rn(alpha, a);
  return use_b ? __fma_rn(beta, b, pa) : pa;  // cuBLAS geam rounding, pinned by tests/golden
}
template <>
__device__ __forceinline__ float combine<float>(float alpha, float a, float beta, float b, bool use_b) {
  float pa = __fmul_rn(alpha, a);
  return use_b ? __fmaf_rn(beta, b, pa) : pa;
}

struct Params {
  int m, n;  // C is m x n; A is n x m
  int tiles_m, tiles_n;
  double alpha, beta;  // converted to T in the kernel
  const void *A;       // A_CPASYNC only
  int lda;
  // The bulk tensor store clips at 16 B granularity (measured: with m = 517 doubles it also wrote row 517), so mapC covers
  // only the m_store = m rounded down to a 16 B multiple leading rows and the transposers store the last m - m_store
  // (< 16 B worth of) rows themselves.
  void *C;
  int ldc, m_store;
};

// mapA: dims {n (contiguous), m}, box {ROW, 64}.  mapB / mapC: dims {m (contiguous), n}, box {ROW, 64}.
// ROW = 128 B worth of elements.
template <typename T, bool USE_B, bool A_CPASYNC>
__global__ void __launch_bounds__(A_CPASYNC ? THREADS_CPASYNC : THREADS, 1)
geam_transpose_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                          const __grid_constant__ CUtensorMap mapC, const Params p) {
  constexpr int ROW = 128 / sizeof(T);        // elements per 128 B smem row (16 doubles / 32 floats)
  constexpr int EPC = 16 / sizeof(T);         // elements per 16 B chunk
  constexpr int BOXES = TILE / ROW;           // boxes per tile side (4 / 2)
  constexpr int BOX_BYTES = TILE * 128;       // one box: 64 rows of 128 B
  constexpr int TILE_BYTES = BOXES * BOX_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t in0 = smem_base, out0 = smem_base + IN_STAGES * TILE_BYTES, bar0 = out0 + OUT_STAGES * TILE_BYTES;
  // barriers: in_full[IN] | in_empty[IN] | out_full[OUT] (B tile landed) | out_ready[OUT] (transposed) | out_free[OUT]
  const uint32_t in_full = bar0, in_empty = bar0 + 8 * IN_STAGES, b_full = in_empty + 8 * IN_STAGES, out_ready = b_full + 8 * OUT_STAGES,
                 out_free = out_ready + 8 * OUT_STAGES;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const T alpha = (T)p.alpha, beta = (T)p.beta;

  if (threadIdx.x == 0) {
    for (int s = 0; s < IN_STAGES; s++) {
      mbar_init(in_full + 8 * s, A_CPASYNC ? PRODUCER_THREADS : 1);
      mbar_init(in_empty + 8 * s, 8);
    }
    for (int s = 0; s < OUT_STAGES; s++) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(out_ready + 8 * s, 8);
      mbar_init(out_free + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const long long num_tiles = (long long)p.tiles_m * p.tiles_n;

  if (warp == 8) {
    // ===================== TMA load warp: runs IN_STAGES tiles ahead of the transposers =====================
    if (lane == 0) {
      int is = 0, os = 0;
      uint32_t iphase = 0, ophase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int i0 = int(tile % p.tiles_m) * TILE, j0 = int(tile / p.tiles_m) * TILE;
        if (!A_CPASYNC) {
          mbar_wait(in_empty + 8 * is, iphase ^ 1);
          mbar_arrive_expect_tx(in_full + 8 * is, TILE_BYTES);
          for (int b = 0; b < BOXES; b++) tma_load_2d(in0 + is * TILE_BYTES + b * BOX_BYTES, &mapA, j0 + b * ROW, i0, in_full + 8 * is);
        }
        if (USE_B) {
          mbar_wait(out_free + 8 * os, ophase ^ 1);  // the store that last used this staging buffer has read it
          mbar_arrive_expect_tx(b_full + 8 * os, TILE_BYTES);
          for (int b = 0; b < BOXES; b++) tma_load_2d(out0 + os * TILE_BYTES + b * BOX_BYTES, &mapB, i0 + b * ROW, j0, b_full + 8 * os);
        }
        if (++is == IN_STAGES) { is = 0; iphase ^= 1; }
        if (++os == OUT_STAGES) { os = 0; ophase ^= 1; }
      }
    }
    return;
  }
  if (warp == 9) {
    // ===================== TMA store warp =====================
    if (lane == 0) {
      int os = 0;
      uint32_t ophase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int i0 = int(tile % p.tiles_m) * TILE, j0 = int(tile / p.tiles_m) * TILE;
        mbar_wait(out_ready + 8 * os, ophase);
        for (int b = 0; b < BOXES; b++) tma_store_2d(&mapC, i0 + b * ROW, j0, out0 + os * TILE_BYTES + b * BOX_BYTES);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // smem of this tile has been read
        mbar_arrive(out_free + 8 * os);
        if (++os == OUT_STAGES) { os = 0; ophase ^= 1; }
      }
      asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // all stores complete before the CTA exits
    }
    return;
  }

  if (A_CPASYNC && warp >= 10) {
    // ===================== cp.async producers for A (4 warps) =====================
    // thread pt owns column jj = pt % 64 of the tile (A's contiguous dimension: a warp copies 32 consecutive elements of
    // one row) and rows pt / 64 + 2 r; the swizzled destination is what the TMA box layout would have produced
    const int pt = threadIdx.x - 10 * 32, jj = pt % TILE, ii0 = pt / TILE;
    const T *Ab = static_cast<const T *>(p.A);
    const uint32_t dst_col = (jj / ROW) * BOX_BYTES + (jj % EPC) * sizeof(T);
    const uint32_t chunk = (jj % ROW) / EPC;
    int is = 0;
    uint32_t iphase = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int i0 = int(tile % p.tiles_m) * TILE, j0 = int(tile / p.tiles_m) * TILE;
      mbar_wait(in_empty + 8 * is, iphase ^ 1);
      const uint32_t in = in0 + is * TILE_BYTES + dst_col;
      const bool jok = j0 + jj < p.n;
      const T *src = Ab + (size_t)(i0 + ii0) * p.lda + (j0 + jj);
#pragma unroll 8
      for (int r = 0; r < TILE / 2; r++) {
        const int i = ii0 + 2 * r;
        const bool ok = jok && i0 + i < p.m;
        cp_async_zfill<sizeof(T)>(in + i * 128 + ((chunk ^ (i & 7)) << 4), ok ? src : Ab, ok ? (int)sizeof(T) : 0);
        src += 2 * (size_t)p.lda;
      }
      cp_async_mbar_arrive_noinc(in_full + 8 * is);
      if (++is == IN_STAGES) { is = 0; iphase ^= 1; }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    return;
  }

  // ===================== transpose warps =====================
  int is = 0, os = 0;
  uint32_t iphase = 0, ophase = 0;
  const int t = threadIdx.x;  // 0..255
  const int i = t % TILE;     // row of the input tile (= index along C's contiguous dimension)
  const int c_first = t / TILE;  // 0..3
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int gi = int(tile % p.tiles_m) * TILE + i, j0 = int(tile / p.tiles_m) * TILE;
    const bool tail_row = gi >= p.m_store && gi < p.m;  // a row the TMA store cannot clip exactly
    mbar_wait(in_full + 8 * is, iphase);
    if (USE_B) mbar_wait(b_full + 8 * os, ophase);
    else mbar_wait(out_free + 8 * os, ophase ^ 1);
    const uint32_t in = in0 + is * TILE_BYTES, out = out0 + os * TILE_BYTES;
    // output position of row i: box i / ROW, chunk (i % ROW) / EPC, element i % EPC
    const uint32_t out_i = (i / ROW) * BOX_BYTES + (i % EPC) * sizeof(T);
    const uint32_t out_chunk = (i % ROW) / EPC;
#pragma unroll
    for (int c = c_first; c < TILE / EPC; c += 4) {
      const int j_base = c * EPC;                 // first of the EPC consecutive j this thread moves
      const int bj = j_base / ROW, cj = (j_base % ROW) / EPC;
      const uint32_t src = in + bj * BOX_BYTES + i * 128 + ((cj ^ (i & 7)) << 4);
      T v[EPC];
      if (sizeof(T) == 8) {
        double2 d;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(d.x), "=d"(d.y) : "r"(src));
        v[0] = (T)d.x;
        v[1] = (T)d.y;
      } else {
        float4 f;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(src));
        v[0] = (T)f.x;
        v[1] = (T)f.y;
        v[2 % EPC] = (T)f.z;
        v[3 % EPC] = (T)f.w;
      }
#pragma unroll
      for (int e = 0; e < EPC; e++) {
        const int j = j_base + e;
        const uint32_t dst = out + out_i + j * 128 + ((out_chunk ^ (j & 7)) << 4);
        T b = T(0);
        if (USE_B) {
          if (sizeof(T) == 8) { double tmp; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(tmp) : "r"(dst)); b = (T)tmp; }
          else { float tmp; asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(tmp) : "r"(dst)); b = (T)tmp; }
        }
        const T r = combine<T>(alpha, v[e], beta, b, USE_B);
        if (sizeof(T) == 8) asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(dst), "d"((double)r) : "memory");
        else asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(dst), "f"((float)r) : "memory");
        if (tail_row && j0 + j < p.n) static_cast<T *>(p.C)[(size_t)(j0 + j) * p.ldc + gi] = r;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // staged tile -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(in_empty + 8 * is);
      mbar_arrive(out_ready + 8 * os);
    }
    if (++is == IN_STAGES) { is = 0; iphase ^= 1; }
    if (++os == OUT_STAGES) { os = 0; ophase ^= 1; }
  }
}

typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
static bool make_map(CUtensorMap *map, const T *base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems) {
  auto encode = reinterpret_cast<encode_tiled_t>(tensor_map_encoder_raw());
  if (!encode) return false;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld_elems * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(T)), TILE};
  cuuint32_t estr[2] = {1, 1};
  return encode(map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<T *>(base), dims, strides,
                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gt

template <typename T>
static bool tma_addressable(const T *ptr, int ld) {
  return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && ((size_t)ld * sizeof(T)) % 16 == 0;
}


// A may be anything (cp.async producers take over when TMA cannot address it); C and B must be TMA-addressable.
template <typename T>
bool geam_transpose_tma_supported(int m, int n, const T *A, int lda, const T *B, int ldb, const T *C, int ldc, bool use_b) {
  (void)A; (void)lda;
  if (!tma_addressable(C, ldc) || (use_b && !tma_addressable(B, ldb))) return false;
  return (long long)m * n >= 256 * 256;  // below that a persistent TMA pipeline is all prologue
}

// C (m x n) = alpha * A^T + beta * B, A is n x m.  Returns NOT_SUPPORTED when the maps cannot be built.
template <typename T>
int geam_transpose_tma(rtat_b200_context *h, int m, int n, T alpha, const T *A, int lda, T beta, const T *B, int ldb, T *C, int ldc) {
  using namespace gt;
  const bool use_b = beta != T(0);
  const bool a_tma = tma_addressable(A, lda) && opt(h, "geam.force_cpasync") == 0;
  CUtensorMap ma, mb, mc;
  const int m_store = m / (int)(16 / sizeof(T)) * (int)(16 / sizeof(T));
  bool ok = m_store > 0 && make_map<T>(&mc, C, m_store, n, ldc);
  if (a_tma) ok = ok && make_map<T>(&ma, A, n, m, lda);
  else ma = mc;
  if (use_b) ok = ok && make_map<T>(&mb, B, m, n, ldb);
  else mb = mc;
  if (!ok) return RTAT_B200_STATUS_NOT_SUPPORTED;
  Params p{m, n, (m + TILE - 1) / TILE, (n + TILE - 1) / TILE, (double)alpha, (double)beta, A, lda, C, ldc, m_store};
  constexpr int tile_bytes = TILE * TILE * sizeof(T);
  const size_t smem = (size_t)(IN_STAGES + OUT_STAGES) * tile_bytes + 1024 + 256;
  long long tiles = (long long)p.tiles_m * p.tiles_n;
  int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
  // one mask per instantiation (the two kernels share a function-pointer type, so a generic lambda would share one)
  static std::atomic<uint64_t> configured[4];
  auto launch = [&](auto kern, std::atomic<uint64_t> &done, int threads, const char *name) {
    if (int st = configure_dynamic_smem(kern, smem, h->device, done)) return st;
    kern<<<grid, threads, smem, h->stream>>>(ma, mb, mc, p);
    return check_launch(h, name);
  };
  if (!a_tma)
    return use_b ? launch(geam_transpose_tma_kernel<T, true, true>, configured[3], THREADS_CPASYNC, "geam_transpose_cpasync_tma_accumulate")
                 : launch(geam_transpose_tma_kernel<T, false, true>, configured[2], THREADS_CPASYNC, "geam_transpose_cpasync_tma");
  return use_b ? launch(geam_transpose_tma_kernel<T, true, false>, configured[1], THREADS, "geam_transpose_tma_accumulate")
               : launch(geam_transpose_tma_kernel<T, false, false>, configured[0], THREADS, "geam_transpose_tma");
}

template bool geam_transpose_tma_supported<double>(int, int, const double *, int, const double *, int, const double *, int, bool);
template bool geam_transpose_tma_supported<float>(int, int, const float *, int, const float *, int, const float *, int, bool);
template int geam_transpose_tma<double>(rtat_b200_context *, int, int, double, const double *, int, double, const double *, int, double *, int);
template int geam_transpose_tma<float>(rtat_b200_context *, int, int, float, const float *, int, float, const float *, int, float *, int);

}  // namespace rtat_b200

// DGEMM on FP64 tensor tiles:  C = alpha * op(A) * op(B) + beta * C   (column-major)
//
// Replaces cublasDgemm behind the reference's MatrixMult / MatrixMultAlloc nodes
// (reference src/matrix_ops/matrixop.h:347-440 via gpuTgemm :15-46).
//
// tcgen05 has no FP64 kind; on sm_100a the FP64 tensor path is DMMA.8x8x4 (every mma.sync f64 shape
// lowers to it, see profiles/r01_probe_fp64_rates.txt: 37.0 TFLOP/s issue peak, DFMA only ~33).
// The kernel is therefore a warp-level DMMA GEMM:
//   * CTA tile BM x BN x BK, warps own WM x WN sub-tiles made of 8x8 DMMA blocks;
//   * a STAGES-deep cp.async ring feeds shared memory; both operand orientations are read in
//     place (the "transpose folded into the load"): an operand whose K index is contiguous in
//     global memory is staged as [mn][BK+4], one whose M/N index is contiguous as [k][BMN+4];
//     the +4 pitch makes every DMMA fragment read (8 rows x 4 k, 64-bit) bank-conflict free;
//   * 16-byte cp.async when the operand is 16 B aligned with even ld, otherwise 8-byte cp.async
//     (odd leading dimensions such as 4097 or 3001 need no pad copy: the "fused pad" path);
//   * epilogue applies alpha/beta from registers; beta == 0 never reads C.
#include "common.cuh"

namespace rtat_b200 {

struct DgemmParams {
  int m, n, k;
  const double *A;
  int lda;
  const double *B;
  int ldb;
  double *C;
  int ldc;
  double alpha, beta;
  int tiles_m, tiles_n;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// One operand tile: BMN (m or n extent) x BK, staged from global X.
//   KC = true : K is contiguous in global, element (mn,k) at X[k + mn*ld]; smem [mn][BK+4]
//   KC = false: MN is contiguous,          element (mn,k) at X[mn + k*ld]; smem [k][BMN+4]
template <int BMN, int BK, bool KC>
struct OperandTile {
  static constexpr int PITCH = KC ? BK + 4 : BMN + 4;
  static constexpr int ROWS = KC ? BMN : BK;
  static constexpr int ELEMS = ROWS * PITCH;
  static constexpr int CHUNKS_PER_ROW = (KC ? BK : BMN) / 2;

  template <int NT, bool ALIGNED>
  __device__ static __forceinline__ void load(double *s, const double *__restrict__ X, int ld, int mn0,
                                              int k0, int MN, int K, int tid) {
#pragma unroll
    for (int idx = tid; idx < ROWS * CHUNKS_PER_ROW; idx += NT) {
      int r = idx / CHUNKS_PER_ROW, c = idx % CHUNKS_PER_ROW;
      int mn = KC ? mn0 + r : mn0 + 2 * c;
      int k = KC ? k0 + 2 * c : k0 + r;
      // number of valid elements in this 2-element chunk along the contiguous dimension
      int valid = KC ? ((mn < MN) ? min(max(K - k, 0), 2) : 0) : ((k < K) ? min(max(MN - mn, 0), 2) : 0);
      const double *src = valid ? (KC ? X + (size_t)mn * ld + k : X + (size_t)k * ld + mn) : X;
      double *dst = s + r * PITCH + 2 * c;
      if (ALIGNED) {
        cp_async16(dst, src, valid * 8);
      } else {
        cp_async8(dst, src, valid >= 1 ? 8 : 0);
        cp_async8(dst + 1, valid >= 2 ? src + 1 : X, valid >= 2 ? 8 : 0);
      }
    }
  }

  // fragment element for DMMA block row/col `idx8` (0..7 within block at mn_off), k index kk
  __device__ static __forceinline__ int frag_offset(int mn, int kk) { return KC ? mn * PITCH + kk : kk * PITCH + mn; }
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, bool ALIGNED>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
dgemm_dmma_kernel(const DgemmParams p) {
  constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, NT = WARPS_M * WARPS_N * 32;
  using TileA = OperandTile<BM, BK, TA>;    // op(A)(m,k): A^T stored => k contiguous
  using TileB = OperandTile<BN, BK, !TB>;   // op(B)(k,n): B stored k x n => k contiguous when not transposed
  constexpr int STAGE_ELEMS = TileA::ELEMS + TileB::ELEMS;
  constexpr int MI = WM / 8, NI = WN / 8, KS = BK / 4;
  extern __shared__ __align__(16) double smem[];

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int q = lane % 4, g = lane / 4;
  const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;

  // grouped rasterisation: GROUP consecutive tile rows share B column strips in L2
  constexpr int GROUP = 8;
  int tile = blockIdx.x;
  int group_size = GROUP * p.tiles_n;
  int group_id = tile / group_size;
  int first_m = group_id * GROUP;
  int gm = min(p.tiles_m - first_m, GROUP);
  int tm = first_m + (tile % group_size) % gm;
  int tn = (tile % group_size) / gm;
  const int m0 = tm * BM, n0 = tn * BN;

  const int KT = (p.k + BK - 1) / BK;

  auto load_stage = [&](int slot, int kt) {
    double *sA = smem + slot * STAGE_ELEMS;
    double *sB = sA + TileA::ELEMS;
    TileA::template load<NT, ALIGNED>(sA, p.A, p.lda, m0, kt * BK, p.m, p.k, tid);
    TileB::template load<NT, ALIGNED>(sB, p.B, p.ldb, n0, kt * BK, p.n, p.k, tid);
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  // per-thread fragment base offsets (everything else is an immediate)
  const int a_base = TileA::frag_offset(wm0 + g, q);
  const int b_base = TileB::frag_offset(wn0 + g, q);

  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double *sA = smem + (kt % STAGES) * STAGE_ELEMS + a_base;
    const double *sB = smem + (kt % STAGES) * STAGE_ELEMS + TileA::ELEMS + b_base;
    double fa[2][MI], fb[2][NI];
#pragma unroll
    for (int i = 0; i < MI; i++) fa[0][i] = sA[TileA::frag_offset(8 * i, 0)];
#pragma unroll
    for (int j = 0; j < NI; j++) fb[0][j] = sB[TileB::frag_offset(8 * j, 0)];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int cur = ks & 1, nxt = cur ^ 1;
      if (ks + 1 < KS) {
#pragma unroll
        for (int i = 0; i < MI; i++) fa[nxt][i] = sA[TileA::frag_offset(8 * i, 4 * (ks + 1))];
#pragma unroll
        for (int j = 0; j < NI; j++) fb[nxt][j] = sB[TileB::frag_offset(8 * j, 4 * (ks + 1))];
      }
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma_8x8x4(acc[i][j], fa[cur][i], fb[cur][j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread (q,g) of block (i,j) holds C[m0+wm0+8i+g][n0+wn0+8j+2q+{0,1}]
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int j = 0; j < NI; j++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int col = n0 + wn0 + 8 * j + 2 * q + e;
      if (col >= p.n) continue;
      double *ccol = p.C + (size_t)col * p.ldc;
#pragma unroll
      for (int i = 0; i < MI; i++) {
        int row = m0 + wm0 + 8 * i + g;
        if (row < p.m) {
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * ccol[row];
          ccol[row] = v;
        }
      }
    }
  }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, bool ALIGNED>
static int launch_cfg(rtat_b200_context *h, DgemmParams &p, const char *name) {
  using TileA = OperandTile<BM, BK, TA>;
  using TileB = OperandTile<BN, BK, !TB>;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr size_t smem = sizeof(double) * STAGES * (TileA::ELEMS + TileB::ELEMS);
  auto kern = dgemm_dmma_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB, ALIGNED>;
  static std::atomic<uint64_t> configured{0};  // per template instantiation
  if (int st = configure_dynamic_smem(kern, smem, h->device, configured)) return st;
  p.tiles_m = (p.m + BM - 1) / BM;
  p.tiles_n = (p.n + BN - 1) / BN;
  long long tiles = (long long)p.tiles_m * p.tiles_n;
  if (tiles > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
  kern<<<(unsigned)tiles, NT, smem, h->stream>>>(p);
  return check_launch(h, name);
}

template <bool TA, bool TB, bool ALIGNED>
static int launch_ops(rtat_b200_context *h, DgemmParams &p, int variant) {
  // variant 1: 128x128 tiles (large problems); variant 2: 64x64 tiles (fill the machine on small ones)
  if (variant == 2)
    return launch_cfg<64, 64, 16, 32, 32, 4, TA, TB, ALIGNED>(h, p, ALIGNED ? "dgemm_dmma_64x64x16_cpasync16" : "dgemm_dmma_64x64x16_cpasync8");
  return launch_cfg<128, 128, 16, 64, 32, 4, TA, TB, ALIGNED>(h, p, ALIGNED ? "dgemm_dmma_128x128x16_cpasync16" : "dgemm_dmma_128x128x16_cpasync8");
}

int dgemm_ws(rtat_b200_context *h, bool ta, bool tb, int m, int n, int k, double alpha, const double *A, int lda,
             const double *B, int ldb, double beta, double *C, int ldc, bool allow_tma, int tile_cfg, int tri = 0);

int dgemm_skinny(rtat_b200_context *h, int r, int N, int K, double alpha, const double *S, long long ss_i, long long ss_k, const double *X,
                 long long xs_k, long long xs_j, double beta, double *Y, long long ys_i, long long ys_j);

int dgemm(rtat_b200_context *h, int transa, int transb, int m, int n, int k, double alpha,
          const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc) {
  if (k == 0 || alpha == 0.0) {
    // C = beta * C  (BLAS semantics: A and B are not referenced)
    if (beta == 0.0) {
      if (cudaMemset2DAsync(C, sizeof(double) * ldc, 0, sizeof(double) * m, n, h->stream) != cudaSuccess)
        return RTAT_B200_STATUS_EXECUTION_FAILED;
      h->last_kernel = "memset2d";
      return RTAT_B200_STATUS_SUCCESS;
    }
    if (beta == 1.0) return RTAT_B200_STATUS_SUCCESS;
    return geam<double>(h, RTAT_B200_OP_N, RTAT_B200_OP_N, m, n, 0.0, nullptr, ldc, beta, C, ldc, C, ldc);
  }
  DgemmParams p{m, n, k, A, lda, B, ldb, C, ldc, alpha, beta, 0, 0};
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  auto al16 = [](const void *ptr, int ld) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && ld % 2 == 0; };
  bool aligned = al16(A, lda) && al16(B, ldb);
  int variant = opt(h, "dgemm.variant");
  if (opt(h, "dgemm.force_unaligned")) aligned = false;
  // GEMV-shaped calls (C at most 4 rows or columns wide; option dgemm.skinny = 0: off): the skinny kernel reads the large
  // operand once instead of padding the thin extent to a tile (1 x 3001 x 2049: 16 against 27 us, 8192 x 1 x 2048: 34 against 83)
  if (variant == 0 && m > 0 && n > 0) {
    const int skinny_max = opt(h, "dgemm.skinny", 1) != 0 ? 4 : 0;
    int st = RTAT_B200_STATUS_NOT_SUPPORTED;
    if (m <= skinny_max && n >= 256)
      st = dgemm_skinny(h, m, n, k, alpha, A, ta ? lda : 1, ta ? 1 : lda, B, tb ? ldb : 1, tb ? 1 : ldb, beta, C, 1, ldc);
    else if (n <= skinny_max && m >= 256)
      st = dgemm_skinny(h, n, m, k, alpha, B, tb ? 1 : ldb, tb ? ldb : 1, A, ta ? 1 : lda, ta ? lda : 1, beta, C, ldc, 1);
    if (st != RTAT_B200_STATUS_NOT_SUPPORTED) return st;
  }
  // variants: 1/2 = cp.async multistage kernels (128x128 / 64x64 tiles, this file);
  //           3 = persistent warp-specialised kernel, TMA producers when the operands allow it;
  //           4 = same with the cp.async producers forced (dgemm_ws.cu)
  if (variant == 0) {
    // the persistent stream-K kernel fills the machine from ~1 tile x 32 k-tiles upward; below that
    // (a handful of k-tiles in total) the small-tile multistage kernel has less fixed cost
    long long units = (long long)((m + 127) / 128) * ((n + 127) / 128) * ((k + 15) / 16);
    variant = units >= 64 ? 3 : 2;
  }
  if (variant == 3 || variant == 4) {
    bool tma = variant == 3 && !opt(h, "dgemm.force_unaligned");
    int st = dgemm_ws(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tma, opt(h, "dgemm.tile"));  // dgemm.tile: 0 = choose, 64 / 128 = forced CTA tile
    if (st == RTAT_B200_STATUS_NOT_SUPPORTED && tma)  // tensor map could not be encoded: same kernel, cp.async producers
      st = dgemm_ws(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, opt(h, "dgemm.tile"));
    return st;
  }
#define RTAT_DISPATCH(TA_, TB_)                                                            \
  (aligned ? launch_ops<TA_, TB_, true>(h, p, variant) : launch_ops<TA_, TB_, false>(h, p, variant))
  if (!ta && !tb) return RTAT_DISPATCH(false, false);
  if (ta && !tb) return RTAT_DISPATCH(true, false);
  if (!ta && tb) return RTAT_DISPATCH(false, true);
  return RTAT_DISPATCH(true, true);
#undef RTAT_DISPATCH
}

}  // namespace rtat_b200

// SGEMM fallback on the FP32 FMA pipe:  C = alpha * op(A) * op(B) + beta * C   (column-major)
//
// Used by rtat_b200::sgemm when the tcgen05 3xTF32 kernel cannot take the call (operands TMA
// cannot address: base not 16 B aligned or ld % 4 != 0, or very small problems).  True FP32
// arithmetic like the reference's cublasSgemm under default math mode.
// 128x128x8 CTA tile, 256 threads, 8x8 outputs per thread, register-staged double buffering.
#include "common.cuh"

namespace rtat_b200 {

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_simt_kernel(int m, int n, int k, float alpha, const float *__restrict__ A, int lda,
                  const float *__restrict__ B, int ldb, float beta, float *C, int ldc, int tiles_m, int tri) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ float sA[2][BK][BM + 4];
  __shared__ float sB[2][BK][BN + 4];
  const int tid = threadIdx.x;
  int m0 = (blockIdx.x % tiles_m) * BM, n0 = (blockIdx.x / tiles_m) * BN;
  if (tri) {  // SSYRK: blockIdx.x = i (i + 1) / 2 + j enumerates the tiles of the triangle, i >= j
    const int t = blockIdx.x;
    int i = int((__fsqrt_rn(8.0f * float(t) + 1.0f) - 1.0f) * 0.5f);
    while ((long long)i * (i + 1) / 2 > t) i--;
    while ((long long)(i + 1) * (i + 2) / 2 <= t) i++;
    const int j = t - int((long long)i * (i + 1) / 2);
    m0 = (tri == 1 ? i : j) * BM;
    n0 = (tri == 1 ? j : i) * BN;
  }
  const int tx = tid % 16, ty = tid / 16;  // thread owns rows tx*4..+3 and 64+tx*4..+3, cols likewise with ty

  // loader mapping: each thread fetches 4 elements of A and 4 of B per k-tile
  // A tile (BM x BK): contiguous along m when !TA, along k when TA
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      int idx = tid + e * 256;  // 0..1023
      int am, ak;
      if (!TA) { am = idx % BM; ak = idx / BM; } else { ak = idx % BK; am = idx / BK; }
      int gm = m0 + am, gk = k0 + ak;
      ra[e] = (gm < m && gk < k) ? (TA ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm]) : 0.f;
      int bn, bk;
      if (!TB) { bk = idx % BK; bn = idx / BK; } else { bn = idx % BN; bk = idx / BN; }
      int gn = n0 + bn;
      gk = k0 + bk;
      rb[e] = (gn < n && gk < k) ? (TB ? B[(size_t)gk * ldb + gn] : B[(size_t)gn * ldb + gk]) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      int idx = tid + e * 256;
      int am, ak;
      if (!TA) { am = idx % BM; ak = idx / BM; } else { ak = idx % BK; am = idx / BK; }
      sA[buf][ak][am] = ra[e];
      int bn, bk;
      if (!TB) { bk = idx % BK; bn = idx / BK; } else { bn = idx % BN; bk = idx / BN; }
      sB[buf][bk][bn] = rb[e];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const int KT = (k + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < KT; kt++) {
    int buf = kt & 1;
    if (kt + 1 < KT) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[8], b[8];
      *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&sA[buf][kk][tx * 4]);
      *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&sA[buf][kk][64 + tx * 4]);
      *reinterpret_cast<float4 *>(&b[0]) = *reinterpret_cast<const float4 *>(&sB[buf][kk][ty * 4]);
      *reinterpret_cast<float4 *>(&b[4]) = *reinterpret_cast<const float4 *>(&sB[buf][kk][64 + ty * 4]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int col = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
    if (col >= n) continue;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int row = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (row < m && !(tri == 1 && row < col) && !(tri == 2 && row > col)) {
        float v = alpha * acc[i][j];
        if (beta != 0.f) v += beta * C[(size_t)col * ldc + row];
        C[(size_t)col * ldc + row] = v;
      }
    }
  }
}

int sgemm_simt(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
               const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri) {
  int tiles_m = (m + 127) / 128, tiles_n = (n + 127) / 128;
  long long tiles = (long long)tiles_m * tiles_n;
  if (tri) tiles = (long long)tiles_m * (tiles_m + 1) / 2;
  if (tiles > 0x7fffffffLL) return RTAT_B200_STATUS_NOT_SUPPORTED;
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  dim3 grid((unsigned)tiles), block(256);
  if (!ta && !tb) sgemm_simt_kernel<false, false><<<grid, block, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tiles_m, tri);
  else if (ta && !tb) sgemm_simt_kernel<true, false><<<grid, block, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tiles_m, tri);
  else if (!ta && tb) sgemm_simt_kernel<false, true><<<grid, block, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tiles_m, tri);
  else sgemm_simt_kernel<true, true><<<grid, block, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tiles_m, tri);
  return check_launch(h, tri ? "ssyrk_simt_fp32_128x128x8" : "sgemm_simt_fp32_128x128x8");
}

}  // namespace rtat_b200

// SGEMM dispatcher.  Replaces cublasSgemm behind MatrixMult / MatrixMultAlloc for T = float
// (reference src/matrix_ops/matrixop.h:33-42).
#include "common.cuh"

namespace rtat_b200 {

int sgemm_simt(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
               const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri);
int sgemm_tcgen05(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
                  const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri);
bool sgemm_tcgen05_supported(const rtat_b200_context *h, int transa, int transb, int m, int n, int k,
                             const float *A, int lda, const float *B, int ldb, const float *C, int ldc);

int sgemm_skinny(rtat_b200_context *h, int r, int N, int K, float alpha, const float *S, long long ss_i, long long ss_k, const float *X,
                 long long xs_k, long long xs_j, float beta, float *Y, long long ys_i, long long ys_j);

// tri: 0 = GEMM; 1 / 2 = SYRK on the lower / upper triangle (syrk.cu; alpha == 0 and k == 0 are handled there)
int sgemm_tri(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
              const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri);

int sgemm(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
          const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc) {
  return sgemm_tri(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0);
}

int sgemm_tri(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
              const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri) {
  if (!tri && (k == 0 || alpha == 0.f)) {
    if (beta == 0.f) {
      if (cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * m, n, h->stream) != cudaSuccess)
        return RTAT_B200_STATUS_EXECUTION_FAILED;
      h->last_kernel = "memset2d";
      return RTAT_B200_STATUS_SUCCESS;
    }
    if (beta == 1.f) return RTAT_B200_STATUS_SUCCESS;
    return geam<float>(h, RTAT_B200_OP_N, RTAT_B200_OP_N, m, n, 0.f, nullptr, ldc, beta, C, ldc, C, ldc);
  }
  int variant = opt(h, "sgemm.variant");  // 0 auto, 1 = FP32 SIMT, 2 = tcgen05 3xTF32
  // GEMV-shaped calls (C at most 4 rows or columns wide): the skinny FMA kernel of dgemm_skinny.cu in FP32 reads the large
  // operand once; the tensor-core kernel would pad the thin extent to a 128-wide tile (1 x 8192 x 4096: 110 us)
  if (variant == 0 && !tri && m > 0 && n > 0 && opt(h, "dgemm.skinny", 1) != 0) {
    const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
    int st = RTAT_B200_STATUS_NOT_SUPPORTED;
    if (m <= 4 && n >= 256)
      st = sgemm_skinny(h, m, n, k, alpha, A, ta ? lda : 1, ta ? 1 : lda, B, tb ? ldb : 1, tb ? 1 : ldb, beta, C, 1, ldc);
    else if (n <= 4 && m >= 256)
      st = sgemm_skinny(h, n, m, k, alpha, B, tb ? 1 : ldb, tb ? ldb : 1, A, ta ? 1 : lda, ta ? lda : 1, beta, C, ldc, 1);
    if (st != RTAT_B200_STATUS_NOT_SUPPORTED) return st;
  }
  bool tc_ok = sgemm_tcgen05_supported(h, transa, transb, m, n, k, A, lda, B, ldb, C, ldc);
  if (variant == 2 && !tc_ok) return RTAT_B200_STATUS_NOT_SUPPORTED;
  // automatic choice: tensor cores whenever TMA can address the operands, except for problems so
  // small that a single 128x128x32 tile pipeline is mostly padding
  const bool worthwhile = (long long)m * n >= 64 * 64 && k >= 32;
  if (variant == 2 || (variant == 0 && tc_ok && worthwhile)) {
    int st = sgemm_tcgen05(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
    if (st != RTAT_B200_STATUS_NOT_SUPPORTED || variant == 2) return st;
  }
  return sgemm_simt(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
}

}  // namespace rtat_b200

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

// SYRK:  C = alpha * op(A) * op(A)^T + beta * C  on one triangle of C  (replaces cublasDsyrk / cublasSsyrk behind
// MatrixSyrk / MatrixSyrkAlloc, reference src/matrix_ops/matrixop.h:48-76, :536-617).
//
// There is no separate SYRK mainloop: the GEMM kernels take a `tri` mode in which the persistent tile schedule
// enumerates only the T (T + 1) / 2 tiles that touch the requested triangle (tile t = i (i + 1) / 2 + j, i >= j)
// and the epilogue masks the strict other side of the diagonal tiles.  A is passed as both operands with
// opposite ops, so FP64 runs on the DMMA warp-specialised kernel (TMA or cp.async producers, stream-K tail) and
// FP32 on the tcgen05 3xTF32 kernel — the same code paths the GEMM parity tests pin.  Elements outside the
// triangle are never read or written (BLAS contract; the reference's transpose_C plan relies on it:
// syrk.cpp:83-97 zeroes the scratch first because SYRK leaves the other triangle alone).
#include "common.cuh"

namespace rtat_b200 {

int dgemm_ws(rtat_b200_context *h, bool ta, bool tb, int m, int n, int k, double alpha, const double *A, int lda,
             const double *B, int ldb, double beta, double *C, int ldc, bool allow_tma, int tile_cfg, int tri);
int sgemm_tri(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
              const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc, int tri);

// C(triangle) = beta * C(triangle); beta == 0 stores zeros without reading (NaN-safe).  blockIdx.y strides over columns.
template <typename T>
__global__ void tri_scale_kernel(T *C, int ldc, int n, T beta, int lower) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int col = blockIdx.y; col < n; col += gridDim.y) {
    if (lower ? row < col : row > col) continue;
    T *c = C + (size_t)col * ldc + row;
    *c = beta == T(0) ? T(0) : beta * *c;
  }
}

template <typename T>
static int tri_scale(rtat_b200_context *h, int uplo, int n, T beta, T *C, int ldc) {
  if (beta == T(1)) return RTAT_B200_STATUS_SUCCESS;
  dim3 grid((n + 255) / 256, n < 65535 ? n : 65535);
  tri_scale_kernel<T><<<grid, 256, 0, h->stream>>>(C, ldc, n, beta, uplo == RTAT_B200_FILL_LOWER);
  return check_launch(h, "syrk_tri_scale");
}

int dsyrk(rtat_b200_context *h, int uplo, int trans, int n, int k, double alpha, const double *A, int lda, double beta,
          double *C, int ldc) {
  if (k == 0 || alpha == 0.0) return tri_scale<double>(h, uplo, n, beta, C, ldc);
  const int tri = uplo == RTAT_B200_FILL_LOWER ? 1 : 2;
  const bool ta = trans != RTAT_B200_OP_N;
  const bool tma = opt(h, "dgemm.variant") != 4 && !opt(h, "dgemm.force_unaligned");
  int st = dgemm_ws(h, ta, !ta, n, n, k, alpha, A, lda, A, lda, beta, C, ldc, tma, opt(h, "dgemm.tile"), tri);
  if (st == RTAT_B200_STATUS_NOT_SUPPORTED && tma)
    st = dgemm_ws(h, ta, !ta, n, n, k, alpha, A, lda, A, lda, beta, C, ldc, false, opt(h, "dgemm.tile"), tri);
  return st;
}

int ssyrk(rtat_b200_context *h, int uplo, int trans, int n, int k, float alpha, const float *A, int lda, float beta,
          float *C, int ldc) {
  if (k == 0 || alpha == 0.f) return tri_scale<float>(h, uplo, n, beta, C, ldc);
  const int tri = uplo == RTAT_B200_FILL_LOWER ? 1 : 2;
  const bool ta = trans != RTAT_B200_OP_N;
  return sgemm_tri(h, ta ? RTAT_B200_OP_T : RTAT_B200_OP_N, ta ? RTAT_B200_OP_N : RTAT_B200_OP_T, n, n, k, alpha, A, lda, A, lda, beta,
                   C, ldc, tri);
}

}  // namespace rtat_b200

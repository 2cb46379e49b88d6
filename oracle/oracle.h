/*
 * oracle.h — CPU restatement of the reference algorithm for rtatblas' GEMM execution path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rtatblas_b200/ may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker (or the reported CPU baseline), never as a product code path.
 *
 * Parity pinning: the reference has no golden vectors for this path (its arithmetic is cuBLAS,
 * SURVEY.md §8c).  The oracle is pinned against outputs of the UNMODIFIED reference executor
 * (oracle/_ref/ref_driver, built from /root/reference/src against cuBLAS 12.9) captured on a B200
 * and committed under tests/golden/ together with the generating script
 * (tests/golden/make_golden.sh).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#ifndef RTAT_ORACLE_H
#define RTAT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* tests/common.h:158-193 (test_gemm): C = beta*C; C += alpha*A(i,l)*B(l,j) for l ascending, all in
 * the element type; op via index swap; column-major with ld. */
void oracle_dgemm(int transa, int transb, int m, int n, int k, double alpha, const double *A, int lda,
                  const double *B, int ldb, double beta, double *C, int ldc);
void oracle_sgemm(int transa, int transb, int m, int n, int k, float alpha, const float *A, int lda,
                  const float *B, int ldb, float beta, float *C, int ldc);

/* Tolerance helpers for BASELINE.json's bound |C - C_ref| <= c*k*eps*(|A||B|) (+ eps*|beta||C0|):
 * exact[i + j*m]  = alpha*sum(a*b) + beta*c0 accumulated in long double (double for the float
 *                   variant), absab[i + j*m] = |alpha| * sum(|a||b|) + |beta||c0|.  Tight m x n outputs. */
void oracle_dgemm_exact(int transa, int transb, int m, int n, int k, double alpha, const double *A,
                        int lda, const double *B, int ldb, double beta, const double *C0, int ldc,
                        double *exact, double *absab);
void oracle_sgemm_exact(int transa, int transb, int m, int n, int k, float alpha, const float *A,
                        int lda, const float *B, int ldb, float beta, const float *C0, int ldc,
                        double *exact, double *absab);

/* cuBLAS geam as the reference calls it (src/matrix_ops/matrixop.h:111-140, :291-304, :329-343):
 * C = alpha*op(A) + beta*op(B), m x n.  beta == 0: B not read, c = alpha*a;  alpha == 0: A not
 * read, c = beta*b;  otherwise c = fma(beta, b, RN(alpha*a)) — the rounding cuBLAS 12.9 geam shows on
 * a B200, pinned bit-for-bit by tests/golden (mul+mul+add and fma(alpha,a,RN(beta*b)) both mismatch). */
void oracle_dgeam(int transa, int transb, int m, int n, double alpha, const double *A, int lda,
                  double beta, const double *B, int ldb, double *C, int ldc);
void oracle_sgeam(int transa, int transb, int m, int n, float alpha, const float *A, int lda,
                  float beta, const float *B, int ldb, float *C, int ldc);

/* Plan algebra: GEMM_Options::form_operation (src/methods/gemm.cpp:182-221) for 3-letter plans
 * "xyz" (transA,transB,transC in {N,T}) and GEMM_Options_Pad::form_operation (gemm.cpp:82-134) for
 * 6-letter plans "xyzabc" (+ padA,padB,padC in {N,P}).  Workspace accounting follows
 * MatrixOp::workspace_req (src/matrix_ops/matrixop.h:164-189); scratch layout follows
 * compute_operands' peel order (matrixop.h:192-210).  Returns bytes (Executor::calculate_workspace,
 * src/methods/executor.h:41-44) or (size_t)-1 for a malformed plan string. */
size_t oracle_gemm_plan_workspace(const char *plan, int transa, int transb, int m, int n, int k,
                                  size_t elem_size);
/* Executes the plan's operation tree on the CPU with the reference's scratch layout.
 * Returns 0, 1 = bad plan, 2 = workspace too small (executor.h:49-51). */
int oracle_dgemm_plan(const char *plan, int transa, int transb, int m, int n, int k, double alpha,
                      const double *A, int lda, const double *B, int ldb, double beta, double *C,
                      int ldc, void *workspace, size_t workspace_bytes);
int oracle_sgemm_plan(const char *plan, int transa, int transb, int m, int n, int k, float alpha,
                      const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc,
                      void *workspace, size_t workspace_bytes);

/* SYRK / TRSM.  The reference holds no CPU restatement of either: its tests check
 * SYRK against its own GEMM and TRSM by solving a product back (tests/executor_test.cpp:132-290).  These follow
 * the BLAS definitions the reference's cuBLAS calls implement (matrixop.h:48-109), with test_gemm's element
 * arithmetic; lower/trans/left/unit are 0/1 flags. */
void oracle_dsyrk(int lower, int trans, int n, int k, double alpha, const double *A, int lda, double beta,
                  double *C, int ldc);
void oracle_ssyrk(int lower, int trans, int n, int k, float alpha, const float *A, int lda, float beta, float *C,
                  int ldc);
void oracle_dtrsm(int left, int lower, int trans, int unit, int m, int n, double alpha, const double *A, int lda,
                  double *B, int ldb);
void oracle_strsm(int left, int lower, int trans, int unit, int m, int n, float alpha, const float *A, int lda,
                  float *B, int ldb);
/* Plan algebra: SYRK_Options "ac" (transpose_A, transpose_C; src/methods/syrk.cpp:72-104) and TRSM_Options "sa"
 * (swap_side, transpose_A; src/methods/trsm.cpp:72-113), letters F/T.  Same return conventions as the GEMM
 * plan functions above. */
size_t oracle_syrk_plan_workspace(const char *plan, int trans, int n, int k, size_t elem_size);
size_t oracle_trsm_plan_workspace(const char *plan, int left, int m, int n, size_t elem_size);
int oracle_dsyrk_plan(const char *plan, int lower, int trans, int n, int k, double alpha, const double *A, int lda,
                      double beta, double *C, int ldc, void *workspace, size_t workspace_bytes);
int oracle_ssyrk_plan(const char *plan, int lower, int trans, int n, int k, float alpha, const float *A, int lda,
                      float beta, float *C, int ldc, void *workspace, size_t workspace_bytes);
int oracle_dtrsm_plan(const char *plan, int left, int lower, int trans, int unit, int m, int n, double alpha,
                      const double *A, int lda, double *B, int ldb, void *workspace, size_t workspace_bytes);
int oracle_strsm_plan(const char *plan, int left, int lower, int trans, int unit, int m, int n, float alpha,
                      const float *A, int lda, float *B, int ldb, void *workspace, size_t workspace_bytes);

/* Same counter-based generator as rtat_b200_fill_uniform_{d,s} (seeded replacement of the
 * reference drivers' unseeded cuRAND fill, src/app/test_harness.h:176). */
void oracle_fill_uniform_d(double *x, size_t n, uint64_t seed, double lo, double hi);
void oracle_fill_uniform_s(float *x, size_t n, uint64_t seed, float lo, float hi);

#ifdef __cplusplus
}
#endif
#endif

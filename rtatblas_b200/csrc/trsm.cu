// TRSM:  solve  op(A) X = alpha B  (left)  or  X op(A) = alpha B  (right)  in place of B   (replaces cublasDtrsm /
// cublasStrsm behind MatrixTrs / MatrixTrsAlloc, reference src/matrix_ops/matrixop.h:78-109, :442-534).
//
// Recursive blocked substitution: the triangular dimension is halved (at multiples of NB) until a block fits the
// base kernel; between the two halves one GEMM folds the solved half into the other half's right-hand side,
//     left,  op(A) lower:  X1 = A11 \ (alpha B1);  B2 = alpha B2 - A21 X1;  X2 = A22 \ B2
// so for m >> NB practically all flops run in the DMMA / tcgen05 GEMM kernels on large k (the recursion hands the
// GEMMs k = m/2, m/4, ... instead of the k = NB panels of a left-looking sweep).  The base kernel keeps one
// right-hand-side vector per thread in NB registers and the NB x NB triangle in shared memory (every lane reads
// the same element: a broadcast), and uses true divisions by the diagonal — no inverted blocks, so the result is
// a plain substitution's, backward stable like the reference BLAS loop.  Only the `uplo` triangle of A is read
// (not even its diagonal when diag == UNIT).
#include "common.cuh"

namespace rtat_b200 {

namespace trsm {

constexpr int NB = 32;        // base block: vector length held in registers
constexpr int VECS = 128;     // right-hand-side vectors per CTA (one per thread)

// Solves F x = alpha b for VECS vectors per CTA, F = the nb x nb block at A (nb <= NB), normalised so that the
// vectors are columns of B (LEFT) or rows of B (!LEFT, where F = op(A)^T).
//   tr      F(i, j) = tr ? A(j, i) : A(i, j)
//   lower_f F is lower triangular (forward substitution), else upper (backward)
template <typename T, bool LEFT>
__global__ void __launch_bounds__(VECS)
trsm_base_kernel(const T *__restrict__ A, int lda, T *B, int ldb, int nb, int nvec, T alpha, int tr, int lower_f, int unit) {
  __shared__ T sF[NB * NB];               // sF[i * NB + j] = F(i, j); zero outside the triangle, 1 on padded diagonal
  __shared__ T sB[LEFT ? VECS * (NB + 1) : 1];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += VECS) {
    // fastest index follows A's contiguous dimension so the global reads coalesce
    const int fast = idx % NB, slow = idx / NB;
    const int i = tr ? slow : fast, j = tr ? fast : slow;  // F(i, j); A element (tr ? (j, i) : (i, j))
    T v = T(0);
    if (i < nb && j < nb) {
      const bool in_tri = lower_f ? (i > j) : (i < j);
      if (in_tri || (i == j && !unit)) v = tr ? A[(size_t)i * lda + j] : A[(size_t)j * lda + i];
      else if (i == j) v = T(1);
    } else if (i == j) {
      v = T(1);
    }
    sF[i * NB + j] = v;
  }
  const int v0 = blockIdx.x * VECS;
  T x[NB];
  if (LEFT) {
    // stage the nb x VECS panel through shared memory: coalesced column segments in, one padded row per thread out
    for (int idx = tid; idx < NB * VECS; idx += VECS) {
      const int i = idx % NB, v = idx / NB;
      sB[v * (NB + 1) + i] = (i < nb && v0 + v < nvec) ? B[(size_t)(v0 + v) * ldb + i] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = alpha * sB[tid * (NB + 1) + i];
  } else {
    __syncthreads();
    const int r = v0 + tid;
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = (i < nb && r < nvec) ? alpha * B[(size_t)i * ldb + r] : T(0);
  }
  if (lower_f) {
#pragma unroll
    for (int i = 0; i < NB; i++) {
      if (!unit) x[i] = x[i] / sF[i * NB + i];
#pragma unroll
      for (int j = i + 1; j < NB; j++) x[j] -= sF[j * NB + i] * x[i];
    }
  } else {
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
      if (!unit) x[i] = x[i] / sF[i * NB + i];
#pragma unroll
      for (int j = 0; j < i; j++) x[j] -= sF[j * NB + i] * x[i];
    }
  }
  if (LEFT) {
#pragma unroll
    for (int i = 0; i < NB; i++) sB[tid * (NB + 1) + i] = x[i];
    __syncthreads();
    for (int idx = tid; idx < NB * VECS; idx += VECS) {
      const int i = idx % NB, v = idx / NB;
      if (i < nb && v0 + v < nvec) B[(size_t)(v0 + v) * ldb + i] = sB[v * (NB + 1) + i];
    }
  } else {
    const int r = v0 + tid;
    if (r < nvec) {
#pragma unroll
      for (int i = 0; i < NB; i++)
        if (i < nb) B[(size_t)i * ldb + r] = x[i];
    }
  }
}

template <typename T>
struct Problem {
  rtat_b200_context *h;
  bool left, lower, trans, unit;
  int m, n;
  const T *A;
  int lda;
  T *B;
  int ldb;
  int (*gemm)(rtat_b200_context *, int, int, int, int, int, T, const T *, int, const T *, int, T, T *, int);
};

template <typename T>
static int base(const Problem<T> &p, int o, int s, T alpha) {
  const T *Ablk = p.A + (size_t)o * p.lda + o;
  const bool tr = p.left ? p.trans : !p.trans;
  const bool lower_f = tr ? !p.lower : p.lower;
  if (p.left) {
    const int grid = (p.n + VECS - 1) / VECS;
    trsm_base_kernel<T, true><<<grid, VECS, 0, p.h->stream>>>(Ablk, p.lda, p.B + o, p.ldb, s, p.n, alpha, tr, lower_f, p.unit);
  } else {
    const int grid = (p.m + VECS - 1) / VECS;
    trsm_base_kernel<T, false><<<grid, VECS, 0, p.h->stream>>>(Ablk, p.lda, p.B + (size_t)o * p.ldb, p.ldb, s, p.m, alpha, tr, lower_f, p.unit);
  }
  return check_launch(p.h, "trsm_base_nb32");
}

// solves the part of the system that lives on [o, o + s) of the triangular dimension
template <typename T>
static int solve(const Problem<T> &p, int o, int s, T alpha) {
  if (s <= NB) return base(p, o, s, alpha);
  const int blocks = (s + NB - 1) / NB;
  const int s1 = (blocks + 1) / 2 * NB, s2 = s - s1, o1 = o, o2 = o + s1;
  const bool lower_eff = p.lower != p.trans;           // shape of op(A)
  const bool first_first = p.left ? lower_eff : !lower_eff;
  // the off-diagonal block of A that couples the halves lies below the diagonal for a lower A, above for an upper A
  const T *Aoff = p.lower ? p.A + (size_t)o1 * p.lda + o2   // A(o2.., o1..): s2 x s1
                          : p.A + (size_t)o2 * p.lda + o1;  // A(o1.., o2..): s1 x s2
  const int op = p.trans ? RTAT_B200_OP_T : RTAT_B200_OP_N;
  int st;
  if (first_first) {
    if ((st = solve(p, o1, s1, alpha))) return st;
    if (p.left)  // B2 = alpha B2 - op(A)[2,1] X1
      st = p.gemm(p.h, op, RTAT_B200_OP_N, s2, p.n, s1, T(-1), Aoff, p.lda, p.B + o1, p.ldb, alpha, p.B + o2, p.ldb);
    else         // B2 = alpha B2 - X1 op(A)[1,2]
      st = p.gemm(p.h, RTAT_B200_OP_N, op, p.m, s2, s1, T(-1), p.B + (size_t)o1 * p.ldb, p.ldb, Aoff, p.lda, alpha, p.B + (size_t)o2 * p.ldb, p.ldb);
    if (st) return st;
    return solve(p, o2, s2, T(1));
  }
  if ((st = solve(p, o2, s2, alpha))) return st;
  if (p.left)  // B1 = alpha B1 - op(A)[1,2] X2
    st = p.gemm(p.h, op, RTAT_B200_OP_N, s1, p.n, s2, T(-1), Aoff, p.lda, p.B + o2, p.ldb, alpha, p.B + o1, p.ldb);
  else         // B1 = alpha B1 - X2 op(A)[2,1]
    st = p.gemm(p.h, RTAT_B200_OP_N, op, p.m, s1, s2, T(-1), p.B + (size_t)o2 * p.ldb, p.ldb, Aoff, p.lda, alpha, p.B + (size_t)o1 * p.ldb, p.ldb);
  if (st) return st;
  return solve(p, o1, s1, T(1));
}

template <typename T>
static int run(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, T alpha, const T *A, int lda, T *B, int ldb,
               int (*gemm)(rtat_b200_context *, int, int, int, int, int, T, const T *, int, const T *, int, T, T *, int), const char *name) {
  if (alpha == T(0)) {  // BLAS: A is not referenced, B = 0
    if (cudaMemset2DAsync(B, sizeof(T) * ldb, 0, sizeof(T) * m, n, h->stream) != cudaSuccess) return RTAT_B200_STATUS_EXECUTION_FAILED;
    h->last_kernel = "memset2d";
    return RTAT_B200_STATUS_SUCCESS;
  }
  Problem<T> p{h, side == RTAT_B200_SIDE_LEFT, uplo == RTAT_B200_FILL_LOWER, trans != RTAT_B200_OP_N, diag == RTAT_B200_DIAG_UNIT,
               m, n, A, lda, B, ldb, gemm};
  int st = solve(p, 0, p.left ? m : n, alpha);
  if (!st) h->last_kernel = name;
  return st;
}

}  // namespace trsm

int dtrsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *A, int lda, double *B,
          int ldb) {
  return trsm::run<double>(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, dgemm, "dtrsm_recursive_nb32_dmma");
}

int strsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *A, int lda, float *B,
          int ldb) {
  return trsm::run<float>(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, sgemm, "strsm_recursive_nb32_3xtf32");
}

}  // namespace rtat_b200

// trsm.h — the TRSM method: problem description, tuning key, plan space and executor
// (reference src/methods/trsm.h, trsm.cpp).
//
//   op(A) X = alpha B  (side Left)   or   X op(A) = alpha B  (side Right),   X overwrites B (m x n)
//
// A plan is two Bool_Op letters "sa":
//   s  swap sides: solve the transposed system in a scratch copy S = B^T (side and op flipped) and copy S^T back;
//   a  materialise A^T first (an explicit MatrixMove pass) and solve with op and uplo flipped.
#pragma once
#include <cstdio>
#include <memory>
#include <string>
#include <vector>
#include "base_options.h"
#include "executor.h"
#include "matrixop.h"

namespace rtat {

template <typename T>
struct TRSM_Inputs {
  using Scalar = T;
  gpu::blasHandle_t handle;
  BLAS_Side side;
  BLAS_Fill_Mode uplo;
  BLAS_Operation trans;
  BLAS_Diag diag;
  const Matrix<T> A;
  Matrix<T> B;
  const T alpha;

  TRSM_Inputs(gpu::blasHandle_t handle, BLAS_Side side, BLAS_Fill_Mode uplo, BLAS_Operation trans, BLAS_Diag diag, const Matrix<T> A,
              Matrix<T> B, T alpha)
      : handle(handle), side(side), uplo(uplo), trans(trans), diag(diag), A(A), B(B), alpha(alpha) {}

  size_t m() const { return B.dims().m; }
  size_t n() const { return B.dims().n; }
};

// Ordered on the text "<side>,<uplo>,<trans>,<diag>,<m>,<n>" (trsm.cpp:8-20), e.g. "Left,Lower,N,Non-Unit,435,527".
struct TRSM_Key {
  BLAS_Side side;
  BLAS_Fill_Mode uplo;
  BLAS_Operation trans;
  BLAS_Diag diag;
  int m, n;

  TRSM_Key(BLAS_Side side, BLAS_Fill_Mode uplo, BLAS_Operation trans, BLAS_Diag diag, int m, int n)
      : side(side), uplo(uplo), trans(trans), diag(diag), m(m), n(n) {}
  template <typename T>
  TRSM_Key(const TRSM_Inputs<T> &i) : TRSM_Key(i.side, i.uplo, i.trans, i.diag, (int)i.m(), (int)i.n()) {}

  operator std::string() const {
    char buf[96];
    std::snprintf(buf, sizeof buf, "%s,%s,%s,%s,%d,%d", std::string(side).c_str(), std::string(uplo).c_str(), std::string(trans).c_str(),
                  std::string(diag).c_str(), m, n);
    return buf;
  }
  bool operator<(const TRSM_Key &rhs) const { return std::string(*this) < std::string(rhs); }
  // multiply-adds of the substitution: na^2 per right-hand-side vector
  double flopcount() const {
    const double na = side == gpu::BLAS_SIDE_LEFT ? m : n, nrhs = side == gpu::BLAS_SIDE_LEFT ? n : m;
    return na * na * nrhs;
  }
  friend std::ostream &operator<<(std::ostream &os, const TRSM_Key &key) { return os << std::string(key); }
};

struct TRSM_Options {
  Bool_Op swap_side, transpose_A;

  TRSM_Options() = default;
  TRSM_Options(Bool_Op swap_side, Bool_Op transpose_A) : swap_side(swap_side), transpose_A(transpose_A) {}
  static TRSM_Options default_opts() { return TRSM_Options(); }

  // FF, FT, TF, TT
  static std::vector<TRSM_Options> enumerate() {
    std::vector<TRSM_Options> all;
    for (int bits = 0; bits < 4; bits++) all.emplace_back(Bool_Op(bool(bits & 2)), Bool_Op(bool(bits & 1)));
    return all;
  }
  operator std::string() const { return std::string(swap_side) + std::string(transpose_A); }
  bool operator<(const TRSM_Options &o) const { return std::string(*this) < std::string(o); }
  bool fully_fused() const { return !swap_side.set && !transpose_A.set; }
  friend std::ostream &operator<<(std::ostream &os, const TRSM_Options o) { return os << std::string(o); }
  friend std::istream &operator>>(std::istream &is, TRSM_Options &o) {
    std::string s;
    is >> s;
    if (s.size() != 2) {
      is.setstate(std::ios::failbit);
      return is;
    }
    o = TRSM_Options(Bool_Op(s.substr(0, 1)), Bool_Op(s.substr(1, 1)));
    return is;
  }

  template <typename T>
  std::unique_ptr<MatrixOp<T>> form_operation(TRSM_Inputs<T> params) {
    std::unique_ptr<MatrixOp<T>> A = std::make_unique<NoOp<T>>(params.A);
    std::unique_ptr<MatrixOp<T>> B = std::make_unique<NoOp<T>>(params.B);
    bool left = params.side == gpu::BLAS_SIDE_LEFT, lower = params.uplo == gpu::BLAS_FILL_MODE_LOWER;
    bool trans = params.trans == gpu::BLAS_OP_T;
    const bool unit = params.diag == gpu::BLAS_DIAG_UNIT;
    if (transpose_A) {
      trans = !trans;
      lower = !lower;
      A = std::make_unique<MatrixMove<T>>(std::move(A), T(1), true, 1);
    }
    if (swap_side) {
      // (op(A) X)^T = X^T op(A)^T: the same triangle with side and op flipped, solved in S = B^T
      std::unique_ptr<MatrixOp<T>> S = std::make_unique<MatrixMove<T>>(std::move(B), T(1), true, 1);
      S = std::make_unique<MatrixTrsAlloc<T>>(std::move(A), std::move(S), !left, lower, !trans, unit, params.alpha);
      B = std::make_unique<NoOp<T>>(params.B);
      return std::make_unique<MatrixAccumulate<T>>(std::move(S), std::move(B), T(1), T(0), true);
    }
    return std::make_unique<MatrixTrs<T>>(std::move(A), std::move(B), left, lower, trans, unit, params.alpha);
  }
};

template <typename T>
class TRSM_Executor : public Executor<TRSM_Inputs<T>, TRSM_Key, TRSM_Options> {
 protected:
  // One solve per side / triangle / op on a 256 x 256 system (base kernel plus the GEMM kernels the recursion
  // reaches) and the transposing geam, in T (the reference warms cublasDtrsm on 128 x 128: trsm.h:96-126).
  void warmup(TRSM_Inputs<T> params, TRSM_Options, Stream) override {
    constexpr size_t n = 256;
    T *buf = nullptr;
    gpuAssert(gpu::Malloc(&buf, 2 * n * n * sizeof(T)));
    gpuAssert(gpu::Memset(buf, 0, 2 * n * n * sizeof(T)));
    Matrix<T> A(Workspace(buf, n * n), n, n, n), B(Workspace(buf + n * n, n * n), n, n, n);
    for (bool left : {false, true})
      for (bool lower : {false, true})
        for (bool trans : {false, true}) {
          gpuTtrsm<T>(params.handle, left, lower, trans, true, A, B, T(1));  // unit diagonal: the zero fill is a valid system
          gpuTgeam<T>(params.handle, true, false, A, B, B, T(1), T(0));
        }
    gpuAssert(gpu::DeviceSynchronize());
    gpuAssert(gpu::Free(buf));
  }
};

}  // namespace rtat

// syrk.h — the SYRK method: problem description, tuning key, plan space and executor
// (reference src/methods/syrk.h, syrk.cpp).
//
//   C(uplo) = alpha * op(A) * op(A)^T + beta * C(uplo),   op(A) n x k
//
// A plan is two Bool_Op letters "ac":
//   a  materialise A^T first (an explicit MatrixMove pass) and run the update with the flipped op;
//   c  compute the update into the OTHER triangle of a zeroed n x n scratch and fold its transpose back into
//      C with a full-matrix accumulate  C = S^T + beta*C.
// Observable quirk kept from the reference (syrk.cpp:92-97): with c set the accumulate covers all of C, so
// the triangle BLAS would leave untouched is scaled by beta (S is zero there).  The reference's own test
// only inspects the requested triangle (tests/executor_test.cpp:273-284); tests/ here pin both halves
// against the oracle.
#pragma once
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "base_options.h"
#include "executor.h"
#include "matrixop.h"

namespace rtat {

template <typename T>
struct SYRK_Inputs {
  using Scalar = T;
  gpu::blasHandle_t handle;
  BLAS_Fill_Mode uplo;
  BLAS_Operation trans;
  const Matrix<T> A;
  Matrix<T> C;
  const T alpha, beta;

  SYRK_Inputs(gpu::blasHandle_t handle, BLAS_Fill_Mode uplo, BLAS_Operation trans, const Matrix<T> A, Matrix<T> C, T alpha, T beta)
      : handle(handle), uplo(uplo), trans(trans), A(A), C(C), alpha(alpha), beta(beta) {}

  size_t n() const { return C.dims().m; }
  size_t k() const { return trans == gpu::BLAS_OP_N ? A.dims().n : A.dims().m; }
};

// Tuned per triangle, op and extents; ordered on the text "<uplo>,<trans>,<n>,<k>" (syrk.cpp:8-19).
struct SYRK_Key {
  BLAS_Fill_Mode uplo;
  BLAS_Operation trans;
  int n, k;

  SYRK_Key(BLAS_Fill_Mode uplo, BLAS_Operation trans, int n, int k) : uplo(uplo), trans(trans), n(n), k(k) {}
  template <typename T>
  SYRK_Key(const SYRK_Inputs<T> &i) : SYRK_Key(i.uplo, i.trans, (int)i.n(), (int)i.k()) {}

  operator std::string() const {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%s,%s,%d,%d", std::string(uplo).c_str(), std::string(trans).c_str(), n, k);
    return buf;
  }
  bool operator<(const SYRK_Key &rhs) const { return std::string(*this) < std::string(rhs); }
  // multiply-adds on one triangle, diagonal included
  double flopcount() const { return double(n) * (double(n) + 1.0) * k; }
  friend std::ostream &operator<<(std::ostream &os, const SYRK_Key &key) { return os << std::string(key); }
};

struct SYRK_Options {
  Bool_Op transpose_A, transpose_C;

  SYRK_Options() = default;
  SYRK_Options(Bool_Op transpose_A, Bool_Op transpose_C) : transpose_A(transpose_A), transpose_C(transpose_C) {}
  static SYRK_Options default_opts() { return SYRK_Options(); }

  // FF, FT, TF, TT
  static std::vector<SYRK_Options> enumerate() {
    std::vector<SYRK_Options> all;
    for (int bits = 0; bits < 4; bits++) all.emplace_back(Bool_Op(bool(bits & 2)), Bool_Op(bool(bits & 1)));
    return all;
  }
  operator std::string() const { return std::string(transpose_A) + std::string(transpose_C); }
  bool operator<(const SYRK_Options &o) const { return std::string(*this) < std::string(o); }
  bool fully_fused() const { return !transpose_A.set && !transpose_C.set; }
  friend std::ostream &operator<<(std::ostream &os, const SYRK_Options o) { return os << std::string(o); }
  // parses the two letters (the reference builds Bool_Op from a char, which is always true: syrk.cpp:64-65)
  friend std::istream &operator>>(std::istream &is, SYRK_Options &o) {
    std::string s;
    is >> s;
    if (s.size() != 2) {
      is.setstate(std::ios::failbit);
      return is;
    }
    o = SYRK_Options(Bool_Op(s.substr(0, 1)), Bool_Op(s.substr(1, 1)));
    return is;
  }

  template <typename T>
  std::unique_ptr<MatrixOp<T>> form_operation(SYRK_Inputs<T> params) {
    std::unique_ptr<MatrixOp<T>> A = std::make_unique<NoOp<T>>(params.A);
    std::unique_ptr<MatrixOp<T>> C = std::make_unique<NoOp<T>>(params.C);
    bool trans = params.trans == gpu::BLAS_OP_T, lower = params.uplo == gpu::BLAS_FILL_MODE_LOWER;
    if (transpose_A) {
      trans = !trans;
      A = std::make_unique<MatrixMove<T>>(std::move(A), T(1), true, 1);
    }
    if (transpose_C) {
      auto S = std::make_unique<MatrixSyrkAlloc<T>>(std::move(A), !lower, trans, params.alpha);
      return std::make_unique<MatrixAccumulate<T>>(std::move(S), std::move(C), T(1), params.beta, true);
    }
    return std::make_unique<MatrixSyrk<T>>(std::move(A), std::move(C), lower, trans, params.alpha, params.beta);
  }
};

template <typename T>
class SYRK_Executor : public Executor<SYRK_Inputs<T>, SYRK_Key, SYRK_Options> {
 protected:
  // Runs every kernel a SYRK plan can reach once (both triangles and ops on a 128 x 128 x 1024 problem — TMA
  // producers — and on a view TMA cannot address; the transposing geam), in T rather than always in double
  // (the reference warms cublasDsyrk on 64 x 64: syrk.h:88-119).
  void warmup(SYRK_Inputs<T> params, SYRK_Options, Stream) override {
    constexpr size_t n = 128, k = 1024;
    T *buf = nullptr;
    gpuAssert(gpu::Malloc(&buf, (2 * n * k + n * n + 8) * sizeof(T)));
    gpuAssert(gpu::Memset(buf, 0, (2 * n * k + n * n + 8) * sizeof(T)));
    for (size_t shift : {size_t(0), size_t(1)})
      for (bool lower : {false, true})
        for (bool trans : {false, true}) {
          T *a = buf + shift, *at = a + n * k, *c = at + n * k;
          const size_t ar = trans ? k : n, ac = trans ? n : k;
          Matrix<T> A(Workspace(a, n * k), ar, ac, ar), At(Workspace(at, n * k), ac, ar, ac), C(Workspace(c, n * n), n, n, n);
          gpuTsyrk<T>(params.handle, lower, trans, A, C, T(1), T(0));
          gpuTgeam<T>(params.handle, true, false, A, At, At, T(1), T(0));
          gpuTgeam<T>(params.handle, true, false, C, C, C, T(0), T(0));
        }
    gpuAssert(gpu::DeviceSynchronize());
    gpuAssert(gpu::Free(buf));
  }
};

}  // namespace rtat

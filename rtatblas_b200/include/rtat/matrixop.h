// matrixop.h — lazy matrix expressions over the B200 back end (reference src/matrix_ops/matrixop.h).
//
// A MatrixOp describes a matrix that may not have been computed yet.  Executing it needs a BLAS
// handle, space for its own output, and scratch for whatever its operands must materialise.
// Node set used by the GEMM plans (reference gemm.cpp:82-134,182-221):
//   NoOp            an existing matrix
//   ScratchMatrix   an uninitialised matrix of given shape in scratch
//   MatrixMove      B = alpha*op(A) into a fresh buffer with ld = roundup(rows, pad)     -> geam
//   MatrixAccumulate  in place  B = alpha*op(A) + beta*B                                   -> geam
//   MatrixMult      in place  C = alpha*op(A)*op(B) + beta*C                               -> gemm
//   MatrixMultAlloc fresh padded  C = alpha*op(A)*op(B)                                    -> gemm
// and by the SYRK plans (reference syrk.cpp:72-104):
//   MatrixSyrk      in place  C(tri) = alpha*op(A)*op(A)^T + beta*C(tri)                   -> syrk
//   MatrixSyrkAlloc fresh zeroed n x n buffer, then its triangle = alpha*op(A)*op(A)^T     -> geam + syrk
// and by the TRSM plans (reference trsm.cpp:72-113):
//   MatrixTrs       in place  B = alpha * op(A)^-1 B  or  alpha * B op(A)^-1               -> trsm
//   MatrixTrsAlloc  same, but B is an operand that was just materialised in this node's own buffer
// Scratch is carved front-to-back in operand order, so the byte layout of a plan's workspace is the
// reference's (matrixop.h:192-210); tests/ pin it against the oracle.
#pragma once
#include <algorithm>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>
#include "gpu-api.h"
#include "matrix.h"

namespace rtat {

// ---- typed entry points to the boundary (reference matrixop.h:15-46, :111-140) ------------------
template <typename T>
inline gpu::blasStatus_t gpuTgemm(gpu::blasHandle_t handle, bool transa, bool transb, Matrix<T> A, Matrix<T> B, Matrix<T> C,
                                  const T alpha, const T beta) {
  static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "GEMM is only double and float");
  const int m = (int)C.dims().m, n = (int)C.dims().n;
  const int k = (int)(transa ? A.dims().m : A.dims().n);
  const auto opa = transa ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, opb = transb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N;
  if constexpr (std::is_same_v<T, double>)
    return gpu::blasDgemm(handle, opa, opb, m, n, k, &alpha, A.ptr(), (int)A.dims().ld, B.ptr(), (int)B.dims().ld, &beta, C.ptr(), (int)C.dims().ld);
  else
    return gpu::blasSgemm(handle, opa, opb, m, n, k, &alpha, A.ptr(), (int)A.dims().ld, B.ptr(), (int)B.dims().ld, &beta, C.ptr(), (int)C.dims().ld);
}

// C = alpha*op(A) + beta*op(B); the shape is taken from B (the callers pass C == B).
template <typename T>
inline gpu::blasStatus_t gpuTgeam(gpu::blasHandle_t handle, bool transa, bool transb, Matrix<T> A, Matrix<T> B, Matrix<T> C,
                                  const T alpha, const T beta) {
  static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "GEAM is only double and float");
  const int m = (int)B.dims().m, n = (int)B.dims().n;
  const auto opa = transa ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, opb = transb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N;
  if constexpr (std::is_same_v<T, double>)
    return gpu::blasDgeam(handle, opa, opb, m, n, &alpha, A.ptr(), (int)A.dims().ld, &beta, B.ptr(), (int)B.dims().ld, C.ptr(), (int)C.dims().ld);
  else
    return gpu::blasSgeam(handle, opa, opb, m, n, &alpha, A.ptr(), (int)A.dims().ld, &beta, B.ptr(), (int)B.dims().ld, C.ptr(), (int)C.dims().ld);
}

// C(tri) = alpha*op(A)*op(A)^T + beta*C(tri); n from C, k from A (reference matrixop.h:48-76)
template <typename T>
inline gpu::blasStatus_t gpuTsyrk(gpu::blasHandle_t handle, bool lower, bool trans, Matrix<T> A, Matrix<T> C, const T alpha, const T beta) {
  static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "SYRK is only double and float");
  const int n = (int)C.dims().n, k = (int)(trans ? A.dims().m : A.dims().n);
  const auto uplo = lower ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER;
  const auto op = trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N;
  if constexpr (std::is_same_v<T, double>)
    return gpu::blasDsyrk(handle, uplo, op, n, k, &alpha, A.ptr(), (int)A.dims().ld, &beta, C.ptr(), (int)C.dims().ld);
  else
    return gpu::blasSsyrk(handle, uplo, op, n, k, &alpha, A.ptr(), (int)A.dims().ld, &beta, C.ptr(), (int)C.dims().ld);
}

// op(A) X = alpha B or X op(A) = alpha B, X over B; m, n from B (reference matrixop.h:78-109)
template <typename T>
inline gpu::blasStatus_t gpuTtrsm(gpu::blasHandle_t handle, bool side_left, bool lower, bool trans, bool unit_diag, Matrix<T> A, Matrix<T> B,
                                  const T alpha) {
  static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "TRSM is only double and float");
  const int m = (int)B.dims().m, n = (int)B.dims().n;
  const auto side = side_left ? gpu::BLAS_SIDE_LEFT : gpu::BLAS_SIDE_RIGHT;
  const auto uplo = lower ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER;
  const auto op = trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N;
  const auto diag = unit_diag ? gpu::BLAS_DIAG_UNIT : gpu::BLAS_DIAG_NON_UNIT;
  if constexpr (std::is_same_v<T, double>)
    return gpu::blasDtrsm(handle, side, uplo, op, diag, m, n, &alpha, A.ptr(), (int)A.dims().ld, B.ptr(), (int)B.dims().ld);
  else
    return gpu::blasStrsm(handle, side, uplo, op, diag, m, n, &alpha, A.ptr(), (int)A.dims().ld, B.ptr(), (int)B.dims().ld);
}

inline size_t round_up_to(size_t x, size_t pad) { return ((x + pad - 1) / pad) * pad; }

// ---- expression node base -------------------------------------------------------------------------
template <typename T>
class MatrixOp {
 protected:
  std::vector<std::unique_ptr<MatrixOp>> operands;
  int output_operand = -1;  // index of the operand whose buffer *is* this node's output, or -1

 public:
  MatrixOp(const MatrixOp &) = delete;
  MatrixOp(MatrixOp &&) = default;
  MatrixOp &operator=(MatrixOp &&) = default;
  explicit MatrixOp(std::vector<std::unique_ptr<MatrixOp>> ops, int output_operand = -1)
      : operands(std::move(ops)), output_operand(output_operand) {}
  virtual ~MatrixOp() = default;

  virtual Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) = 0;
  virtual size_t output_space_req() const = 0;  // elements this node needs for its own result
  virtual MatrixDims dims() const = 0;

  // Elements needed in total: own output + the high-water mark of operand outputs that must be
  // alive at once plus the scratch of the operand being computed (reference matrixop.h:172-185).
  size_t workspace_req() const {
    size_t live = 0, peak = 0;
    for (int i = 0; i < (int)operands.size(); i++) {
      if (i != output_operand) live += operands[i]->output_space_req();
      peak = std::max(peak, live + operands[i]->scratch_space_req());
    }
    return output_space_req() + peak;
  }
  size_t scratch_space_req() const { return workspace_req() - output_space_req(); }
  size_t workspace_req_bytes() const { return sizeof(T) * workspace_req(); }
  size_t scratch_space_req_bytes() const { return sizeof(T) * scratch_space_req(); }

 protected:
  // Materialise every operand.  The operand that doubles as the output gets `out_space`; the others
  // are peeled off the front of `scratch_space` in order, and each runs with what is left.
  std::vector<Matrix<T>> compute_operands(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) {
    std::vector<Matrix<T>> result;
    result.reserve(operands.size());
    for (int i = 0; i < (int)operands.size(); i++) {
      Workspace home = (i == output_operand) ? out_space : scratch_space.template peel<T>(operands[i]->output_space_req());
      result.push_back(operands[i]->execute(handle, home, scratch_space));
    }
    return result;
  }

  void require_space(Workspace &out_space, Workspace &scratch_space, const char *who) const {
    if (out_space.size<T>() < output_space_req() || scratch_space.size<T>() < scratch_space_req()) {
      std::cout << who << " NOT ENOUGH SPACE" << std::endl;
      throw "Not enough space";
    }
  }
};

template <typename T>
class NoOp : public MatrixOp<T> {
  Matrix<T> A;

 public:
  explicit NoOp(Matrix<T> A) : MatrixOp<T>({}), A(A) {}
  Matrix<T> execute(gpu::blasHandle_t, Workspace, Workspace) override { return A; }
  size_t output_space_req() const override { return 0; }
  MatrixDims dims() const override { return A.dims(); }
};

template <typename T>
class ScratchMatrix : public MatrixOp<T> {
  MatrixDims d;

 public:
  explicit ScratchMatrix(MatrixDims dims) : MatrixOp<T>({}), d(dims) {}
  ScratchMatrix(size_t m, size_t n, size_t ld) : ScratchMatrix(MatrixDims(m, n, ld)) {}
  ScratchMatrix(std::unique_ptr<MatrixOp<T>> &like, int pad)
      : ScratchMatrix(like->dims().m, like->dims().n, round_up_to(like->dims().m, pad)) {}
  size_t output_space_req() const override { return d.footprint(); }
  MatrixDims dims() const override { return d; }
  Matrix<T> execute(gpu::blasHandle_t, Workspace out_space, Workspace) override { return Matrix<T>(out_space, d); }
};

// in place: B = alpha*op(A) + beta*B   (B is operand 1 and the output)
template <typename T>
class MatrixAccumulate : public MatrixOp<T> {
  T alpha, beta;
  bool transpose;

 public:
  MatrixAccumulate(std::unique_ptr<MatrixOp<T>> Aop, std::unique_ptr<MatrixOp<T>> Bop, T alpha, T beta, bool transpose)
      : MatrixOp<T>({}, 1), alpha(alpha), beta(beta), transpose(transpose) {
    const MatrixDims a = Aop->dims(), b = Bop->dims();
    const bool ok = transpose ? (a.m == b.n && a.n == b.m) : (a.m == b.m && a.n == b.n);
    if (!ok) {
      std::cout << "Bad matrix accumulate, Adims=" << a.m << "," << a.n << " Bdims=" << b.m << "," << b.n << " "
                << (transpose ? "trans" : "notrans") << std::endl;
      std::terminate();  // the reference's bare `throw;` (matrixop.h:283)
    }
    this->operands.push_back(std::move(Aop));
    this->operands.push_back(std::move(Bop));
  }
  size_t output_space_req() const override { return 0; }
  MatrixDims dims() const override { return this->operands[1]->dims(); }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    this->require_space(out_space, scratch_space, "ACCUMULATE");
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    blas_status_check(gpuTgeam<T>(handle, transpose, false, ms[0], ms[1], ms[1], alpha, beta), "MatrixAccumulate: geam");
    return ms[1];
  }
};

// fresh buffer: B = alpha*op(A), ld(B) = roundup(rows(B), pad); rows [m, ld) are never written
template <typename T>
class MatrixMove : public MatrixOp<T> {
  T alpha;
  bool transpose;
  size_t pad;

 public:
  MatrixMove(std::unique_ptr<MatrixOp<T>> Aop, T alpha, bool transpose, size_t pad)
      : MatrixOp<T>({}), alpha(alpha), transpose(transpose), pad(pad) {
    this->operands.push_back(std::move(Aop));
  }
  size_t output_space_req() const override { return dims().footprint(); }
  MatrixDims dims() const override {
    const MatrixDims a = this->operands[0]->dims();
    const size_t m = transpose ? a.n : a.m, n = transpose ? a.m : a.n;
    return MatrixDims(m, n, round_up_to(m, pad));
  }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    this->require_space(out_space, scratch_space, "MATRIX MOVE");
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    Matrix<T> B(out_space, dims());
    // beta = 0: the back end never reads B, which is uninitialised scratch here
    blas_status_check(gpuTgeam<T>(handle, transpose, false, ms[0], B, B, alpha, T(0)), "MatrixMove: geam");
    return B;
  }
};

// in place: C = alpha*op(A)*op(B) + beta*C   (C is operand 2 and the output)
template <typename T>
class MatrixMult : public MatrixOp<T> {
 protected:
  bool transa, transb;
  T alpha, beta;

 public:
  MatrixMult(std::unique_ptr<MatrixOp<T>> Aop, std::unique_ptr<MatrixOp<T>> Bop, std::unique_ptr<MatrixOp<T>> Cop, bool transa,
             bool transb, T alpha, T beta)
      : MatrixOp<T>({}, 2), transa(transa), transb(transb), alpha(alpha), beta(beta) {
    const MatrixDims a = Aop->dims(), b = Bop->dims(), c = Cop->dims();
    const size_t kA = transa ? a.m : a.n, kB = transb ? b.n : b.m;
    const size_t mA = transa ? a.n : a.m, nB = transb ? b.m : b.n;
    if (kA != kB || mA != c.m || nB != c.n) {
      // the reference only prints on an m/n mismatch (matrixop.h:366-369); every mismatch is fatal here
      std::cout << "Bad matrix mult, kA=" << kA << " kB=" << kB << " mA=" << mA << " mC=" << c.m << " nB=" << nB << " nC=" << c.n << std::endl;
      std::terminate();
    }
    this->operands.push_back(std::move(Aop));
    this->operands.push_back(std::move(Bop));
    this->operands.push_back(std::move(Cop));
  }
  MatrixDims dims() const override { return this->operands[2]->dims(); }
  size_t output_space_req() const override { return 0; }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    blas_status_check(gpuTgemm<T>(handle, transa, transb, ms[0], ms[1], ms[2], alpha, beta), "MatrixMult: gemm");
    return ms[2];
  }
};

// fresh buffer: C = alpha*op(A)*op(B), ld(C) = roundup(m, pad)
template <typename T>
class MatrixMultAlloc : public MatrixOp<T> {
  bool transa, transb;
  T alpha;
  size_t pad;

 public:
  MatrixMultAlloc(std::unique_ptr<MatrixOp<T>> Aop, std::unique_ptr<MatrixOp<T>> Bop, bool transa, bool transb, T alpha, size_t pad)
      : MatrixOp<T>({}), transa(transa), transb(transb), alpha(alpha), pad(pad) {
    const MatrixDims a = Aop->dims(), b = Bop->dims();
    const size_t kA = transa ? a.m : a.n, kB = transb ? b.n : b.m;
    if (kA != kB) {
      std::cout << "Bad matrix mult, kA=" << kA << " kB=" << kB << std::endl;
      std::terminate();
    }
    this->operands.push_back(std::move(Aop));
    this->operands.push_back(std::move(Bop));
  }
  MatrixDims dims() const override {
    const MatrixDims a = this->operands[0]->dims(), b = this->operands[1]->dims();
    const size_t m = transa ? a.n : a.m, n = transb ? b.m : b.n;
    return MatrixDims(m, n, round_up_to(m, pad));
  }
  size_t output_space_req() const override { return dims().footprint(); }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    Matrix<T> C(out_space, dims());
    blas_status_check(gpuTgemm<T>(handle, transa, transb, ms[0], ms[1], C, alpha, T(0)), "MatrixMultAlloc: gemm");
    return C;
  }
};

// Triangular solve in place of operand 1.  OWNS_OUTPUT = false: B is an existing matrix (MatrixTrs, reference
// matrixop.h:442-486).  OWNS_OUTPUT = true: B is produced by operand 1 directly into this node's output buffer,
// whose footprint this node therefore reports (MatrixTrsAlloc, matrixop.h:488-534).
template <typename T, bool OWNS_OUTPUT>
class MatrixTrsBase : public MatrixOp<T> {
 protected:
  bool side_left, lower, trans, unit_diag;
  T alpha;
  size_t pad;

 public:
  MatrixTrsBase(std::unique_ptr<MatrixOp<T>> Aop, std::unique_ptr<MatrixOp<T>> Bop, bool side_left, bool lower, bool trans, bool unit_diag,
                T alpha, size_t pad = 1)
      : MatrixOp<T>({}, 1), side_left(side_left), lower(lower), trans(trans), unit_diag(unit_diag), alpha(alpha), pad(pad) {
    const MatrixDims a = Aop->dims(), b = Bop->dims();
    if (a.m != a.n || (side_left ? b.m : b.n) != a.m) {
      std::cout << "Bad matrix trs, mA=" << a.m << " nA=" << a.n << " mB=" << b.m << " nB=" << b.n << " " << (side_left ? "LEFT" : "RIGHT")
                << std::endl;
      std::terminate();  // the reference's bare `throw;` (matrixop.h:459, :508)
    }
    this->operands.push_back(std::move(Aop));
    this->operands.push_back(std::move(Bop));
  }
  MatrixDims dims() const override {
    const MatrixDims b = this->operands[1]->dims();
    return OWNS_OUTPUT ? MatrixDims(b.m, b.n, round_up_to(b.m, pad)) : b;
  }
  size_t output_space_req() const override { return OWNS_OUTPUT ? dims().footprint() : 0; }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    blas_status_check(gpuTtrsm<T>(handle, side_left, lower, trans, unit_diag, ms[0], ms[1], alpha), "MatrixTrs: trsm");
    return ms[1];
  }
};
template <typename T>
using MatrixTrs = MatrixTrsBase<T, false>;
template <typename T>
using MatrixTrsAlloc = MatrixTrsBase<T, true>;

// in place: C(tri) = alpha*op(A)*op(A)^T + beta*C(tri)   (C is operand 1 and the output)
template <typename T>
class MatrixSyrk : public MatrixOp<T> {
 protected:
  bool lower, trans;
  T alpha, beta;

 public:
  MatrixSyrk(std::unique_ptr<MatrixOp<T>> Aop, std::unique_ptr<MatrixOp<T>> Cop, bool lower, bool trans, T alpha, T beta)
      : MatrixOp<T>({}, 1), lower(lower), trans(trans), alpha(alpha), beta(beta) {
    const MatrixDims a = Aop->dims(), c = Cop->dims();
    if ((trans ? a.n : a.m) != c.n || c.m != c.n) {
      std::cout << "Bad matrix syrk, mA=" << a.m << " nA=" << a.n << " mC=" << c.m << " nC=" << c.n << std::endl;
      std::terminate();  // the reference's bare `throw;` (matrixop.h:552)
    }
    this->operands.push_back(std::move(Aop));
    this->operands.push_back(std::move(Cop));
  }
  MatrixDims dims() const override { return this->operands[1]->dims(); }
  size_t output_space_req() const override { return 0; }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    blas_status_check(gpuTsyrk<T>(handle, lower, trans, ms[0], ms[1], alpha, beta), "MatrixSyrk: syrk");
    return ms[1];
  }
};

// fresh n x n buffer, zero-filled, whose `lower`/upper triangle then receives alpha*op(A)*op(A)^T.  The zero
// fill matters: the caller folds the WHOLE buffer back with a transposed accumulate (syrk.cpp:92-97).
template <typename T>
class MatrixSyrkAlloc : public MatrixOp<T> {
 protected:
  bool lower, trans;
  T alpha;
  size_t pad;

 public:
  MatrixSyrkAlloc(std::unique_ptr<MatrixOp<T>> Aop, bool lower, bool trans, T alpha, size_t pad = 1)
      : MatrixOp<T>({}), lower(lower), trans(trans), alpha(alpha), pad(pad) {
    this->operands.push_back(std::move(Aop));
  }
  MatrixDims dims() const override {
    const MatrixDims a = this->operands[0]->dims();
    const size_t n = trans ? a.n : a.m;
    return MatrixDims(n, n, round_up_to(n, pad));
  }
  size_t output_space_req() const override { return dims().footprint(); }
  Matrix<T> execute(gpu::blasHandle_t handle, Workspace out_space, Workspace scratch_space) override {
    auto ms = this->compute_operands(handle, out_space, scratch_space);
    Matrix<T> C(out_space, dims());
    // alpha = beta = 0: neither operand is read, C is just cleared (reference matrixop.h:612)
    blas_status_check(gpuTgeam<T>(handle, false, false, C, C, C, T(0), T(0)), "MatrixSyrkAlloc: geam");
    blas_status_check(gpuTsyrk<T>(handle, lower, trans, ms[0], C, alpha, T(0)), "MatrixSyrkAlloc: syrk");
    return C;
  }
};

}  // namespace rtat

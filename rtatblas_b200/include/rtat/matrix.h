// matrix.h — column-major matrix views (reference src/matrix_ops/matrix.h).
#pragma once
#include <ostream>
#include "workspace.h"

namespace rtat {

// m x n, column-major, leading dimension ld >= m; occupies ld*n elements.
struct MatrixDims {
  size_t m = 0, n = 0, ld = 0;
  MatrixDims() = default;
  MatrixDims(size_t m, size_t n, size_t ld) : m(m), n(n), ld(ld) {
    if (ld < m) throw "ld < m in MatrixDims";
  }
  size_t footprint() const { return ld * n; }
  friend std::ostream &operator<<(std::ostream &os, const MatrixDims &d) { return os << "(" << d.m << "," << d.n << "," << d.ld << ")"; }
};

// A MatrixDims laid over a Workspace.  Unlike the reference (whose size check compares a value
// with itself, matrix.h:33-34) the span really is checked against the footprint.
template <typename T>
class Matrix {
  Workspace home;
  MatrixDims dimensions;

 public:
  Matrix() = default;
  Matrix(Workspace home, MatrixDims dims) : home(home), dimensions(dims) {
    if (this->home.template size<T>() < dims.footprint()) throw "MATRIX WORKSPACE TOO SMALL";
  }
  Matrix(Workspace home, size_t m, size_t n, size_t ld) : Matrix(home, MatrixDims(m, n, ld)) {}

  // Column block [col, col + ncols): a zero-copy view (what the multi-GPU column split uses).
  Matrix column_block(size_t col, size_t ncols) const {
    if (col + ncols > dimensions.n) throw "Matrix::column_block out of range";
    Workspace copy = home;
    Workspace sub(copy, col * dimensions.ld * sizeof(T), ncols * dimensions.ld * sizeof(T));
    return Matrix(sub, dimensions.m, ncols, dimensions.ld);
  }

  size_t footprint() const { return dimensions.footprint(); }
  const MatrixDims dims() const { return dimensions; }
  T *ptr() { return static_cast<T *>(home); }
  const T *ptr() const { return static_cast<T *>(const_cast<Workspace &>(home)); }
};

}  // namespace rtat

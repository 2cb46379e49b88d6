// workspace.h — caller-provided device scratch (reference src/matrix_ops/workspace.h).
// Workspace is a non-owning byte span that sub-operations carve pieces off the front of;
// ManagedWorkspace owns a cudaMalloc'd span and can grow (contents are not preserved).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <exception>
#include <iostream>
#include "gpu-api.h"

namespace rtat {

class Workspace {
 protected:
  size_t count = 0;  // bytes
  char *ptr = nullptr;

 public:
  Workspace() = default;
  virtual ~Workspace() = default;
  template <typename T>
  Workspace(T *p, size_t n) : count(n * sizeof(T)), ptr(reinterpret_cast<char *>(p)) {}

  // Sub-range [offset, offset + bytes) of `other`.  Stepping outside the parent is a programming
  // error and ends the process, which is what the reference's death test pins
  // (workspace.h:24-30 `throw;` with no active exception; tests/workspace_test.cpp:48).
  Workspace(Workspace &other, size_t offset, size_t bytes) : count(bytes), ptr(other.ptr + offset) {
    if (offset + bytes > other.count) {
      std::cout << "WORKSPACE OFFSET ERROR" << std::endl;
      std::terminate();
    }
  }

  // Remove `n` elements of T from the front and hand them out.
  template <typename T>
  Workspace peel(size_t n) {
    const size_t bytes = n * sizeof(T);
    Workspace front(ptr, bytes);
    ptr += bytes;
    count -= bytes;
    return front;
  }

  template <typename T>
  operator T *() { return reinterpret_cast<T *>(ptr); }
  template <typename T>
  size_t size() const { return count / sizeof(T); }
};

class ManagedWorkspace : public Workspace {
 public:
  explicit ManagedWorkspace(size_t bytes) {
    gpuAssert(gpu::Malloc(&ptr, bytes));
    count = bytes;
  }
  ManagedWorkspace(const ManagedWorkspace &) = delete;
  ManagedWorkspace &operator=(const ManagedWorkspace &) = delete;
  ~ManagedWorkspace() override { gpuAssert(gpu::Free(ptr)); }

  // Grow-only; device-wide sync because queued work may still use the old span (workspace.h:57-65).
  template <typename T>
  void grow_to_fit(size_t n) {
    if (size<T>() >= n) return;
    gpuAssert(gpu::DeviceSynchronize());
    gpuAssert(gpu::Free(ptr));
    ptr = nullptr;
    gpuAssert(gpu::Malloc(&ptr, n * sizeof(T)));
    count = n * sizeof(T);
  }
};

}  // namespace rtat

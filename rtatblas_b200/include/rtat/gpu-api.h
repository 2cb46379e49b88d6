// gpu-api.h — the rtat::gpu:: vocabulary on B200.
//
// Mirrors the reference's alias table (reference src/gpu-api/gpu-api.h:30-126): runtime entries
// still resolve to the CUDA runtime, but every `blas*` name now resolves to the sm_100a back end
// behind the C ABI in include/rtat_b200.h instead of cuBLAS.  RAII Stream / Event / Device_RNG and
// the string-backed enums keep the reference's observable behaviour (gpu-api.h:142-325,
// gpu-api.cpp:1-76): copies share the underlying object, Event::elapsed_time yields NaN on error,
// gpuAssert reports and carries on.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <memory>
#include <vector>
#include <stdexcept>
#include <string>
#include <utility>
#include "../../../include/rtat_b200.h"

namespace rtat {
namespace gpu {

// ---- runtime ------------------------------------------------------------------------------------
using Error_t = cudaError_t;
using Stream_t = cudaStream_t;
using Event_t = cudaEvent_t;
constexpr Error_t Success = cudaSuccess;
constexpr Error_t ErrorInvalidResourceHandle = cudaErrorInvalidResourceHandle;

inline const char *GetErrorString(Error_t e) { return cudaGetErrorString(e); }
inline Error_t SetDevice(int d) { return cudaSetDevice(d); }
inline Error_t GetDevice(int *d) { return cudaGetDevice(d); }
inline Error_t GetDeviceCount(int *n) { return cudaGetDeviceCount(n); }
inline Error_t DeviceSynchronize() { return cudaDeviceSynchronize(); }
inline Error_t MemGetInfo(size_t *free_b, size_t *total_b) { return cudaMemGetInfo(free_b, total_b); }
template <typename T>
inline Error_t Malloc(T **ptr, size_t bytes) { return cudaMalloc(reinterpret_cast<void **>(ptr), bytes); }
inline Error_t Free(void *ptr) { return cudaFree(ptr); }
inline Error_t StreamCreate(Stream_t *s) { return cudaStreamCreate(s); }
inline Error_t StreamDestroy(Stream_t s) { return cudaStreamDestroy(s); }
inline Error_t StreamSynchronize(Stream_t s) { return cudaStreamSynchronize(s); }
inline Error_t StreamWaitEvent(Stream_t s, Event_t e, unsigned flags) { return cudaStreamWaitEvent(s, e, flags); }
inline Error_t EventCreate(Event_t *e) { return cudaEventCreate(e); }
inline Error_t EventDestroy(Event_t e) { return cudaEventDestroy(e); }
inline Error_t EventRecord(Event_t e, Stream_t s) { return cudaEventRecord(e, s); }
inline Error_t EventSynchronize(Event_t e) { return cudaEventSynchronize(e); }
inline Error_t EventElapsedTime(float *ms, Event_t a, Event_t b) { return cudaEventElapsedTime(ms, a, b); }
inline Error_t EventQuery(Event_t e) { return cudaEventQuery(e); }
constexpr auto MemcpyDeviceToDevice = cudaMemcpyDeviceToDevice;
constexpr auto MemcpyDeviceToHost = cudaMemcpyDeviceToHost;
constexpr auto MemcpyHostToDevice = cudaMemcpyHostToDevice;
inline Error_t Memcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k) { return cudaMemcpy(dst, src, n, k); }
inline Error_t MemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, Stream_t s) { return cudaMemcpyAsync(dst, src, n, k, s); }
inline Error_t Memset(void *dst, int v, size_t n) { return cudaMemset(dst, v, n); }

// ---- BLAS: the plug-in boundary -----------------------------------------------------------------
using blasHandle_t = rtat_b200_handle_t;
using blasStatus_t = int;
enum blasOperation_t { BLAS_OP_N = RTAT_B200_OP_N, BLAS_OP_T = RTAT_B200_OP_T };
enum blasFillMode_t { BLAS_FILL_MODE_LOWER = 0, BLAS_FILL_MODE_UPPER = 1 };
enum blasSideMode_t { BLAS_SIDE_LEFT = 0, BLAS_SIDE_RIGHT = 1 };
enum blasDiagType_t { BLAS_DIAG_NON_UNIT = 0, BLAS_DIAG_UNIT = 1 };
constexpr blasStatus_t BLAS_STATUS_SUCCESS = RTAT_B200_STATUS_SUCCESS;

inline blasStatus_t blasCreate(blasHandle_t *h) { return rtat_b200_create(h); }
inline blasStatus_t blasDestroy(blasHandle_t h) { return rtat_b200_destroy(h); }
inline blasStatus_t blasSetStream(blasHandle_t h, Stream_t s) { return rtat_b200_set_stream(h, s); }
inline blasStatus_t blasGetStream(blasHandle_t h, Stream_t *s) { return rtat_b200_get_stream(h, reinterpret_cast<void **>(s)); }
inline blasStatus_t blasDgemm(blasHandle_t h, blasOperation_t ta, blasOperation_t tb, int m, int n, int k, const double *alpha,
                              const double *A, int lda, const double *B, int ldb, const double *beta, double *C, int ldc) {
  return rtat_b200_dgemm(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
inline blasStatus_t blasSgemm(blasHandle_t h, blasOperation_t ta, blasOperation_t tb, int m, int n, int k, const float *alpha,
                              const float *A, int lda, const float *B, int ldb, const float *beta, float *C, int ldc) {
  return rtat_b200_sgemm(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
inline blasStatus_t blasDgeam(blasHandle_t h, blasOperation_t ta, blasOperation_t tb, int m, int n, const double *alpha,
                              const double *A, int lda, const double *beta, const double *B, int ldb, double *C, int ldc) {
  return rtat_b200_dgeam(h, ta, tb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
}
inline blasStatus_t blasSgeam(blasHandle_t h, blasOperation_t ta, blasOperation_t tb, int m, int n, const float *alpha,
                              const float *A, int lda, const float *beta, const float *B, int ldb, float *C, int ldc) {
  return rtat_b200_sgeam(h, ta, tb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
}
inline blasStatus_t blasDsyrk(blasHandle_t h, blasFillMode_t uplo, blasOperation_t trans, int n, int k, const double *alpha,
                              const double *A, int lda, const double *beta, double *C, int ldc) {
  return rtat_b200_dsyrk(h, uplo, trans, n, k, alpha, A, lda, beta, C, ldc);
}
inline blasStatus_t blasSsyrk(blasHandle_t h, blasFillMode_t uplo, blasOperation_t trans, int n, int k, const float *alpha,
                              const float *A, int lda, const float *beta, float *C, int ldc) {
  return rtat_b200_ssyrk(h, uplo, trans, n, k, alpha, A, lda, beta, C, ldc);
}
inline blasStatus_t blasDtrsm(blasHandle_t h, blasSideMode_t side, blasFillMode_t uplo, blasOperation_t trans, blasDiagType_t diag, int m,
                              int n, const double *alpha, const double *A, int lda, double *B, int ldb) {
  return rtat_b200_dtrsm(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);
}
inline blasStatus_t blasStrsm(blasHandle_t h, blasSideMode_t side, blasFillMode_t uplo, blasOperation_t trans, blasDiagType_t diag, int m,
                              int n, const float *alpha, const float *A, int lda, float *B, int ldb) {
  return rtat_b200_strsm(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);
}

}  // namespace gpu

// Report-and-continue, like the reference (gpu-api.h:128-138): CUDA errors are printed, not thrown.
inline void gpu_error_check(gpu::Error_t code, const char *file, int line) {
  if (code != gpu::Success) std::cerr << "GPU Error: " << gpu::GetErrorString(code) << " " << file << " " << line << std::endl;
}
#define gpuAssert(ans) ::rtat::gpu_error_check((ans), __FILE__, __LINE__)

// BLAS status is checked here, unlike the reference (which drops it: matrixop.h:302,340,391,436).
inline void blas_status_check(gpu::blasStatus_t st, const char *what) {
  if (st != gpu::BLAS_STATUS_SUCCESS)
    throw std::runtime_error(std::string(what) + " failed: " + rtat_b200_status_string(st));
}

class Event;

// Shared-ownership stream.  Default construction creates a stream; construction from a raw
// handle borrows it (never destroyed here).
class Stream {
  std::shared_ptr<CUstream_st> impl;

 public:
  Stream() {
    gpu::Stream_t s = nullptr;
    gpuAssert(gpu::StreamCreate(&s));
    impl = std::shared_ptr<CUstream_st>(s, [](gpu::Stream_t p) { if (p) gpuAssert(gpu::StreamDestroy(p)); });
  }
  Stream(gpu::Stream_t borrowed) : impl(borrowed, [](gpu::Stream_t) {}) {}
  operator gpu::Stream_t() const { return impl.get(); }
  void wait_event(const Event &e);
  void synchronize() { gpuAssert(gpu::StreamSynchronize(*this)); }
};

namespace detail {
// Every timed execute needs two events (reference src/methods/executor.h:18-33); creating and destroying them per call
// costs more host time than launching a small GEMM, so released events go back to a per-thread, per-device free list.
// The lists are deliberately never torn down: the driver reclaims the events with the context.
struct Event_Pool {
  std::vector<std::vector<gpu::Event_t>> free_by_device;
  static Event_Pool &get() {
    static thread_local Event_Pool *pool = new Event_Pool;
    return *pool;
  }
  static std::vector<gpu::Event_t> &list_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
    auto &lists = get().free_by_device;
    if ((int)lists.size() <= dev) lists.resize(dev + 1);
    return lists[dev];
  }
};
}  // namespace detail

class Event {
  std::shared_ptr<CUevent_st> impl;

 public:
  Event() {
    auto &free_list = detail::Event_Pool::list_for_current_device();
    gpu::Event_t e = nullptr;
    if (!free_list.empty()) {
      e = free_list.back();
      free_list.pop_back();
    } else {
      gpuAssert(gpu::EventCreate(&e));
    }
    int dev = 0;
    (void)cudaGetDevice(&dev);
    // the deleter files the event under the device it was created on, in the pool of whichever thread drops the last reference
    impl = std::shared_ptr<CUevent_st>(e, [dev](gpu::Event_t p) {
      if (!p) return;
      auto &lists = detail::Event_Pool::get().free_by_device;
      if ((int)lists.size() <= dev) lists.resize(dev + 1);
      if (lists[dev].size() < 4096) lists[dev].push_back(p);
      else gpuAssert(gpu::EventDestroy(p));
    });
  }
  Event(gpu::Event_t borrowed) : impl(borrowed, [](gpu::Event_t) {}) {}
  operator gpu::Event_t() const { return impl.get(); }
  void record(Stream s) { gpuAssert(gpu::EventRecord(*this, s)); }
  void synchronize() { gpuAssert(gpu::EventSynchronize(*this)); }
  bool query() { return gpu::EventQuery(*this) == gpu::Success; }
  // milliseconds; NaN when the pair cannot be timed (reference gpu-api.cpp:69-75)
  static float elapsed_time(Event start, Event end) {
    float ms = 0.f;
    if (gpu::EventElapsedTime(&ms, start, end) != gpu::Success) {
      (void)cudaGetLastError();
      return NAN;
    }
    return ms;
  }
};

inline void Stream::wait_event(const Event &e) { gpuAssert(gpu::StreamWaitEvent(*this, e, 0)); }

// Seeded device fill (the reference's cuRAND generator is unseeded: gpu-api.cpp:7).  Each call
// advances the stream index so successive fills differ, and the sequence is reproducible.
class Device_RNG {
  struct State {
    gpu::blasHandle_t handle = nullptr;
    uint64_t seed, calls = 0;
    explicit State(uint64_t s) : seed(s) { blas_status_check(gpu::blasCreate(&handle), "Device_RNG: blasCreate"); }
    ~State() { if (handle) gpu::blasDestroy(handle); }
  };
  std::shared_ptr<State> st;

 public:
  explicit Device_RNG(uint64_t seed = 0xB200) : st(std::make_shared<State>(seed)) {}
  Device_RNG(Stream s, uint64_t seed = 0xB200) : Device_RNG(seed) { set_stream(s); }
  void set_stream(Stream s) { gpu::blasSetStream(st->handle, s); }
  void reseed(uint64_t seed) { st->seed = seed; st->calls = 0; }
  template <typename T>
  void uniform(T *x, size_t n) {  // U(0,1], like curandGenerateUniform[Double]
    uint64_t s = st->seed + 0x9E3779B97F4A7C15ull * (st->calls++);
    if constexpr (std::is_same_v<T, double>) blas_status_check(rtat_b200_fill_uniform_d(st->handle, x, n, s, 0.0, 1.0), "fill_uniform_d");
    else blas_status_check(rtat_b200_fill_uniform_s(st->handle, x, n, s, 0.f, 1.f), "fill_uniform_s");
  }
};

// ---- string-backed enums (reference gpu-api.h:246-325) -------------------------------------------
// Traits supply the two (value, name) pairs; operator! flips between them.
template <class Traits>
class String_Rep {
  using T = typename Traits::value_type;
  T val;

 public:
  String_Rep(T v) : val(v) {}
  String_Rep(const std::string &name) {
    for (auto &entry : Traits::table)
      if (name == entry.second) { val = entry.first; return; }
    throw std::runtime_error("Invalid string " + name + " passed to string rep");
  }
  String_Rep(const char *name) : String_Rep(std::string(name)) {}
  operator T() const { return val; }
  operator std::string() const {
    for (auto &entry : Traits::table)
      if (entry.first == val) return entry.second;
    throw std::runtime_error("Invalid string rep value");
  }
  String_Rep operator!() const { return String_Rep(val == Traits::table[0].first ? Traits::table[1].first : Traits::table[0].first); }
  bool operator==(T o) const { return val == o; }
  bool operator!=(T o) const { return val != o; }
  friend std::ostream &operator<<(std::ostream &os, const String_Rep &r) { return os << std::string(r); }
};

struct BLAS_Operation_Traits {
  using value_type = gpu::blasOperation_t;
  static constexpr std::pair<value_type, const char *> table[2] = {{gpu::BLAS_OP_N, "N"}, {gpu::BLAS_OP_T, "T"}};
};
struct BLAS_Fill_Mode_Traits {
  using value_type = gpu::blasFillMode_t;
  static constexpr std::pair<value_type, const char *> table[2] = {{gpu::BLAS_FILL_MODE_LOWER, "Lower"}, {gpu::BLAS_FILL_MODE_UPPER, "Upper"}};
};
struct BLAS_Side_Traits {
  using value_type = gpu::blasSideMode_t;
  static constexpr std::pair<value_type, const char *> table[2] = {{gpu::BLAS_SIDE_LEFT, "Left"}, {gpu::BLAS_SIDE_RIGHT, "Right"}};
};
struct BLAS_Diag_Traits {
  using value_type = gpu::blasDiagType_t;
  static constexpr std::pair<value_type, const char *> table[2] = {{gpu::BLAS_DIAG_UNIT, "Unit"}, {gpu::BLAS_DIAG_NON_UNIT, "Non-Unit"}};
};
using BLAS_Operation = String_Rep<BLAS_Operation_Traits>;
using BLAS_Fill_Mode = String_Rep<BLAS_Fill_Mode_Traits>;
using BLAS_Side = String_Rep<BLAS_Side_Traits>;
using BLAS_Diag = String_Rep<BLAS_Diag_Traits>;

}  // namespace rtat

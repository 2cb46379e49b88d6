// Internal declarations shared by the kernels behind the C ABI (include/rtat_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>
#include <map>
#include "../../include/rtat_b200.h"

struct rtat_b200_context {
  int device = 0;
  int sm_count = 0;
  size_t l2_bytes = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  // library-owned device arena: stream-K partial tiles, pre-split operands, host-entry staging
  void *arena = nullptr;
  size_t arena_bytes = 0;
  // TRSM: inverted 32 x 32 diagonal blocks of the current call (separate from the arena, which the GEMM updates use)
  void *trsm_inv = nullptr;
  size_t trsm_inv_bytes = 0;
  // SGEMM: 16 B-aligned copy of an op(A) TMA cannot address (leading dimension not a multiple of 4 floats, 4 / 8 B base)
  void *sgemm_a_pad = nullptr;
  size_t sgemm_a_pad_bytes = 0;
  // Library-owned scratch (arena, trsm_inv, stream-K flags) is ordered by the stream the work runs on.  When the
  // caller rebinds the handle to another stream while such work may still be queued, set_stream makes the new stream
  // wait for it, so a later call cannot overwrite scratch an earlier one is still reading.
  bool scratch_in_flight = false;
  cudaEvent_t order_event = nullptr;
  // tuning knobs (rtat_b200_set_option)
  std::map<std::string, int, std::less<>> options;  // transparent comparator: lookups by const char* build no std::string
  // introspection
  const char *last_kernel = "none";
  uint64_t launches = 0;
  // stream-K fix-up flags (1024 x u32) and the generation they are compared against
  void *sk_flags = nullptr;
  unsigned sk_generation = 0;
  // skinny DGEMM: arrival counters of the column blocks (zero between launches)
  void *skinny_counters = nullptr;
  // host-entry staging
  void *staging = nullptr;
  size_t staging_bytes = 0;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H side streams of the host entry points
  cudaEvent_t host_events[96] = {};  // entry fence + one per staged piece + one per finished column block
};

namespace rtat_b200 {

// cuTensorMapEncodeTiled resolved through the runtime, so the library has no link-time dependency
// on libcuda (it must load, and fail cleanly in rtat_b200_create, on a box without a driver).
typedef int (*tensor_map_encode_fn)(void *map, int dtype, unsigned rank, void *base, const unsigned long long *dims,
                                    const unsigned long long *strides, const unsigned *box, const unsigned *estrides,
                                    int interleave, int swizzle, int l2promo, int oobfill);
void *tensor_map_encoder_raw();

inline int opt(const rtat_b200_context *h, const char *key, int dflt = 0) {
  auto it = h->options.find(key);
  return it == h->options.end() ? dflt : it->second;
}

inline int check_launch(rtat_b200_context *h, const char *kernel, int nlaunch = 1) {
  cudaError_t e = cudaPeekAtLastError();
  h->last_kernel = kernel;
  h->launches += nlaunch;
  if (e != cudaSuccess) {
    fprintf(stderr, "rtat_b200: launch of %s failed: %s\n", kernel, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_EXECUTION_FAILED;
  }
  return RTAT_B200_STATUS_SUCCESS;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per (kernel, device).  `done` is a per-call-site
// bit mask of device ordinals (static std::atomic<uint64_t> next to the launch); racing threads may both make the
// (idempotent) runtime call.  Devices >= 64 are simply configured on every launch.
template <typename Kernel>
inline int configure_dynamic_smem(Kernel kernel, size_t bytes, int device, std::atomic<uint64_t> &done) {
  const uint64_t bit = device >= 0 && device < 64 ? uint64_t(1) << device : 0;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return RTAT_B200_STATUS_SUCCESS;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_INTERNAL_ERROR;
  }
  if (bit) done.fetch_or(bit, std::memory_order_release);
  return RTAT_B200_STATUS_SUCCESS;
}

// grow-only device arena owned by the handle (contents are scratch; never preserved across calls)
int ensure_arena(rtat_b200_context *h, size_t bytes);

// kernel families (each defined in its own .cu)
template <typename T>
int geam(rtat_b200_context *h, int transa, int transb, int m, int n, T alpha, const T *A, int lda,
         T beta, const T *B, int ldb, T *C, int ldc);

int dgemm(rtat_b200_context *h, int transa, int transb, int m, int n, int k, double alpha,
          const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
int sgemm(rtat_b200_context *h, int transa, int transb, int m, int n, int k, float alpha,
          const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc);

int dsyrk(rtat_b200_context *h, int uplo, int trans, int n, int k, double alpha, const double *A, int lda, double beta,
          double *C, int ldc);
int ssyrk(rtat_b200_context *h, int uplo, int trans, int n, int k, float alpha, const float *A, int lda, float beta,
          float *C, int ldc);

int dtrsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, double alpha, const double *A, int lda, double *B,
          int ldb);
int strsm(rtat_b200_context *h, int side, int uplo, int trans, int diag, int m, int n, float alpha, const float *A, int lda, float *B,
          int ldb);

template <typename T>
int fill_uniform(rtat_b200_context *h, T *x, size_t n, uint64_t seed, T lo, T hi);

}  // namespace rtat_b200

// C ABI entry points (include/rtat_b200.h): handle management, argument validation, dispatch.
// Replaces the cuBLAS side of the reference's rtat::gpu:: alias table (src/gpu-api/gpu-api.h:87-113).
#include <string>
#include <cstdlib>
#include "common.cuh"
#include "host_pipeline.cuh"
#include <cstring>
#include <new>

using namespace rtat_b200;

namespace rtat_b200 {
void *tensor_map_encoder_raw() {
  static void *fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
    }
    return p;
  }();
  return fn;
}

int ensure_arena(rtat_b200_context *h, size_t bytes) {
  h->scratch_in_flight = true;  // the caller is about to queue work that uses the arena
  if (bytes <= h->arena_bytes) return RTAT_B200_STATUS_SUCCESS;
  if (h->arena) {
    // the arena may still be in use by work queued on the handle's stream
    cudaStreamSynchronize(h->stream);
    cudaFree(h->arena);
    h->arena = nullptr;
    h->arena_bytes = 0;
  }
  size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
  if (cudaMalloc(&h->arena, want) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  h->arena_bytes = want;
  return RTAT_B200_STATUS_SUCCESS;
}
}  // namespace rtat_b200

extern "C" {

int rtat_b200_version(void) { return RTAT_B200_VERSION; }

const char *rtat_b200_status_string(int s) {
  switch (s) {
    case RTAT_B200_STATUS_SUCCESS: return "RTAT_B200_STATUS_SUCCESS";
    case RTAT_B200_STATUS_NOT_INITIALIZED: return "RTAT_B200_STATUS_NOT_INITIALIZED";
    case RTAT_B200_STATUS_ALLOC_FAILED: return "RTAT_B200_STATUS_ALLOC_FAILED";
    case RTAT_B200_STATUS_INVALID_VALUE: return "RTAT_B200_STATUS_INVALID_VALUE";
    case RTAT_B200_STATUS_ARCH_MISMATCH: return "RTAT_B200_STATUS_ARCH_MISMATCH";
    case RTAT_B200_STATUS_EXECUTION_FAILED: return "RTAT_B200_STATUS_EXECUTION_FAILED";
    case RTAT_B200_STATUS_INTERNAL_ERROR: return "RTAT_B200_STATUS_INTERNAL_ERROR";
    case RTAT_B200_STATUS_NOT_SUPPORTED: return "RTAT_B200_STATUS_NOT_SUPPORTED";
    default: return "RTAT_B200_STATUS_UNKNOWN";
  }
}

int rtat_b200_create(rtat_b200_handle_t *handle) {
  if (!handle) return RTAT_B200_STATUS_INVALID_VALUE;
  *handle = nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_NOT_INITIALIZED;  // no usable GPU: there is no CPU fallback
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_NOT_INITIALIZED;
  }
  if (prop.major != 10) {
    fprintf(stderr, "rtat_b200: device %d is sm_%d%d; this library only contains sm_100a code\n", dev,
            prop.major, prop.minor);
    return RTAT_B200_STATUS_ARCH_MISMATCH;
  }
  auto *h = new (std::nothrow) rtat_b200_context();
  if (!h) return RTAT_B200_STATUS_ALLOC_FAILED;
  h->device = dev;
  h->sm_count = prop.multiProcessorCount;
  h->l2_bytes = (size_t)prop.l2CacheSize;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->stream = nullptr;  // legacy default stream until set_stream, as with cuBLAS
  // RTAT_B200_OPTIONS="key=value,key=value": option overrides for every handle of the process, so that experiments reach
  // the kernels through any caller (the C++ planner, run_tests, the reference's own host code over this ABI)
  if (const char *env = getenv("RTAT_B200_OPTIONS")) {
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t end = all.find(',', pos);
      if (end == std::string::npos) end = all.size();
      const std::string kv = all.substr(pos, end - pos);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos && eq > 0) h->options[kv.substr(0, eq)] = atoi(kv.c_str() + eq + 1);
      pos = end + 1;
    }
  }
  *handle = h;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_destroy(rtat_b200_handle_t h) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaStreamSynchronize(h->stream);
  if (h->arena) cudaFree(h->arena);
  if (h->staging) cudaFree(h->staging);
  if (h->trsm_inv) cudaFree(h->trsm_inv);
  if (h->sgemm_a_pad) cudaFree(h->sgemm_a_pad);
  if (h->sk_flags) cudaFree(h->sk_flags);
  if (h->skinny_counters) cudaFree(h->skinny_counters);
  if (h->order_event) cudaEventDestroy(h->order_event);
  if (h->copy_in) cudaStreamDestroy(h->copy_in);
  if (h->copy_out) cudaStreamDestroy(h->copy_out);
  for (auto e : h->host_events) if (e) cudaEventDestroy(e);
  delete h;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_set_stream(rtat_b200_handle_t h, void *stream) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaStream_t next = static_cast<cudaStream_t>(stream);
  if (next != h->stream && h->scratch_in_flight) {
    // order the new stream after whatever the old one still has queued against the handle's scratch
    if (!h->order_event && cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_ALLOC_FAILED;
    }
    if (cudaEventRecord(h->order_event, h->stream) != cudaSuccess || cudaStreamWaitEvent(next, h->order_event, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    }
    h->scratch_in_flight = false;
  }
  h->stream = next;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_get_stream(rtat_b200_handle_t h, void **stream) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!stream) return RTAT_B200_STATUS_INVALID_VALUE;
  *stream = static_cast<void *>(h->stream);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_set_option(rtat_b200_handle_t h, const char *key, int value) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!key) return RTAT_B200_STATUS_INVALID_VALUE;
  h->options[key] = value;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_get_option(rtat_b200_handle_t h, const char *key, int *value) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!key || !value) return RTAT_B200_STATUS_INVALID_VALUE;
  *value = opt(h, key, 0);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_host_schedule(int elem_size, int transa, int m, int n, int k, int lda, int world, int opt_panels, int opt_lead, int *out,
                            int cap) {
  if ((elem_size != 4 && elem_size != 8) || m < 0 || n < 0 || k < 0 || lda < 1 || world < 0 || !out) return -1;
  const size_t cols_a = transa == RTAT_B200_OP_N ? k : m;
  const size_t bytes_a = (size_t(elem_size) * lda * cols_a + 511) & ~size_t(511);
  double share = 1.0, rate = 50e9;
  if (world > 0) rtat_b200::mg_link_model(world, share, rate);
  const rtat_b200::HostSchedule sc = rtat_b200::plan_host_schedule(m, n, k, elem_size, bytes_a, share, rate, world > 0, opt_panels, opt_lead);
  const int need = 4 + 2 * sc.blocks + 2 * sc.npanels;
  if (need > cap) return -1;
  int *o = out;
  *o++ = sc.blocks; *o++ = sc.lead; *o++ = sc.npanels; *o++ = sc.nb;
  for (int j = 0; j < sc.blocks; j++) *o++ = sc.col0[j];
  for (int j = 0; j < sc.blocks; j++) *o++ = sc.cols[j];
  for (int p = 0; p < sc.npanels; p++) *o++ = sc.k0s[p];
  for (int p = 0; p < sc.npanels; p++) *o++ = sc.kcs[p];
  return need;
}

const char *rtat_b200_last_kernel(rtat_b200_handle_t h) { return h ? h->last_kernel : "none"; }
uint64_t rtat_b200_launch_count(rtat_b200_handle_t h) { return h ? h->launches : 0; }

// ---- argument validation shared by gemm entry points (cuBLAS rules) ---------------------------
static int validate_gemm(int transa, int transb, int m, int n, int k, int lda, int ldb, int ldc) {
  if ((transa != RTAT_B200_OP_N && transa != RTAT_B200_OP_T) ||
      (transb != RTAT_B200_OP_N && transb != RTAT_B200_OP_T))
    return RTAT_B200_STATUS_INVALID_VALUE;
  if (m < 0 || n < 0 || k < 0) return RTAT_B200_STATUS_INVALID_VALUE;
  int rows_a = transa == RTAT_B200_OP_N ? m : k;
  int rows_b = transb == RTAT_B200_OP_N ? k : n;
  if (lda < (rows_a > 1 ? rows_a : 1) || ldb < (rows_b > 1 ? rows_b : 1) || ldc < (m > 1 ? m : 1))
    return RTAT_B200_STATUS_INVALID_VALUE;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_dgemm(rtat_b200_handle_t h, int transa, int transb, int m, int n, int k,
                    const double *alpha, const double *A, int lda, const double *B, int ldb,
                    const double *beta, double *C, int ldc) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_gemm(transa, transb, m, n, k, lda, ldb, ldc);
  if (st) return st;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!C || (k > 0 && *alpha != 0.0 && (!A || !B))) return RTAT_B200_STATUS_INVALID_VALUE;
  return dgemm(h, transa, transb, m, n, k, *alpha, A, lda, B, ldb, *beta, C, ldc);
}

int rtat_b200_sgemm(rtat_b200_handle_t h, int transa, int transb, int m, int n, int k,
                    const float *alpha, const float *A, int lda, const float *B, int ldb,
                    const float *beta, float *C, int ldc) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_gemm(transa, transb, m, n, k, lda, ldb, ldc);
  if (st) return st;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!C || (k > 0 && *alpha != 0.f && (!A || !B))) return RTAT_B200_STATUS_INVALID_VALUE;
  return sgemm(h, transa, transb, m, n, k, *alpha, A, lda, B, ldb, *beta, C, ldc);
}

static int validate_syrk(int uplo, int trans, int n, int k, int lda, int ldc) {
  if (uplo != RTAT_B200_FILL_LOWER && uplo != RTAT_B200_FILL_UPPER) return RTAT_B200_STATUS_INVALID_VALUE;
  if (trans != RTAT_B200_OP_N && trans != RTAT_B200_OP_T) return RTAT_B200_STATUS_INVALID_VALUE;
  if (n < 0 || k < 0) return RTAT_B200_STATUS_INVALID_VALUE;
  int rows_a = trans == RTAT_B200_OP_N ? n : k;
  if (lda < (rows_a > 1 ? rows_a : 1) || ldc < (n > 1 ? n : 1)) return RTAT_B200_STATUS_INVALID_VALUE;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_dsyrk(rtat_b200_handle_t h, int uplo, int trans, int n, int k, const double *alpha, const double *A,
                    int lda, const double *beta, double *C, int ldc) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_syrk(uplo, trans, n, k, lda, ldc);
  if (st) return st;
  if (n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!C || (k > 0 && *alpha != 0.0 && !A)) return RTAT_B200_STATUS_INVALID_VALUE;
  return dsyrk(h, uplo, trans, n, k, *alpha, A, lda, *beta, C, ldc);
}

int rtat_b200_ssyrk(rtat_b200_handle_t h, int uplo, int trans, int n, int k, const float *alpha, const float *A,
                    int lda, const float *beta, float *C, int ldc) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_syrk(uplo, trans, n, k, lda, ldc);
  if (st) return st;
  if (n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!C || (k > 0 && *alpha != 0.f && !A)) return RTAT_B200_STATUS_INVALID_VALUE;
  return ssyrk(h, uplo, trans, n, k, *alpha, A, lda, *beta, C, ldc);
}

static int validate_trsm(int side, int uplo, int trans, int diag, int m, int n, int lda, int ldb) {
  if (side != RTAT_B200_SIDE_LEFT && side != RTAT_B200_SIDE_RIGHT) return RTAT_B200_STATUS_INVALID_VALUE;
  if (uplo != RTAT_B200_FILL_LOWER && uplo != RTAT_B200_FILL_UPPER) return RTAT_B200_STATUS_INVALID_VALUE;
  if (trans != RTAT_B200_OP_N && trans != RTAT_B200_OP_T) return RTAT_B200_STATUS_INVALID_VALUE;
  if (diag != RTAT_B200_DIAG_NON_UNIT && diag != RTAT_B200_DIAG_UNIT) return RTAT_B200_STATUS_INVALID_VALUE;
  if (m < 0 || n < 0) return RTAT_B200_STATUS_INVALID_VALUE;
  int na = side == RTAT_B200_SIDE_LEFT ? m : n;
  if (lda < (na > 1 ? na : 1) || ldb < (m > 1 ? m : 1)) return RTAT_B200_STATUS_INVALID_VALUE;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_dtrsm(rtat_b200_handle_t h, int side, int uplo, int trans, int diag, int m, int n, const double *alpha,
                    const double *A, int lda, double *B, int ldb) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_trsm(side, uplo, trans, diag, m, n, lda, ldb);
  if (st) return st;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!B || (*alpha != 0.0 && !A)) return RTAT_B200_STATUS_INVALID_VALUE;
  return dtrsm(h, side, uplo, trans, diag, m, n, *alpha, A, lda, B, ldb);
}

int rtat_b200_strsm(rtat_b200_handle_t h, int side, int uplo, int trans, int diag, int m, int n, const float *alpha,
                    const float *A, int lda, float *B, int ldb) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_trsm(side, uplo, trans, diag, m, n, lda, ldb);
  if (st) return st;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!B || (*alpha != 0.f && !A)) return RTAT_B200_STATUS_INVALID_VALUE;
  return strsm(h, side, uplo, trans, diag, m, n, *alpha, A, lda, B, ldb);
}

}  // extern "C"

template <typename T>
static int geam_entry(rtat_b200_handle_t h, int transa, int transb, int m, int n, const T *alpha,
                      const T *A, int lda, const T *beta, const T *B, int ldb, T *C, int ldc) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  if ((transa != RTAT_B200_OP_N && transa != RTAT_B200_OP_T) ||
      (transb != RTAT_B200_OP_N && transb != RTAT_B200_OP_T))
    return RTAT_B200_STATUS_INVALID_VALUE;
  if (m < 0 || n < 0) return RTAT_B200_STATUS_INVALID_VALUE;
  int rows_a = transa == RTAT_B200_OP_N ? m : n;
  int rows_b = transb == RTAT_B200_OP_N ? m : n;
  if (lda < (rows_a > 1 ? rows_a : 1) || ldb < (rows_b > 1 ? rows_b : 1) || ldc < (m > 1 ? m : 1))
    return RTAT_B200_STATUS_INVALID_VALUE;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!C) return RTAT_B200_STATUS_INVALID_VALUE;
  if (*alpha != T(0) && !A) return RTAT_B200_STATUS_INVALID_VALUE;
  if (*beta != T(0) && !B) return RTAT_B200_STATUS_INVALID_VALUE;
  // in-place rules (cuBLAS geam): C may alias an operand only if that operand is not transposed
  // and shares C's leading dimension
  if (*alpha != T(0) && (const T *)C == A && (transa != RTAT_B200_OP_N || lda != ldc))
    return RTAT_B200_STATUS_INVALID_VALUE;
  if (*beta != T(0) && (const T *)C == B && (transb != RTAT_B200_OP_N || ldb != ldc))
    return RTAT_B200_STATUS_INVALID_VALUE;
  return geam<T>(h, transa, transb, m, n, *alpha, A, lda, *beta, B, ldb, C, ldc);
}

extern "C" {

int rtat_b200_dgeam(rtat_b200_handle_t h, int transa, int transb, int m, int n, const double *alpha,
                    const double *A, int lda, const double *beta, const double *B, int ldb,
                    double *C, int ldc) {
  return geam_entry<double>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
}

int rtat_b200_sgeam(rtat_b200_handle_t h, int transa, int transb, int m, int n, const float *alpha,
                    const float *A, int lda, const float *beta, const float *B, int ldb, float *C,
                    int ldc) {
  return geam_entry<float>(h, transa, transb, m, n, alpha, A, lda, beta, B, ldb, C, ldc);
}

int rtat_b200_fill_uniform_d(rtat_b200_handle_t h, double *x, size_t n, uint64_t seed, double lo,
                             double hi) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!x) return RTAT_B200_STATUS_INVALID_VALUE;
  return fill_uniform<double>(h, x, n, seed, lo, hi);
}

int rtat_b200_fill_uniform_s(rtat_b200_handle_t h, float *x, size_t n, uint64_t seed, float lo,
                             float hi) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!x) return RTAT_B200_STATUS_INVALID_VALUE;
  return fill_uniform<float>(h, x, n, seed, lo, hi);
}

}  // extern "C"

// ---- host-buffer entry points (pipeline: host_pipeline.cuh) ---------------------------------------
template <typename T, typename GemmFn>
static int gemm_host(rtat_b200_handle_t h, int transa, int transb, int m, int n, int k,
                     const T *alpha, const T *A, int lda, const T *B, int ldb, const T *beta, T *C,
                     int ldc, GemmFn gemm_fn) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  int st = validate_gemm(transa, transb, m, n, k, lda, ldb, ldc);
  if (st) return st;
  if (m == 0 || n == 0) return RTAT_B200_STATUS_SUCCESS;
  if (!A || !B || !C) return RTAT_B200_STATUS_INVALID_VALUE;
  return gemm_host_pipeline<T>(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gemm_fn, nullptr);
}

extern "C" {

int rtat_b200_dgemm_host(rtat_b200_handle_t h, int transa, int transb, int m, int n, int k,
                         const double *alpha, const double *A, int lda, const double *B, int ldb,
                         const double *beta, double *C, int ldc) {
  return gemm_host<double>(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dgemm);
}

int rtat_b200_sgemm_host(rtat_b200_handle_t h, int transa, int transb, int m, int n, int k,
                         const float *alpha, const float *A, int lda, const float *B, int ldb,
                         const float *beta, float *C, int ldc) {
  return gemm_host<float>(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, sgemm);
}

}  // extern "C"

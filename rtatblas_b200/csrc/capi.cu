// C ABI entry points (include/rtat_b200.h): handle management, argument validation, dispatch.
// Replaces the cuBLAS side of the reference's rtat::gpu:: alias table (src/gpu-api/gpu-api.h:87-113).
#include "common.cuh"
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
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->stream = nullptr;  // legacy default stream until set_stream, as with cuBLAS
  *handle = h;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_destroy(rtat_b200_handle_t h) {
  if (!h) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaStreamSynchronize(h->stream);
  if (h->arena) cudaFree(h->arena);
  if (h->staging) cudaFree(h->staging);
  if (h->trsm_inv) cudaFree(h->trsm_inv);
  if (h->sk_flags) cudaFree(h->sk_flags);
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

// ---- host-buffer entry points ------------------------------------------------------------------
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
  size_t cols_a = transa == RTAT_B200_OP_N ? k : m, cols_b = transb == RTAT_B200_OP_N ? n : k;
  auto round = [](size_t b) { return (b + 511) & ~size_t(511); };
  size_t bytes_a = round(sizeof(T) * lda * cols_a), bytes_b = round(sizeof(T) * ldb * cols_b),
         bytes_c = round(sizeof(T) * size_t(ldc) * n);
  size_t need = bytes_a + bytes_b + bytes_c;
  if (need > h->staging_bytes) {
    cudaStreamSynchronize(h->stream);
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    if (cudaMalloc(&h->staging, need) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_ALLOC_FAILED;
    }
    h->staging_bytes = need;
  }
  char *base = static_cast<char *>(h->staging);
  T *dA = reinterpret_cast<T *>(base), *dB = reinterpret_cast<T *>(base + bytes_a), *dC = reinterpret_cast<T *>(base + bytes_a + bytes_b);
  auto cuda_ok = [](cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    fprintf(stderr, "rtat_b200 host entry: %s: %s\n", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return false;
  };
  // Lazily created side streams / events: H2D, math and D2H run as a column-block pipeline.
  if (!h->copy_in) {
    if (!cuda_ok(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking), "stream") ||
        !cuda_ok(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking), "stream"))
      return RTAT_B200_STATUS_ALLOC_FAILED;
    for (auto &e : h->host_events)
      if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  cudaStream_t s = h->stream, sin = h->copy_in, sout = h->copy_out;
  constexpr int MAX_BLOCKS = 8, MAX_PANELS = 8;
  // Column blocks of op(B) / C: at most 8, at least 256 columns each, multiples of 128.
  int nb = (n + MAX_BLOCKS - 1) / MAX_BLOCKS;
  nb = ((nb < 256 ? 256 : nb) + 127) / 128 * 128;
  int blocks = (n + nb - 1) / nb;
  int col0[MAX_BLOCKS + 2], cols[MAX_BLOCKS + 2];
  for (int j = 0; j < blocks; j++) {
    col0[j] = j * nb;
    cols[j] = (col0[j] + nb > n) ? n - col0[j] : nb;
  }
  // the last block's copy back to the host is the one transfer nothing can hide: halve it
  if (blocks > 1 && cols[blocks - 1] >= 512) {
    const int half = cols[blocks - 1] / 2 / 128 * 128;
    col0[blocks] = col0[blocks - 1] + half;
    cols[blocks] = cols[blocks - 1] - half;
    cols[blocks - 1] = half;
    blocks++;
  }
  // Schedule.  Every column block needs all of A, so a plain column-block pipeline idles the GPU while A crosses
  // PCIe.  Phase 1 therefore runs the first `lead` column blocks K-panel by K-panel: panel p of A and the matching
  // pieces of those blocks are copied, then multiplied (beta, then 1) while panel p + 1 is in flight.  `lead` is
  // chosen so that those blocks' math takes about as long as copying A plus their own share of B.  Phase 2 is the
  // column-block pipeline for the rest (A is resident by then): copy B_j, multiply, copy C_j back.
  int panels = 1, lead = 0;
  if (blocks > 1 && k >= 2048 && bytes_a >= (size_t(32) << 20)) {
    panels = k / 1024 < MAX_PANELS ? k / 1024 : MAX_PANELS;
    const double pcie = 50e9, rate = sizeof(T) == 8 ? 33e12 : 150e12;  // nominal: only the split point depends on them
    const double t_a = double(bytes_a) / pcie, t_b = double(sizeof(T)) * k * nb / pcie, t_mm = 2.0 * m * nb * double(k) / rate;
    lead = t_mm > t_b ? int(t_a / (t_mm - t_b) + 0.5) : blocks;  // in units of nb columns: leading blocks are full
    if (lead < 1) lead = 1;
    // host.panels / host.lead: overrides for experiments (0 = the choice above)
    if (opt(h, "host.panels") > 0) panels = opt(h, "host.panels") < MAX_PANELS ? opt(h, "host.panels") : MAX_PANELS;
    if (opt(h, "host.lead") > 0) lead = opt(h, "host.lead");
    if (lead > blocks) lead = blocks;
  }
  const int kp = panels > 1 ? ((k + panels - 1) / panels + 15) / 16 * 16 : k;
  // K-panel table; the first panel is split in two so that the math starts after half a panel of A has landed
  int k0s[MAX_PANELS + 2], kcs[MAX_PANELS + 2], npanels = 0;
  for (int p0 = 0; p0 < k; p0 += kp) {
    k0s[npanels] = p0;
    kcs[npanels++] = p0 + kp > k ? k - p0 : kp;
  }
  if (panels > 1 && kcs[0] >= 1024) {
    for (int p = npanels; p > 1; p--) { k0s[p] = k0s[p - 1]; kcs[p] = kcs[p - 1]; }
    const int half = kcs[0] / 2 / 16 * 16;
    k0s[1] = half;
    kcs[1] = kcs[0] - half;
    kcs[0] = half;
    npanels++;
  }
  cudaEvent_t *ev = h->host_events;
  int next_event = 1;  // [0] is the entry fence
  // everything queued on the H2D stream so far is visible to the math stream after this.  A failed record / wait would
  // silently drop an ordering edge of the pipeline, so every one is checked.
  auto landed = [&]() {
    cudaEvent_t e = ev[next_event++];
    return cuda_ok(cudaEventRecord(e, sin), "record H2D") && cuda_ok(cudaStreamWaitEvent(s, e, 0), "wait for H2D");
  };
  // stored element (row, col) of a column-major host/device pair -> 2-D copy of a rows x cols window
  auto h2d_window = [&](T *dst, const T *src, int ld, size_t row0, size_t col0, size_t rows, size_t cols) {
    const size_t off = col0 * ld + row0;
    if (rows == (size_t)ld) return cudaMemcpyAsync(dst + off, src + off, sizeof(T) * rows * cols, cudaMemcpyHostToDevice, sin);
    return cudaMemcpy2DAsync(dst + off, sizeof(T) * ld, src + off, sizeof(T) * ld, sizeof(T) * rows, cols, cudaMemcpyHostToDevice, sin);
  };
  auto d2h_block = [&](int j0, int cnt) {  // only the m x cnt window (ldc padding rows on the host are left untouched)
    cudaEvent_t e = ev[next_event++];
    return cuda_ok(cudaEventRecord(e, s), "record GEMM") && cuda_ok(cudaStreamWaitEvent(sout, e, 0), "wait for GEMM") &&
           cuda_ok(cudaMemcpy2DAsync(C + (size_t)j0 * ldc, sizeof(T) * ldc, dC + (size_t)j0 * ldc, sizeof(T) * ldc, sizeof(T) * m, cnt,
                                     cudaMemcpyDeviceToHost, sout), "D2H C");
  };
  if (!cuda_ok(cudaEventRecord(ev[0], s), "fence")) return RTAT_B200_STATUS_EXECUTION_FAILED;
  if (!cuda_ok(cudaStreamWaitEvent(sin, ev[0], 0), "fence")) return RTAT_B200_STATUS_EXECUTION_FAILED;

  // ---- phase 1: K-panels of A against the first `lead` column blocks (panels == 1: all of A, no blocks) ----
  for (int p = 0; p < npanels; p++) {
    const int p0 = k0s[p], kc = kcs[p];
    // A is stored m x k (N) or k x m (T): the panel is a column block or a row block
    cudaError_t e = transa == RTAT_B200_OP_N ? h2d_window(dA, A, lda, 0, p0, m, kc) : h2d_window(dA, A, lda, p0, 0, kc, m);
    if (!cuda_ok(e, "H2D A")) return RTAT_B200_STATUS_EXECUTION_FAILED;
    for (int j = 0; j < lead; j++) {
      const int j0 = col0[j], cnt = cols[j];
      // op(B) block j, K-panel p: B is stored k x n (N) or n x k (T)
      e = transb == RTAT_B200_OP_N ? h2d_window(dB, B, ldb, p0, j0, kc, cnt) : h2d_window(dB, B, ldb, j0, p0, cnt, kc);
      if (!cuda_ok(e, "H2D B")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      if (p == 0 && *beta != T(0) && !cuda_ok(h2d_window(dC, C, ldc, 0, j0, m, cnt), "H2D C")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      if (!landed()) return RTAT_B200_STATUS_EXECUTION_FAILED;
      const T *Ap = transa == RTAT_B200_OP_N ? dA + (size_t)p0 * lda : dA + p0;
      const T *Bjp = transb == RTAT_B200_OP_N ? dB + (size_t)j0 * ldb + p0 : dB + (size_t)p0 * ldb + j0;
      st = gemm_fn(h, transa, transb, m, cnt, kc, *alpha, Ap, lda, Bjp, ldb, p == 0 ? *beta : T(1), dC + (size_t)j0 * ldc, ldc);
      if (st) return st;
      if (p0 + kc >= k && !d2h_block(j0, cnt)) return RTAT_B200_STATUS_EXECUTION_FAILED;
    }
  }
  // ---- phase 2: whole-K column blocks ----
  for (int j = lead; j < blocks; j++) {
    const int j0 = col0[j], cnt = cols[j];
    cudaError_t e = transb == RTAT_B200_OP_N ? h2d_window(dB, B, ldb, 0, j0, k, cnt) : h2d_window(dB, B, ldb, j0, 0, cnt, k);
    if (!cuda_ok(e, "H2D B")) return RTAT_B200_STATUS_EXECUTION_FAILED;
    if (*beta != T(0) && !cuda_ok(h2d_window(dC, C, ldc, 0, j0, m, cnt), "H2D C")) return RTAT_B200_STATUS_EXECUTION_FAILED;
    if (!landed()) return RTAT_B200_STATUS_EXECUTION_FAILED;
    const T *Bj = transb == RTAT_B200_OP_N ? dB + (size_t)j0 * ldb : dB + j0;
    st = gemm_fn(h, transa, transb, m, cnt, k, *alpha, dA, lda, Bj, ldb, *beta, dC + (size_t)j0 * ldc, ldc);
    if (st) return st;
    if (!d2h_block(j0, cnt)) return RTAT_B200_STATUS_EXECUTION_FAILED;
  }
  if (!cuda_ok(cudaStreamSynchronize(sout), "synchronize") || !cuda_ok(cudaStreamSynchronize(s), "synchronize"))
    return RTAT_B200_STATUS_EXECUTION_FAILED;
  return RTAT_B200_STATUS_SUCCESS;
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

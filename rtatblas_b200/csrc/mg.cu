// Column-split multi-GPU DGEMM (include/rtat_b200_mg.h): K-panel pipeline of peer copies and DMMA GEMMs.
#include <vector>
#include "common.cuh"
#include "../../include/rtat_b200_mg.h"

struct rtat_b200_mg_context {
  rtat_b200_context *h = nullptr;
  int rank = 0, world = 1;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> landed;    // panel p has arrived in A_stage
  std::vector<cudaEvent_t> consumed;  // panel p of the previous call has been read by its GEMM
  bool consumed_valid = false;
};

static void ensure_events(rtat_b200_mg_context *mg, int panels) {
  while ((int)mg->landed.size() < panels) {
    cudaEvent_t a, b;
    cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
    mg->landed.push_back(a);
    mg->consumed.push_back(b);
  }
}

extern "C" {

int rtat_b200_mg_partition(int n, int world, int rank, int *first, int *count) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) return RTAT_B200_STATUS_INVALID_VALUE;
  // blocks differ by at most one column; the first n % world ranks take the extra one
  int base = n / world, extra = n % world;
  *count = base + (rank < extra ? 1 : 0);
  *first = rank * base + (rank < extra ? rank : extra);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_panel(int k, int panels, int p, int *first, int *count) {
  if (k < 0 || panels < 1 || p < 0 || p >= panels || !first || !count) return RTAT_B200_STATUS_INVALID_VALUE;
  // panel width rounded up to the GEMM's k-tile (16) so that only the last panel is ragged
  long long w = ((long long)k + panels - 1) / panels;
  w = (w + 15) / 16 * 16;
  long long f = w * p;
  if (f > k) f = k;
  long long c = (f + w > k) ? k - f : w;
  *first = (int)f;
  *count = (int)c;
  return RTAT_B200_STATUS_SUCCESS;
}

// Automatic schedule (panels == 0 in rtat_b200_mg_dgemm): four panels of 1/16, 3/16, 4/16 and 8/16 of K.
// Only the first panel's transfer is exposed, so it is small; every further panel costs one extra
// read-modify-write of C_loc (the panel GEMMs accumulate with beta = 1), so there are few of them.
int rtat_b200_mg_auto_panel(int k, int p, int *first, int *count) {
  if (k < 0 || p < 0 || p >= 4 || !first || !count) return RTAT_B200_STATUS_INVALID_VALUE;
  static const int sixteenths[5] = {0, 1, 4, 8, 16};
  auto cut = [&](int i) {
    long long c = ((long long)k * sixteenths[i] / 16 + 15) / 16 * 16;
    return (int)(c > k || i == 4 ? k : c);
  };
  *first = cut(p);
  *count = cut(p + 1) - cut(p);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_create(rtat_b200_mg_t *out, rtat_b200_handle_t handle, int rank, int world) {
  if (!out || !handle || world < 1 || rank < 0 || rank >= world) return RTAT_B200_STATUS_INVALID_VALUE;
  auto *mg = new rtat_b200_mg_context();
  mg->h = handle;
  mg->rank = rank;
  mg->world = world;
  if (cudaStreamCreateWithFlags(&mg->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    (void)cudaGetLastError();
    delete mg;
    return RTAT_B200_STATUS_NOT_INITIALIZED;
  }
  *out = mg;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_destroy(rtat_b200_mg_t mg) {
  if (!mg) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaStreamSynchronize(mg->copy_stream);
  for (auto e : mg->landed) cudaEventDestroy(e);
  for (auto e : mg->consumed) cudaEventDestroy(e);
  cudaStreamDestroy(mg->copy_stream);
  delete mg;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_alloc_shared(rtat_b200_mg_t mg, size_t bytes, void **dptr, unsigned char handle_out[RTAT_B200_MG_HANDLE_BYTES]) {
  if (!mg || !dptr || !handle_out) return RTAT_B200_STATUS_INVALID_VALUE;
  static_assert(sizeof(cudaIpcMemHandle_t) == RTAT_B200_MG_HANDLE_BYTES, "IPC handle size");
  if (cudaMalloc(dptr, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  cudaIpcMemHandle_t hnd;
  if (cudaIpcGetMemHandle(&hnd, *dptr) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(*dptr);
    *dptr = nullptr;
    return RTAT_B200_STATUS_NOT_SUPPORTED;
  }
  memcpy(handle_out, &hnd, sizeof hnd);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_free_shared(rtat_b200_mg_t mg, void *dptr) {
  if (!mg) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaDeviceSynchronize();
  return cudaFree(dptr) == cudaSuccess ? RTAT_B200_STATUS_SUCCESS : RTAT_B200_STATUS_EXECUTION_FAILED;
}

int rtat_b200_mg_open_shared(rtat_b200_mg_t mg, const unsigned char handle[RTAT_B200_MG_HANDLE_BYTES], void **peer_dptr) {
  if (!mg || !handle || !peer_dptr) return RTAT_B200_STATUS_INVALID_VALUE;
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, handle, sizeof hnd);
  if (cudaIpcOpenMemHandle(peer_dptr, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    fprintf(stderr, "rtat_b200_mg: cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return RTAT_B200_STATUS_NOT_SUPPORTED;
  }
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_close_shared(rtat_b200_mg_t mg, void *peer_dptr) {
  if (!mg) return RTAT_B200_STATUS_NOT_INITIALIZED;
  cudaStreamSynchronize(mg->copy_stream);
  return cudaIpcCloseMemHandle(peer_dptr) == cudaSuccess ? RTAT_B200_STATUS_SUCCESS : RTAT_B200_STATUS_EXECUTION_FAILED;
}

int rtat_b200_mg_dgemm(rtat_b200_mg_t mg, int transa, int transb, int m, int n_loc, int k, const double *alpha, const double *A_src,
                       int lda, double *A_stage, const double *B_loc, int ldb, const double *beta, double *C_loc, int ldc, int panels) {
  if (!mg || !alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  if (panels < 0) return RTAT_B200_STATUS_INVALID_VALUE;
  const bool auto_schedule = panels == 0;
  if (auto_schedule) panels = 4;
  rtat_b200_context *h = mg->h;
  const bool local = (A_stage == nullptr || A_stage == A_src);
  if (local || k == 0 || m == 0 || n_loc == 0)  // nothing to replicate: one GEMM on the resident operand
    return rtat_b200_dgemm(h, transa, transb, m, n_loc, k, alpha, A_src, lda, B_loc, ldb, beta, C_loc, ldc);
  if (!auto_schedule && panels > (k + 15) / 16) panels = (k + 15) / 16;
  ensure_events(mg, panels);
  const double one = 1.0;
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  bool first_gemm = true;
  for (int p = 0; p < panels; p++) {
    int k0, kw;
    if (auto_schedule) rtat_b200_mg_auto_panel(k, p, &k0, &kw);
    else rtat_b200_mg_panel(k, panels, p, &k0, &kw);
    if (kw <= 0) continue;
    // ---- copy engine: panel p of A, root -> local stage (waits until last call's GEMM has read the slot)
    // a failed record / wait would silently drop an ordering edge of the pipeline, so every one is checked
    auto sync_ok = [](cudaError_t err, const char *what) {
      if (err == cudaSuccess) return true;
      fprintf(stderr, "rtat_b200_mg: %s failed: %s\n", what, cudaGetErrorString(err));
      (void)cudaGetLastError();
      return false;
    };
    if (mg->consumed_valid && !sync_ok(cudaStreamWaitEvent(mg->copy_stream, mg->consumed[p], 0), "wait for the slot"))
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    cudaError_t e;
    if (!ta) {
      // A is m x k: panel = columns [k0, k0 + kw), lda*kw contiguous elements when lda == m, else a strided block
      e = cudaMemcpy2DAsync(A_stage + (size_t)k0 * lda, sizeof(double) * lda, A_src + (size_t)k0 * lda, sizeof(double) * lda,
                            sizeof(double) * m, kw, cudaMemcpyDeviceToDevice, mg->copy_stream);
    } else {
      // A is k x m: panel = rows [k0, k0 + kw) of every column
      e = cudaMemcpy2DAsync(A_stage + k0, sizeof(double) * lda, A_src + k0, sizeof(double) * lda, sizeof(double) * kw, m,
                            cudaMemcpyDeviceToDevice, mg->copy_stream);
    }
    if (e != cudaSuccess) {
      fprintf(stderr, "rtat_b200_mg: panel copy failed: %s\n", cudaGetErrorString(e));
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    }
    // ---- math: C += alpha * op(A)[:, panel] * op(B)[panel, :]
    if (!sync_ok(cudaEventRecord(mg->landed[p], mg->copy_stream), "record panel landed") ||
        !sync_ok(cudaStreamWaitEvent(h->stream, mg->landed[p], 0), "wait for the panel"))
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    const double *Ap = ta ? A_stage + k0 : A_stage + (size_t)k0 * lda;
    const double *Bp = tb ? B_loc + (size_t)k0 * ldb : B_loc + k0;
    int st = rtat_b200_dgemm(h, transa, transb, m, n_loc, kw, alpha, Ap, lda, Bp, ldb, first_gemm ? beta : &one, C_loc, ldc);
    first_gemm = false;
    if (st) return st;
    if (!sync_ok(cudaEventRecord(mg->consumed[p], h->stream), "record panel consumed")) return RTAT_B200_STATUS_EXECUTION_FAILED;
  }
  mg->consumed_valid = true;
  return RTAT_B200_STATUS_SUCCESS;
}

}  // extern "C"

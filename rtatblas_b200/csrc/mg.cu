// Column-split multi-GPU DGEMM (include/rtat_b200_mg.h): K-panel pipeline of peer copies and DMMA GEMMs.
#include <vector>
#include "common.cuh"
#include "host_pipeline.cuh"
#include "../../include/rtat_b200_mg.h"

using rtat_b200::HostPipelineHooks;
using rtat_b200::host_cuda_ok;
using rtat_b200::opt;

struct rtat_b200_mg_context {
  rtat_b200_context *h = nullptr;
  int rank = 0, world = 1;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> landed;    // panel p has arrived in A_stage
  std::vector<cudaEvent_t> consumed;  // panel p of the previous call has been read by its GEMM
  bool consumed_valid = false;
  // What the previous rtat_b200_mg_dgemm call staged where: the per-panel `consumed` events order the reuse of A_stage only
  // when this call cuts the same bytes into the same panels; any other call waits for ALL of the previous call's GEMMs.
  struct StageSignature {
    const void *stage = nullptr;
    int transa = -1, m = 0, k = 0, lda = 0, panels = 0;
    bool auto_schedule = false;
    bool operator==(const StageSignature &o) const {
      return stage == o.stage && transa == o.transa && m == o.m && k == o.k && lda == o.lda && panels == o.panels && auto_schedule == o.auto_schedule;
    }
  } last_signature;
  cudaEvent_t last_consumed = nullptr;  // recorded on the math stream after the previous call's last GEMM
  // ---- host-buffer entry (rtat_b200_mg_dgemm_host): every rank owns a replica of A in IPC-exportable memory, preceded by
  // a flag page that the peers write into
  char *replica = nullptr;                       // own allocation: [flag page][A]
  size_t replica_bytes = 0;                      // bytes of A it can hold
  char *peer[RTAT_B200_MG_MAX_WORLD] = {};       // the other ranks' allocations mapped here (own entry = replica)
  cudaStream_t pull_stream = nullptr;            // peer pulls (copy engines) and the flag waits in front of them
  cudaEvent_t pulled = nullptr;
  int signal_mode = 1;                           // flag transport (mg.signal), fixed by _host_setup: who writes which page
  unsigned generation = 0;                       // call counter: flags carry it, so they never need clearing
  unsigned *generation_pinned = nullptr;         // pinned host word the generation is uploaded from
};

// Flag page layout (u32 words).  arrived[q][p]: rank q's share of K-panel p of call `generation` is in q's replica.
// done[q]: rank q has finished reading this rank's replica for call `generation`.  status: a wait timed out.
namespace {
constexpr size_t FLAG_PAGE_BYTES = 8192;
constexpr int MAX_PANEL_FLAGS = 64;
__host__ __device__ inline size_t arrived_word(int q, int p) { return size_t(q) * MAX_PANEL_FLAGS + p; }
__host__ __device__ inline size_t done_word(int q) { return 1024 + q; }
constexpr size_t STATUS_WORD = 1024 + 64;
constexpr size_t GEN_WORD = 1024 + 65;  // this rank's current generation, the source of the 4-byte flag copies

// one thread: spin until *flag >= value (wrap-safe); gives up after 30 s and raises *status instead of hanging the GPU.
// Fallback for drivers without stream memory operations (see wait_flag).
__global__ void mg_wait_kernel(const unsigned *flag, unsigned value, unsigned *status) {
  unsigned long long t0, now;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
  // one timed-out wait fails the call: the waits queued behind it give up at once instead of each spending its own 30 s
  if (*reinterpret_cast<volatile unsigned *>(status)) return;
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(flag) : "memory");
    if ((int)(v - value) >= 0) return;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
    if (now - t0 > 30000000000ull) {
      atomicExch(status, 1u);
      return;
    }
    __nanosleep(500);
  }
}

// cuStreamWaitValue32 through the runtime (the library has no link-time dependency on libcuda)
typedef int (*stream_wait_value32_fn)(cudaStream_t, unsigned long long addr, unsigned value, unsigned flags);
stream_wait_value32_fn stream_wait_value32() {
  static stream_wait_value32_fn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<stream_wait_value32_fn>(p);
  }();
  return fn;
}
}  // namespace

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
  if (mg->last_consumed) cudaEventDestroy(mg->last_consumed);
  cudaStreamDestroy(mg->copy_stream);
  rtat_b200_mg_host_teardown(mg);
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
  // validated here, before the first panel copy is queued (the GEMM entry would only catch it after the copy engine had
  // been told to move lda * k elements)
  if ((transa != RTAT_B200_OP_N && transa != RTAT_B200_OP_T) || (transb != RTAT_B200_OP_N && transb != RTAT_B200_OP_T) || m < 0 || n_loc < 0 ||
      k < 0)
    return RTAT_B200_STATUS_INVALID_VALUE;
  {
    const int rows_a = transa == RTAT_B200_OP_N ? m : k, rows_b = transb == RTAT_B200_OP_N ? k : n_loc;
    if (lda < (rows_a > 1 ? rows_a : 1) || ldb < (rows_b > 1 ? rows_b : 1) || ldc < (m > 1 ? m : 1)) return RTAT_B200_STATUS_INVALID_VALUE;
  }
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
  // Reuse of A_stage: with the schedule of the previous call, panel p may be overwritten as soon as ITS previous GEMM is done
  // (so the next step's first panel streams in under this step's last GEMMs); with any other schedule a new panel can cover
  // bytes that a different, later panel of the previous call is still being read from, so the copy stream waits for all of it.
  rtat_b200_mg_context::StageSignature sig;
  sig.stage = A_stage; sig.transa = transa; sig.m = m; sig.k = k; sig.lda = lda; sig.panels = panels; sig.auto_schedule = auto_schedule;
  const bool same_schedule = mg->consumed_valid && sig == mg->last_signature;
  if (mg->consumed_valid && !same_schedule) {
    if (cudaStreamWaitEvent(mg->copy_stream, mg->last_consumed, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    }
  }
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
    if (same_schedule && !sync_ok(cudaStreamWaitEvent(mg->copy_stream, mg->consumed[p], 0), "wait for the slot"))
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
  if (!mg->last_consumed && cudaEventCreateWithFlags(&mg->last_consumed, cudaEventDisableTiming) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  if (cudaEventRecord(mg->last_consumed, h->stream) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_EXECUTION_FAILED;
  }
  mg->last_signature = sig;
  mg->consumed_valid = true;
  return RTAT_B200_STATUS_SUCCESS;
}


// ---- host-buffer column split ------------------------------------------------------------------------------------
int rtat_b200_mg_host_setup(rtat_b200_mg_t mg, size_t bytes_a, unsigned char handle_out[RTAT_B200_MG_HANDLE_BYTES]) {
  if (!mg || !handle_out || mg->world > RTAT_B200_MG_MAX_WORLD) return RTAT_B200_STATUS_INVALID_VALUE;
  if (mg->replica) return RTAT_B200_STATUS_INVALID_VALUE;  // one setup per context
  void *p = nullptr;
  if (cudaMalloc(&p, FLAG_PAGE_BYTES + bytes_a) != cudaSuccess) {
    (void)cudaGetLastError();
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  if (cudaMemset(p, 0, FLAG_PAGE_BYTES) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p);
    return RTAT_B200_STATUS_EXECUTION_FAILED;
  }
  cudaIpcMemHandle_t hnd;
  if (cudaIpcGetMemHandle(&hnd, p) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return RTAT_B200_STATUS_NOT_SUPPORTED;
  }
  memcpy(handle_out, &hnd, sizeof hnd);
  if (cudaStreamCreateWithFlags(&mg->pull_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&mg->pulled, cudaEventDisableTiming) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  if (cudaMallocHost(&mg->generation_pinned, 64) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  mg->replica = static_cast<char *>(p);
  mg->replica_bytes = bytes_a;
  mg->peer[mg->rank] = mg->replica;
  mg->generation = 0;
  mg->signal_mode = opt(mg->h, "mg.signal", 1);
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_host_connect(rtat_b200_mg_t mg, const unsigned char *handles) {
  if (!mg || !mg->replica || (mg->world > 1 && !handles)) return RTAT_B200_STATUS_INVALID_VALUE;
  for (int q = 0; q < mg->world; q++) {
    if (q == mg->rank || mg->peer[q]) continue;
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handles + (size_t)q * RTAT_B200_MG_HANDLE_BYTES, sizeof hnd);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      fprintf(stderr, "rtat_b200_mg: cudaIpcOpenMemHandle(rank %d) failed: %s\n", q, cudaGetErrorString(cudaGetLastError()));
      return RTAT_B200_STATUS_NOT_SUPPORTED;
    }
    mg->peer[q] = static_cast<char *>(p);
  }
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_host_teardown(rtat_b200_mg_t mg) {
  if (!mg) return RTAT_B200_STATUS_NOT_INITIALIZED;
  if (!mg->replica) return RTAT_B200_STATUS_SUCCESS;
  cudaDeviceSynchronize();
  for (int q = 0; q < mg->world; q++)
    if (q != mg->rank && mg->peer[q]) cudaIpcCloseMemHandle(mg->peer[q]);
  for (auto &p : mg->peer) p = nullptr;
  cudaFree(mg->replica);
  mg->replica = nullptr;
  if (mg->pulled) cudaEventDestroy(mg->pulled);
  if (mg->pull_stream) cudaStreamDestroy(mg->pull_stream);
  if (mg->generation_pinned) cudaFreeHost(mg->generation_pinned);
  mg->generation_pinned = nullptr;
  mg->pulled = nullptr;
  mg->pull_stream = nullptr;
  return RTAT_B200_STATUS_SUCCESS;
}

int rtat_b200_mg_dgemm_host(rtat_b200_mg_t mg, int transa, int transb, int m, int n_loc, int k, const double *alpha, const double *A,
                            int lda, const double *B_loc, int ldb, const double *beta, double *C_loc, int ldc) {
  if (!mg || !alpha || !beta) return RTAT_B200_STATUS_INVALID_VALUE;
  if ((transa != RTAT_B200_OP_N && transa != RTAT_B200_OP_T) || (transb != RTAT_B200_OP_N && transb != RTAT_B200_OP_T) || m < 0 ||
      n_loc < 0 || k < 0)
    return RTAT_B200_STATUS_INVALID_VALUE;
  const bool ta = transa == RTAT_B200_OP_T, tb = transb == RTAT_B200_OP_T;
  const int rows_a = ta ? k : m, rows_b = tb ? n_loc : k;
  if (lda < (rows_a > 1 ? rows_a : 1) || ldb < (rows_b > 1 ? rows_b : 1) || ldc < (m > 1 ? m : 1)) return RTAT_B200_STATUS_INVALID_VALUE;
  if (!mg->replica) return RTAT_B200_STATUS_NOT_INITIALIZED;
  const size_t cols_a = ta ? m : k;
  if (sizeof(double) * lda * cols_a > mg->replica_bytes) return RTAT_B200_STATUS_INVALID_VALUE;
  for (int q = 0; q < mg->world; q++)
    if (!mg->peer[q]) return RTAT_B200_STATUS_NOT_INITIALIZED;  // host_connect has not run
  if (m == 0) return RTAT_B200_STATUS_SUCCESS;  // the same on every rank: nobody signals, nobody waits
  // every rank carries its share of A even if its own column block were empty; keep the contract simple instead
  if (n_loc == 0 || !A || !B_loc || !C_loc) return RTAT_B200_STATUS_INVALID_VALUE;
  rtat_b200_context *h = mg->h;
  const int R = mg->rank, W = mg->world;
  // The call is collective (every rank makes it, the same number of times): the flags carry the call's generation number,
  // so they never need clearing and a stale value can never satisfy a wait.
  const unsigned gen = ++mg->generation;
  double *dA = reinterpret_cast<double *>(mg->replica + FLAG_PAGE_BYTES);
  unsigned *my_flags = reinterpret_cast<unsigned *>(mg->replica);
  // How a flag travels (mg.signal):
  //   1 (default) — the writer sets the word in its OWN page with a 4-byte H2D copy on the stream that carried the data (same
  //       copy engine as the upload, so strictly behind it), and a reader polls the writer's page over NVLink with a one-thread
  //       kernel.  Every cross-GPU access is a read by the side that needs the data; nothing is pushed.
  //   0 — the writer pushes the word into every peer's page with 4-byte peer copies and a reader waits on its own page with a
  //       stream memory operation (cuStreamWaitValue32; mg.spin_wait = 1: the polling kernel).  Measured on 2 GPUs: the pushed
  //       words land only when the writer's upload stream has drained (~150 ms late), which serialises the whole exchange.
  const int signal_mode = mg->signal_mode;
  const bool spin = opt(h, "mg.spin_wait") != 0 || !stream_wait_value32();
  // wait on `stream` until word `word` written by rank q reaches `value`
  auto wait_flag = [&](cudaStream_t stream, int q, size_t word, unsigned value) {
    unsigned *flag = (signal_mode == 1 ? reinterpret_cast<unsigned *>(mg->peer[q]) : my_flags) + word;
    if (signal_mode != 1 && !spin) {
      int rc = stream_wait_value32()(stream, (unsigned long long)(uintptr_t)flag, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */);
      if (rc == 0) return true;
      fprintf(stderr, "rtat_b200_mg: cuStreamWaitValue32 failed (%d)\n", rc);
      return false;
    }
    mg_wait_kernel<<<1, 1, 0, stream>>>(flag, value, my_flags + STATUS_WORD);
    return host_cuda_ok(cudaPeekAtLastError(), "flag wait kernel");
  };
  // publish `gen` as word `word` (one of this rank's words) in `stream` order
  auto signal_peers = [&](cudaStream_t stream, size_t word) {
    if (signal_mode == 1)
      return host_cuda_ok(cudaMemcpyAsync(my_flags + word, &mg->generation_pinned[0], sizeof(unsigned), cudaMemcpyHostToDevice, stream),
                          "flag write");
    for (int q = 0; q < W; q++) {
      if (q == R) continue;
      if (!host_cuda_ok(cudaMemcpyAsync(reinterpret_cast<unsigned *>(mg->peer[q]) + word, my_flags + GEN_WORD, sizeof(unsigned),
                                        cudaMemcpyDeviceToDevice, stream), "flag write"))
        return false;
    }
    return true;
  };
  int panel = 0;
  bool started = false;
  HostPipelineHooks<double> hooks;
  hooks.dA = dA;
  rtat_b200::mg_link_model(W, hooks.a_share, hooks.pcie_rate);
  hooks.fetch_a = [&](size_t row0, size_t col0, size_t rows, size_t cols) {
    cudaStream_t sin = h->copy_in, spull = mg->pull_stream;
    const int p = panel++;
    if (p >= MAX_PANEL_FLAGS) return false;
    if (!started && W > 1) {
      started = true;
      // this call's generation number, where the copy engines can read it from
      if (!host_cuda_ok(cudaMemcpyAsync(my_flags + GEN_WORD, &mg->generation_pinned[0], sizeof(unsigned), cudaMemcpyHostToDevice, sin),
                        "generation word"))
        return false;
      // this rank's replica is about to be overwritten: every peer must have finished reading last call's contents
      if (gen > 1)
        for (int q = 0; q < W; q++)
          if (q != R && !wait_flag(sin, q, done_word(q), gen - 1)) return false;
    }
    auto window = [&](int q, size_t &c0, size_t &cn) {
      int f, c;
      rtat_b200_mg_partition((int)cols, W, q, &f, &c);
      c0 = col0 + f;
      cn = c;
    };
    size_t c0, cn;
    window(R, c0, cn);
    if (cn > 0) {
      const size_t off = c0 * lda + row0;
      cudaError_t e = rows == (size_t)lda
                          ? cudaMemcpyAsync(dA + off, A + off, sizeof(double) * rows * cn, cudaMemcpyHostToDevice, sin)
                          : cudaMemcpy2DAsync(dA + off, sizeof(double) * lda, A + off, sizeof(double) * lda, sizeof(double) * rows, cn,
                                              cudaMemcpyHostToDevice, sin);
      if (!host_cuda_ok(e, "H2D share of A")) return false;
    }
    if (W > 1) {
      if (!signal_peers(sin, arrived_word(R, p))) return false;
      for (int i = 1; i < W; i++) {
        const int q = (R + i) % W;  // every rank starts with a different peer
        window(q, c0, cn);
        if (!wait_flag(spull, q, arrived_word(q, p), gen)) return false;
        if (cn == 0) continue;
        const size_t off = c0 * lda + row0;
        const double *src = reinterpret_cast<const double *>(mg->peer[q] + FLAG_PAGE_BYTES) + off;
        if (!host_cuda_ok(cudaMemcpy2DAsync(dA + off, sizeof(double) * lda, src, sizeof(double) * lda, sizeof(double) * rows, cn,
                                            cudaMemcpyDeviceToDevice, spull), "peer pull of A"))
          return false;
        if (hooks.mark) hooks.mark(spull, "pull", p, q);
      }
    }
    return true;
  };
  hooks.fence_a = [&](cudaStream_t s) {
    if (W == 1) return true;
    return host_cuda_ok(cudaEventRecord(mg->pulled, mg->pull_stream), "record pulls") &&
           host_cuda_ok(cudaStreamWaitEvent(s, mg->pulled, 0), "wait for pulls");
  };
  mg->generation_pinned[0] = gen;
  auto gemm_fn = [](rtat_b200_context *hh, int ta_, int tb_, int mm, int nn, int kk, double al, const double *a, int la, const double *b,
                    int lb, double be, double *c, int lc) { return rtat_b200::dgemm(hh, ta_, tb_, mm, nn, kk, al, a, la, b, lb, be, c, lc); };
  int st = rtat_b200::gemm_host_pipeline<double>(h, transa, transb, m, n_loc, k, alpha, A, lda, B_loc, ldb, beta, C_loc, ldc, gemm_fn, &hooks);
  if (W > 1) {
    // pushed flags (mg.signal = 0) are device-to-device copies OF the generation word: (re)write it on the stream that is
    // about to read it, so the read cannot overtake the upload that copy_in carried at the start of the call
    if ((!started || signal_mode != 1) &&
        !host_cuda_ok(cudaMemcpyAsync(my_flags + GEN_WORD, &mg->generation_pinned[0], sizeof(unsigned), cudaMemcpyHostToDevice,
                                      mg->pull_stream), "generation word"))
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    // tell every peer that this rank no longer reads their replicas of this generation (all pulls are behind us on pull_stream)
    if (!signal_peers(mg->pull_stream, done_word(R))) return RTAT_B200_STATUS_EXECUTION_FAILED;
    unsigned status = 0;
    if (!host_cuda_ok(cudaStreamSynchronize(mg->pull_stream), "synchronize pulls") ||
        !host_cuda_ok(cudaStreamSynchronize(h->copy_in), "synchronize H2D") ||
        !host_cuda_ok(cudaMemcpy(&status, my_flags + STATUS_WORD, sizeof status, cudaMemcpyDeviceToHost), "read status"))
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    if (status) {
      // report once: the word is cleared so that a later, healthy call on this context is not failed by this one
      (void)cudaMemset(my_flags + STATUS_WORD, 0, sizeof(unsigned));
      fprintf(stderr, "rtat_b200_mg: rank %d timed out waiting for a peer's share of A (is every rank making this call?)\n", R);
      return RTAT_B200_STATUS_EXECUTION_FAILED;
    }
  }
  return st;
}

}  // extern "C"

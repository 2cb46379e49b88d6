// Host-buffer GEMM pipeline shared by the single-GPU host entry points (capi.cu: rtat_b200_{d,s}gemm_host) and the
// multi-GPU one (mg.cu: rtat_b200_mg_dgemm_host).  A, B, C live in (pinned) host memory; H2D copies, GEMMs and D2H copies
// run as a pipeline on three streams so that only the first piece in and the last piece out are exposed.
//
// What differs between the callers is how a K-panel of A reaches the device: the single-GPU entry copies it over PCIe;
// the multi-GPU entry copies only this rank's 1/world share over PCIe and pulls the other shares from the peers' replicas
// over NVLink (copy engines).  That step is the `fetch_a` hook.
#pragma once
#include <functional>
#include <string>
#include <vector>
#include "common.cuh"
#include "host_schedule.h"

namespace rtat_b200 {

template <typename T>
struct HostPipelineHooks {
  // Bring the stored window rows [row0, row0 + rows) x columns [col0, col0 + cols) of A (leading dimension lda) into
  // dA at the same offsets.  Work that the plain H2D stream `sin` carries is fenced by the pipeline itself; anything
  // queued elsewhere must be covered by `fence_a`.  Null: plain H2D copy of the window.
  std::function<bool(size_t row0, size_t col0, size_t rows, size_t cols)> fetch_a;
  // Make `s` wait for whatever fetch_a queued outside `sin` (peer pulls).  Null: nothing.
  std::function<bool(cudaStream_t s)> fence_a;
  // set by the pipeline when host.trace is on: timestamp `stream` under a label (what, panel, index)
  std::function<void(cudaStream_t stream, const char *what, int a, int b)> mark;
  T *dA = nullptr;          // device home of A (null: the handle's staging)
  double a_share = 1.0;     // fraction of A that crosses this GPU's PCIe link
  double pcie_rate = 50e9;  // bytes/s this GPU's link can expect (only the split point of the schedule depends on it)
};

inline bool host_cuda_ok(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  fprintf(stderr, "rtat_b200 host entry: %s: %s\n", what, cudaGetErrorString(e));
  (void)cudaGetLastError();
  return false;
}

// Arguments are validated by the caller.  gemm_fn(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc) is the
// device GEMM (value alpha / beta).  Returns after C has landed in host memory.
template <typename T, typename GemmFn>
int gemm_host_pipeline(rtat_b200_context *h, int transa, int transb, int m, int n, int k, const T *alpha, const T *A, int lda,
                       const T *B, int ldb, const T *beta, T *C, int ldc, GemmFn gemm_fn, HostPipelineHooks<T> *hooks) {
  auto cuda_ok = host_cuda_ok;
  size_t cols_a = transa == RTAT_B200_OP_N ? k : m, cols_b = transb == RTAT_B200_OP_N ? n : k;
  auto round = [](size_t b) { return (b + 511) & ~size_t(511); };
  const bool own_a = !(hooks && hooks->dA);
  size_t bytes_a = round(sizeof(T) * lda * cols_a), bytes_b = round(sizeof(T) * ldb * cols_b),
         bytes_c = round(sizeof(T) * size_t(ldc) * n);
  size_t need = (own_a ? bytes_a : 0) + bytes_b + bytes_c;
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
  T *dA = own_a ? reinterpret_cast<T *>(base) : hooks->dA;
  T *dB = reinterpret_cast<T *>(base + (own_a ? bytes_a : 0)), *dC = reinterpret_cast<T *>(base + (own_a ? bytes_a : 0) + bytes_b);
  // Lazily created side streams / events: H2D, math and D2H run as a column-block pipeline.
  if (!h->copy_in) {
    if (!cuda_ok(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking), "stream") ||
        !cuda_ok(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking), "stream"))
      return RTAT_B200_STATUS_ALLOC_FAILED;
    for (auto &e : h->host_events)
      if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return RTAT_B200_STATUS_ALLOC_FAILED;
  }
  cudaStream_t s = h->stream, sin = h->copy_in, sout = h->copy_out;
  // Schedule (host_schedule.h): column blocks, K-panels and the number of leading blocks that take the K-panel phase.
  const HostSchedule sc = plan_host_schedule(m, n, k, sizeof(T), bytes_a, hooks ? hooks->a_share : 1.0, hooks ? hooks->pcie_rate : 50e9,
                                             hooks != nullptr, opt(h, "host.panels"), opt(h, "host.lead"));
  const int blocks = sc.blocks, lead = sc.lead, npanels = sc.npanels;
  const int *col0 = sc.col0, *cols = sc.cols, *k0s = sc.k0s, *kcs = sc.kcs;
  cudaEvent_t *ev = h->host_events;
  int next_event = 1;  // [0] is the entry fence
  // host.trace = 1: timestamp every stage of the pipeline (timed events on the three streams) and print the timeline to
  // stderr when the call ends — how the schedule was tuned and how the multi-GPU exchange is debugged
  const bool trace = opt(h, "host.trace") != 0;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  auto mark = [&](cudaStream_t st_, const char *what, int a, int b) {
    if (!trace) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st_);
    char buf[64];
    snprintf(buf, sizeof buf, "%s p%d j%d", what, a, b);
    marks.emplace_back(buf, e);
  };
  mark(s, "start", 0, 0);
  if (hooks && trace) hooks->mark = mark;
  // everything queued on the H2D stream so far is visible to the math stream after this.  A failed record / wait would
  // silently drop an ordering edge of the pipeline, so every one is checked.
  auto landed = [&]() {
    cudaEvent_t e = ev[next_event++];
    if (!(cuda_ok(cudaEventRecord(e, sin), "record H2D") && cuda_ok(cudaStreamWaitEvent(s, e, 0), "wait for H2D"))) return false;
    return !(hooks && hooks->fence_a) || hooks->fence_a(s);
  };
  // stored element (row, col) of a column-major host/device pair -> 2-D copy of a rows x cols window
  auto h2d_window = [&](T *dst, const T *src, int ld, size_t row0, size_t col0, size_t rows, size_t cols) {
    const size_t off = col0 * ld + row0;
    if (rows == (size_t)ld) return cudaMemcpyAsync(dst + off, src + off, sizeof(T) * rows * cols, cudaMemcpyHostToDevice, sin);
    return cudaMemcpy2DAsync(dst + off, sizeof(T) * ld, src + off, sizeof(T) * ld, sizeof(T) * rows, cols, cudaMemcpyHostToDevice, sin);
  };
  auto fetch_a = [&](size_t row0, size_t col0, size_t rows, size_t cols) {
    if (hooks && hooks->fetch_a) return hooks->fetch_a(row0, col0, rows, cols);
    return cuda_ok(h2d_window(dA, A, lda, row0, col0, rows, cols), "H2D A");
  };
  auto d2h_block = [&](int j0, int cnt) {  // only the m x cnt window (ldc padding rows on the host are left untouched)
    cudaEvent_t e = ev[next_event++];
    return cuda_ok(cudaEventRecord(e, s), "record GEMM") && cuda_ok(cudaStreamWaitEvent(sout, e, 0), "wait for GEMM") &&
           cuda_ok(cudaMemcpy2DAsync(C + (size_t)j0 * ldc, sizeof(T) * ldc, dC + (size_t)j0 * ldc, sizeof(T) * ldc, sizeof(T) * m, cnt,
                                     cudaMemcpyDeviceToHost, sout), "D2H C");
  };
  if (!cuda_ok(cudaEventRecord(ev[0], s), "fence")) return RTAT_B200_STATUS_EXECUTION_FAILED;
  if (!cuda_ok(cudaStreamWaitEvent(sin, ev[0], 0), "fence")) return RTAT_B200_STATUS_EXECUTION_FAILED;

  int st = RTAT_B200_STATUS_SUCCESS;
  // ---- phase 1: K-panels of A against the first `lead` column blocks (panels == 1: all of A, no blocks) ----
  // The leading blocks are adjacent, so each panel is ONE copy of op(B)'s rows [p0, p0 + kc) under all of them and ONE
  // GEMM lead_cols wide (one GEMM per block measured 27.6 TF over the phase on config 2: 1024-wide, 1024-deep products
  // and a launch gap between each; one product per panel runs at the kernel's large-problem rate).
  const int lead_cols = lead > 0 ? col0[lead - 1] + cols[lead - 1] : 0;
  for (int p = 0; p < npanels; p++) {
    const int p0 = k0s[p], kc = kcs[p];
    // A is stored m x k (N) or k x m (T): the panel is a column block or a row block
    if (!(transa == RTAT_B200_OP_N ? fetch_a(0, p0, m, kc) : fetch_a(p0, 0, kc, m))) return RTAT_B200_STATUS_EXECUTION_FAILED;
    mark(sin, "A-h2d", p, -1);
    if (lead_cols == 0) continue;
    const T *Ap = transa == RTAT_B200_OP_N ? dA + (size_t)p0 * lda : dA + p0;
    if (p0 + kc < k) {
      // op(B) under the leading blocks, K-panel p: B is stored k x n (N) or n x k (T)
      cudaError_t e = transb == RTAT_B200_OP_N ? h2d_window(dB, B, ldb, p0, 0, kc, lead_cols) : h2d_window(dB, B, ldb, 0, p0, lead_cols, kc);
      if (!cuda_ok(e, "H2D B")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      if (p == 0 && *beta != T(0) && !cuda_ok(h2d_window(dC, C, ldc, 0, 0, m, lead_cols), "H2D C")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      mark(sin, "B-h2d", p, 0);
      if (!landed()) return RTAT_B200_STATUS_EXECUTION_FAILED;
      mark(s, "landed", p, 0);
      const T *Bp = transb == RTAT_B200_OP_N ? dB + p0 : dB + (size_t)p0 * ldb;
      st = gemm_fn(h, transa, transb, m, lead_cols, kc, *alpha, Ap, lda, Bp, ldb, p == 0 ? *beta : T(1), dC, ldc);
      if (st) return st;
      mark(s, "gemm", p, 0);
      continue;
    }
    // the last panel finishes the leading blocks: block by block, so that each one's copy back to the host starts as soon
    // as it is done (config 3, where every block is a leading one: one product for the whole panel left all of C's
    // 98 MB to leave after the math, 13.0 -> 10.6 TF end to end)
    for (int j = 0; j < lead; j++) {
      const int j0 = col0[j], cnt = cols[j];
      cudaError_t e = transb == RTAT_B200_OP_N ? h2d_window(dB, B, ldb, p0, j0, kc, cnt) : h2d_window(dB, B, ldb, j0, p0, cnt, kc);
      if (!cuda_ok(e, "H2D B")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      if (p == 0 && *beta != T(0) && !cuda_ok(h2d_window(dC, C, ldc, 0, j0, m, cnt), "H2D C")) return RTAT_B200_STATUS_EXECUTION_FAILED;
      mark(sin, "B-h2d", p, j);
      if (!landed()) return RTAT_B200_STATUS_EXECUTION_FAILED;
      const T *Bjp = transb == RTAT_B200_OP_N ? dB + (size_t)j0 * ldb + p0 : dB + (size_t)p0 * ldb + j0;
      st = gemm_fn(h, transa, transb, m, cnt, kc, *alpha, Ap, lda, Bjp, ldb, p == 0 ? *beta : T(1), dC + (size_t)j0 * ldc, ldc);
      if (st) return st;
      mark(s, "gemm", p, j);
      if (!d2h_block(j0, cnt)) return RTAT_B200_STATUS_EXECUTION_FAILED;
      mark(sout, "C-d2h", p, j);
    }
  }
  // ---- phase 2: whole-K column blocks ----
  for (int j = lead; j < blocks; j++) {
    const int j0 = col0[j], cnt = cols[j];
    cudaError_t e = transb == RTAT_B200_OP_N ? h2d_window(dB, B, ldb, 0, j0, k, cnt) : h2d_window(dB, B, ldb, j0, 0, cnt, k);
    if (!cuda_ok(e, "H2D B")) return RTAT_B200_STATUS_EXECUTION_FAILED;
    if (*beta != T(0) && !cuda_ok(h2d_window(dC, C, ldc, 0, j0, m, cnt), "H2D C")) return RTAT_B200_STATUS_EXECUTION_FAILED;
    mark(sin, "B-h2d", -1, j);
    if (!landed()) return RTAT_B200_STATUS_EXECUTION_FAILED;
    const T *Bj = transb == RTAT_B200_OP_N ? dB + (size_t)j0 * ldb : dB + j0;
    st = gemm_fn(h, transa, transb, m, cnt, k, *alpha, dA, lda, Bj, ldb, *beta, dC + (size_t)j0 * ldc, ldc);
    if (st) return st;
    mark(s, "gemm", -1, j);
    if (!d2h_block(j0, cnt)) return RTAT_B200_STATUS_EXECUTION_FAILED;
    mark(sout, "C-d2h", -1, j);
  }
  if (!cuda_ok(cudaStreamSynchronize(sout), "synchronize") || !cuda_ok(cudaStreamSynchronize(s), "synchronize"))
    return RTAT_B200_STATUS_EXECUTION_FAILED;
  if (trace) {
    cudaDeviceSynchronize();
    fprintf(stderr, "rtat_b200 host pipeline: m %d n %d k %d, %d K-panels, %d column blocks, %d leading\n", m, n, k, npanels, blocks, lead);
    std::string out;
    for (auto &mk : marks) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[0].second, mk.second);
      char line[128];
      snprintf(line, sizeof line, "  dev%d %9.3f ms  %s\n", h->device, ms, mk.first.c_str());
      out += line;
    }
    fputs(out.c_str(), stderr);
    for (auto &mk : marks) cudaEventDestroy(mk.second);
  }
  return RTAT_B200_STATUS_SUCCESS;
}

}  // namespace rtat_b200

// ref_driver — drives the UNMODIFIED reference (headers/sources compiled in place from
// /root/reference/src against cuBLAS) on seeded inputs and dumps its outputs, so that the CPU
// oracle and the sm_100a kernels can be pinned against what the reference itself computes.
// Built by oracle/Makefile into oracle/_ref/ (git-ignored).  TEST INFRASTRUCTURE ONLY.
//
//   ref_driver golden <out.bin> <out.json>      run the fixed golden case list (needs a GPU)
//   ref_driver golden_l3 <out.bin> <out.json>   SYRK / TRSM executors and raw calls (needs a GPU)
//   ref_driver time_move <reps>                 time the reference's MatrixMove (cublas?geam) on bench.py's matrix_op
//                                               shapes with event pairs; one JSON object on stdout (needs a GPU)
//
// Inputs use the same counter-based generator as oracle_fill_uniform_* / rtat_b200_fill_uniform_*
// so the consumers regenerate them from (seed, count) instead of storing them.
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <gemm.h>       // reference: src/methods/gemm.h
#include <matrixop.h>   // reference: src/matrix_ops/matrixop.h
#include <syrk.h>       // reference: src/methods/syrk.h
#include <trsm.h>       // reference: src/methods/trsm.h

using namespace rtat;

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <typename T>
static void fill(std::vector<T> &v, uint64_t seed) {
  for (size_t i = 0; i < v.size(); i++) {
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + i);
    double u = (double)((r >> 11) + 1) * (1.0 / 9007199254740992.0);
    v[i] = (T)(-1.0 + 2.0 * u);
  }
}
// a few special values at fixed positions (signed zero, denormal, huge) for the geam cases
template <typename T>
static void add_specials(std::vector<T> &v) {
  if (v.size() < 8) return;
  v[0] = (T)-0.0;
  v[1] = std::is_same<T, double>::value ? (T)1e-310 : (T)1e-42f;
  v[2] = std::is_same<T, double>::value ? (T)1e300 : (T)1e30f;
  v[3] = (T)0.0;
}

template <typename T>
struct DevMat {
  std::vector<T> host;
  ManagedWorkspace space;
  size_t m, n, ld;
  DevMat(size_t m, size_t n, size_t ld, uint64_t seed, bool specials = false)
      : host(ld * n), space(ld * n * sizeof(T) + 16), m(m), n(n), ld(ld) {
    fill(host, seed);
    if (specials) add_specials(host);
    upload();
  }
  void upload() { gpuAssert(gpu::Memcpy(space, host.data(), host.size() * sizeof(T), gpu::MemcpyHostToDevice)); }
  void download() { gpuAssert(gpu::Memcpy(host.data(), space, host.size() * sizeof(T), gpu::MemcpyDeviceToHost)); }
  Matrix<T> matrix() { return Matrix<T>(space, m, n, ld); }
};

struct Sink {
  FILE *bin, *json;
  size_t offset = 0;
  bool first = true;
  template <typename T>
  void emit(const std::string &meta, const std::vector<T> &data) {
    fprintf(json, "%s\n  {%s, \"dtype\": \"%s\", \"offset\": %zu, \"count\": %zu}", first ? "" : ",", meta.c_str(),
            std::is_same<T, double>::value ? "f64" : "f32", offset, data.size());
    first = false;
    fwrite(data.data(), sizeof(T), data.size(), bin);
    offset += data.size() * sizeof(T);
  }
};

static std::string kv(const char *k, long long v) { return std::string("\"") + k + "\": " + std::to_string(v); }
static std::string kvs(const char *k, const std::string &v) { return std::string("\"") + k + "\": \"" + v + "\""; }
static std::string kvd(const char *k, double v) { char b[64]; snprintf(b, sizeof b, "\"%s\": %.17g", k, v); return b; }

template <typename T, typename Exec, typename Opts>
static void gemm_plan_case(Sink &out, gpu::blasHandle_t handle, Stream s, const char *kind, bool ta, bool tb, int m,
                           int n, int k, T alpha, T beta, uint64_t seed) {
  for (auto &opts : Opts::enumerate()) {
    size_t ar = ta ? k : m, ac = ta ? m : k, br = tb ? n : k, bc = tb ? k : n;
    DevMat<T> A(ar, ac, ar, seed), B(br, bc, br, seed + 1), C(m, n, m, seed + 2);
    GEMM_Inputs<T> in(handle, ta ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, tb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, A.matrix(),
                      B.matrix(), C.matrix(), alpha, beta);
    Exec exec;
    size_t ws = exec.calculate_workspace(in, opts);
    ManagedWorkspace space(ws + 16);
    Workspace exact(( char * )space, ws);
    exec.execute(in, opts, exact, s);
    s.synchronize();
    C.download();
    std::string meta = kvs("kind", kind) + ", " + kvs("plan", std::string(opts)) + ", " + kv("transa", ta) + ", " +
                       kv("transb", tb) + ", " + kv("m", m) + ", " + kv("n", n) + ", " + kv("k", k) + ", " +
                       kvd("alpha", alpha) + ", " + kvd("beta", beta) + ", " + kv("seed", seed) + ", " +
                       kv("workspace_bytes", (long long)ws);
    out.emit(meta, C.host);
  }
}

template <typename T>
static void raw_gemm_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool ta, bool tb, int m, int n, int k, T alpha,
                          T beta, uint64_t seed) {
  size_t ar = ta ? k : m, ac = ta ? m : k, br = tb ? n : k, bc = tb ? k : n;
  DevMat<T> A(ar, ac, ar + 3, seed), B(br, bc, br + 1, seed + 1), C(m, n, m + 2, seed + 2);
  gpuTgemm<T>(handle, ta, tb, A.matrix(), B.matrix(), C.matrix(), alpha, beta);
  s.synchronize();
  C.download();
  std::string meta = kvs("kind", "raw_gemm") + ", " + kv("transa", ta) + ", " + kv("transb", tb) + ", " + kv("m", m) +
                     ", " + kv("n", n) + ", " + kv("k", k) + ", " + kv("lda", ar + 3) + ", " + kv("ldb", br + 1) + ", " +
                     kv("ldc", m + 2) + ", " + kvd("alpha", alpha) + ", " + kvd("beta", beta) + ", " + kv("seed", seed);
  out.emit(meta, C.host);
}

// raw geam through the reference's alias table (gpu::blas{D,S}geam), out of place, all op pairs
template <typename T>
static void raw_geam_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool ta, bool tb, int m, int n, T alpha, T beta,
                          uint64_t seed) {
  size_t ar = ta ? n : m, ac = ta ? m : n, br = tb ? n : m, bc = tb ? m : n;
  DevMat<T> A(ar, ac, ar + 1, seed, true), B(br, bc, br + 2, seed + 1), C(m, n, m + 3, seed + 2);
  if constexpr (std::is_same<T, double>::value)
    gpu::blasDgeam(handle, ta ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, tb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, m, n, &alpha,
                   A.matrix().ptr(), A.ld, &beta, B.matrix().ptr(), B.ld, C.matrix().ptr(), C.ld);
  else
    gpu::blasSgeam(handle, ta ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, tb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, m, n, &alpha,
                   A.matrix().ptr(), A.ld, &beta, B.matrix().ptr(), B.ld, C.matrix().ptr(), C.ld);
  s.synchronize();
  C.download();
  std::string meta = kvs("kind", "raw_geam") + ", " + kv("transa", ta) + ", " + kv("transb", tb) + ", " + kv("m", m) +
                     ", " + kv("n", n) + ", " + kv("lda", ar + 1) + ", " + kv("ldb", br + 2) + ", " + kv("ldc", m + 3) +
                     ", " + kvd("alpha", alpha) + ", " + kvd("beta", beta) + ", " + kv("seed", seed);
  out.emit(meta, C.host);
}

// MatrixMove (matrixop.h:307-344): B = alpha*op(A) into fresh padded buffer
template <typename T>
static void move_case(Sink &out, gpu::blasHandle_t handle, Stream s, int m, int n, T alpha, bool transpose, int pad,
                      uint64_t seed) {
  DevMat<T> A(m, n, m, seed, true);
  std::unique_ptr<MatrixOp<T>> Aop = std::make_unique<NoOp<T>>(A.matrix());
  MatrixMove<T> mv(std::move(Aop), alpha, transpose, pad);
  auto d = mv.dims();
  DevMat<T> B(d.m, d.n, d.ld, seed + 1);  // pre-filled: pad rows must stay untouched
  ManagedWorkspace scratch(mv.scratch_space_req_bytes() + 16);
  mv.execute(handle, Workspace((char *)B.space, d.footprint() * sizeof(T)), scratch);
  s.synchronize();
  B.download();
  std::string meta = kvs("kind", "move") + ", " + kv("m", m) + ", " + kv("n", n) + ", " + kvd("alpha", alpha) + ", " +
                     kv("transpose", transpose) + ", " + kv("pad", pad) + ", " + kv("out_m", d.m) + ", " +
                     kv("out_n", d.n) + ", " + kv("out_ld", d.ld) + ", " + kv("seed", seed);
  out.emit(meta, B.host);
}

// MatrixAccumulate (matrixop.h:256-305): B = alpha*op(A) + beta*B in place
template <typename T>
static void accumulate_case(Sink &out, gpu::blasHandle_t handle, Stream s, int m, int n, int lda_extra, T alpha, T beta,
                            bool transpose, uint64_t seed) {
  size_t ar = transpose ? n : m, ac = transpose ? m : n;
  DevMat<T> A(ar, ac, ar + lda_extra, seed, true), B(m, n, m + 1, seed + 1);
  std::unique_ptr<MatrixOp<T>> Aop = std::make_unique<NoOp<T>>(A.matrix());
  std::unique_ptr<MatrixOp<T>> Bop = std::make_unique<NoOp<T>>(B.matrix());
  MatrixAccumulate<T> acc(std::move(Aop), std::move(Bop), alpha, beta, transpose);
  acc.execute(handle, Workspace(), Workspace());
  s.synchronize();
  B.download();
  std::string meta = kvs("kind", "accumulate") + ", " + kv("m", m) + ", " + kv("n", n) + ", " + kv("lda", ar + lda_extra) +
                     ", " + kv("ldb", m + 1) + ", " + kvd("alpha", alpha) + ", " + kvd("beta", beta) + ", " +
                     kv("transpose", transpose) + ", " + kv("seed", seed);
  out.emit(meta, B.host);
}

template <typename T>
static void all_cases(Sink &out, gpu::blasHandle_t handle, Stream s, uint64_t base) {
  uint64_t seed = base;
  // every plan and op pair (as planning_test.cpp:86-114 / executor_test.cpp:12-88 do) and op pair, alpha=1 beta=0 (test_harness.h:244-247)
  for (bool ta : {false, true})
    for (bool tb : {false, true}) {
      gemm_plan_case<T, GEMM_Executor<T>, GEMM_Options>(out, handle, s, "gemm_plan", ta, tb, 13, 10, 17, (T)1, (T)0, seed);
      seed += 3;
    }
  gemm_plan_case<T, GEMM_Executor<T>, GEMM_Options>(out, handle, s, "gemm_plan", false, false, 13, 10, 17, (T)-0.75, (T)0.5, seed);
  seed += 3;
  // 64 pad plans on odd sizes (C3-like op pair NT for double, TN for float)
  bool is_d = std::is_same<T, double>::value;
  gemm_plan_case<T, GEMM_Executor_Pad<T>, GEMM_Options_Pad>(out, handle, s, "gemm_pad_plan", !is_d, is_d, 13, 9, 11, (T)1, (T)0, seed);
  seed += 3;
  gemm_plan_case<T, GEMM_Executor_Pad<T>, GEMM_Options_Pad>(out, handle, s, "gemm_pad_plan", false, false, 12, 10, 7, (T)1.5, (T)-0.25, seed);
  seed += 3;
  // executor_test.cpp:12-88 shape, raw gemm with non-tight ld
  for (bool ta : {false, true})
    for (bool tb : {false, true}) {
      if (ta == tb) raw_gemm_case<T>(out, handle, s, ta, tb, 70, 45, 62, (T)1, (T)0, seed);
      else raw_gemm_case<T>(out, handle, s, ta, tb, 20, 12, 17, (T)1, (T)0, seed);
      seed += 3;
      raw_gemm_case<T>(out, handle, s, ta, tb, 20, 12, 17, (T)2, (T)-1, seed);
      seed += 3;
    }
  for (bool ta : {false, true})
    for (bool tb : {false, true})
      for (int c = 0; c < 4; c++) {
        T alphas[4] = {(T)1, (T)-2, (T)0.3, (T)0};
        T betas[4] = {(T)0, (T)1, (T)-1.7, (T)0.6};
        raw_geam_case<T>(out, handle, s, ta, tb, 13, 7, alphas[c], betas[c], seed);
        seed += 3;
      }
  move_case<T>(out, handle, s, 10, 12, (T)1, false, 32, seed); seed += 3;   // matrixop_test.cpp:10-31
  move_case<T>(out, handle, s, 14, 25, (T)-2, true, 8, seed); seed += 3;    // matrixop_test.cpp:252
  move_case<T>(out, handle, s, 42, 23, (T)0.5, true, 16, seed); seed += 3;  // matrixop_test.cpp:289
  move_case<T>(out, handle, s, 67, 70, (T)1, true, 1, seed); seed += 3;
  move_case<T>(out, handle, s, 129, 65, (T)1, false, 1, seed); seed += 3;
  for (bool tr : {false, true})
    for (int c = 0; c < 4; c++) {
      T betas[4] = {(T)0, (T)1, (T)0.5, (T)-1.25};
      accumulate_case<T>(out, handle, s, 29, 18, c, (T)1, betas[c], tr, seed);
      seed += 3;
    }
}

// SYRK_Executor over all four plans (executor_test.cpp:227-290 walks the same loops); the WHOLE of C is dumped so
// that what each plan does to the triangle BLAS leaves alone is pinned too.
template <typename T>
static void syrk_plan_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool lower, bool trans, int n, int k, T alpha,
                           T beta, uint64_t seed) {
  for (auto &opts : SYRK_Options::enumerate()) {
    size_t ar = trans ? k : n, ac = trans ? n : k;
    DevMat<T> A(ar, ac, ar, seed), C(n, n, n, seed + 1);
    SYRK_Inputs<T> in(handle, lower ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER,
                      trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, A.matrix(), C.matrix(), alpha, beta);
    SYRK_Executor<T> exec;
    size_t ws = exec.calculate_workspace(in, opts);
    ManagedWorkspace space(ws + 16);
    gpuAssert(gpu::Memset(space, 0xFF, ws + 16));  // NaN-patterned scratch: shows whether geam(0, 0) really clears it
    Workspace exact((char *)space, ws);
    exec.execute(in, opts, exact, s);
    s.synchronize();
    C.download();
    std::string meta = kvs("kind", "syrk_plan") + ", " + kvs("plan", std::string(opts)) + ", " + kv("lower", lower) + ", " +
                       kv("trans", trans) + ", " + kv("n", n) + ", " + kv("k", k) + ", " + kvd("alpha", alpha) + ", " +
                       kvd("beta", beta) + ", " + kv("seed", seed) + ", " + kv("workspace_bytes", (long long)ws) + ", " +
                       kvs("key", std::string(SYRK_Key(in)));
    out.emit(meta, C.host);
  }
}

template <typename T>
static void raw_syrk_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool lower, bool trans, int n, int k, T alpha,
                          T beta, uint64_t seed) {
  size_t ar = trans ? k : n, ac = trans ? n : k;
  DevMat<T> A(ar, ac, ar + 3, seed), C(n, n, n + 2, seed + 1);
  gpuTsyrk<T>(handle, lower, trans, A.matrix(), C.matrix(), alpha, beta);
  s.synchronize();
  C.download();
  std::string meta = kvs("kind", "raw_syrk") + ", " + kv("lower", lower) + ", " + kv("trans", trans) + ", " + kv("n", n) +
                     ", " + kv("k", k) + ", " + kv("lda", ar + 3) + ", " + kv("ldc", n + 2) + ", " + kvd("alpha", alpha) +
                     ", " + kvd("beta", beta) + ", " + kv("seed", seed);
  out.emit(meta, C.host);
}

// Triangular test matrix: uniform fill, then the diagonal is pushed away from zero (non-unit) or the
// off-diagonals are shrunk (unit), as executor_test.cpp:146-176 does; the unused triangle keeps its random
// contents so that reading it would show.
template <typename T>
static void condition_triangle(DevMat<T> &A, bool unit) {
  size_t n = A.m;
  for (size_t j = 0; j < n; j++)
    for (size_t i = 0; i < n; i++) {
      T &a = A.host[j * A.ld + i];
      if (i == j && !unit) a += (a < 0 ? -(T)2 : (T)2);
      if (i != j) a /= (T)4;
    }
  A.upload();
}

template <typename T>
static void trsm_plan_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool left, bool lower, bool trans, bool unit,
                           int m, int n, T alpha, uint64_t seed) {
  for (auto &opts : TRSM_Options::enumerate()) {
    size_t na = left ? m : n;
    DevMat<T> A(na, na, na, seed), B(m, n, m, seed + 1);
    condition_triangle(A, unit);
    TRSM_Inputs<T> in(handle, left ? gpu::BLAS_SIDE_LEFT : gpu::BLAS_SIDE_RIGHT,
                      lower ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER, trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N,
                      unit ? gpu::BLAS_DIAG_UNIT : gpu::BLAS_DIAG_NON_UNIT, A.matrix(), B.matrix(), alpha);
    TRSM_Executor<T> exec;
    size_t ws = exec.calculate_workspace(in, opts);
    ManagedWorkspace space(ws + 16);
    gpuAssert(gpu::Memset(space, 0xFF, ws + 16));
    Workspace exact((char *)space, ws);
    exec.execute(in, opts, exact, s);
    s.synchronize();
    B.download();
    std::string meta = kvs("kind", "trsm_plan") + ", " + kvs("plan", std::string(opts)) + ", " + kv("left", left) + ", " +
                       kv("lower", lower) + ", " + kv("trans", trans) + ", " + kv("unit", unit) + ", " + kv("m", m) + ", " +
                       kv("n", n) + ", " + kvd("alpha", alpha) + ", " + kv("seed", seed) + ", " +
                       kv("workspace_bytes", (long long)ws) + ", " + kvs("key", std::string(TRSM_Key(in)));
    out.emit(meta, B.host);
  }
}

template <typename T>
static void raw_trsm_case(Sink &out, gpu::blasHandle_t handle, Stream s, bool left, bool lower, bool trans, bool unit, int m,
                          int n, T alpha, uint64_t seed) {
  size_t na = left ? m : n;
  DevMat<T> A(na, na, na + 1, seed), B(m, n, m + 3, seed + 1);
  condition_triangle(A, unit);
  gpuTtrsm<T>(handle, left, lower, trans, unit, A.matrix(), B.matrix(), alpha);
  s.synchronize();
  B.download();
  std::string meta = kvs("kind", "raw_trsm") + ", " + kv("left", left) + ", " + kv("lower", lower) + ", " + kv("trans", trans) +
                     ", " + kv("unit", unit) + ", " + kv("m", m) + ", " + kv("n", n) + ", " + kv("lda", na + 1) + ", " +
                     kv("ldb", m + 3) + ", " + kvd("alpha", alpha) + ", " + kv("seed", seed);
  out.emit(meta, B.host);
}

template <typename T>
static void l3_cases(Sink &out, gpu::blasHandle_t handle, Stream s, uint64_t base) {
  uint64_t seed = base;
  for (bool lower : {false, true})
    for (bool trans : {false, true}) {
      syrk_plan_case<T>(out, handle, s, lower, trans, 37, 19, (T)1, (T)0, seed); seed += 2;     // harness defaults (test_harness.h:274-277)
      syrk_plan_case<T>(out, handle, s, lower, trans, 21, 30, (T)-1, (T)1, seed); seed += 2;    // executor_test.cpp:262-266
      raw_syrk_case<T>(out, handle, s, lower, trans, 45, 23, (T)0.75, (T)-0.5, seed); seed += 2;
      raw_syrk_case<T>(out, handle, s, lower, trans, 18, 9, (T)0, (T)2, seed); seed += 2;       // alpha == 0: scale only
    }
  for (bool left : {false, true})
    for (bool lower : {false, true})
      for (bool trans : {false, true})
        for (bool unit : {false, true}) {
          trsm_plan_case<T>(out, handle, s, left, lower, trans, unit, 21, 17, (T)1, seed); seed += 2;
          raw_trsm_case<T>(out, handle, s, left, lower, trans, unit, 40, 33, (T)-1.5, seed); seed += 2;
        }
}

// MatrixMove (reference src/matrix_ops/matrixop.h:307-344) of an m x n source, timed like Executor::execute times a
// plan (event pair on the handle's stream, reference src/methods/executor.h:18-33): best and median of `reps` runs.
template <typename T>
static void time_move_case(gpu::blasHandle_t handle, Stream s, const char *name, size_t m, size_t n, bool transpose, size_t pad,
                           int reps, bool last) {
  ManagedWorkspace src(m * n * sizeof(T) + 16);
  gpuAssert(gpu::Memset(src, 0x3c, m * n * sizeof(T)));
  Matrix<T> A(src, m, n, m);
  std::unique_ptr<MatrixOp<T>> leaf = std::make_unique<NoOp<T>>(A);
  MatrixMove<T> move(std::move(leaf), (T)1, transpose, pad);
  const size_t out_bytes = move.output_space_req() * sizeof(T);
  ManagedWorkspace dst(out_bytes + 16);
  Workspace out((char *)dst, out_bytes), none((char *)dst + out_bytes, 0);
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; r++) {
    Event a, b;
    a.record(s);
    move.execute(handle, out, none);
    b.record(s);
    s.synchronize();
    if (r >= 3) ms.push_back(Event::elapsed_time(a, b));
  }
  std::sort(ms.begin(), ms.end());
  const double bytes = 2.0 * m * n * sizeof(T);
  printf("  \"%s\": {\"rows\": %zu, \"cols\": %zu, \"transpose\": %d, \"ld_out\": %zu, \"bytes\": %.0f, \"best_ms\": %.6f, \"median_ms\": %.6f, "
         "\"best_gbs\": %.1f, \"median_gbs\": %.1f}%s\n",
         name, m, n, (int)transpose, move.dims().ld, bytes, ms.front(), ms[ms.size() / 2], bytes / ms.front() * 1e-6, bytes / ms[ms.size() / 2] * 1e-6,
         last ? "" : ",");
}

static int time_move_main(int reps) {
  Stream s;
  gpu::blasHandle_t handle;
  gpu::blasCreate(&handle);
  gpu::blasSetStream(handle, s);
  printf("{\n");
  time_move_case<double>(handle, s, "transpose_8192_f64", 8192, 8192, true, 1, reps, false);
  time_move_case<double>(handle, s, "copy_8192_f64", 8192, 8192, false, 1, reps, false);
  time_move_case<double>(handle, s, "pad_copy_c3_A", 4097, 2049, false, 32, reps, false);   // config 3: A 4097 x 2049 -> ld 4128
  time_move_case<double>(handle, s, "pad_copy_c3_B", 3001, 2049, false, 32, reps, false);   // config 3: B 3001 x 2049 -> ld 3008
  time_move_case<double>(handle, s, "transpose_pad_c3_A", 4097, 2049, true, 32, reps, false);  // -> 2049 x 4097, ld 2080
  time_move_case<double>(handle, s, "transpose_c3_A_tight", 4097, 2049, true, 1, reps, false); // -> ld 2049 (odd target)
  time_move_case<float>(handle, s, "transpose_131072x128_f32", 131072, 128, true, 1, reps, true);
  printf("}\n");
  gpu::blasDestroy(handle);
  return 0;
}

int main(int argc, char **argv) {
  if (argc == 3 && std::string(argv[1]) == "time_move") return time_move_main(atoi(argv[2]) > 0 ? atoi(argv[2]) : 20);
  if (argc != 4 || (std::string(argv[1]) != "golden" && std::string(argv[1]) != "golden_l3")) {
    fprintf(stderr, "usage: ref_driver golden|golden_l3 <out.bin> <out.json>\n");
    return 2;
  }
  const bool l3 = std::string(argv[1]) == "golden_l3";
  Sink out;
  out.bin = fopen(argv[2], "wb");
  out.json = fopen(argv[3], "w");
  if (!out.bin || !out.json) return 3;
  Stream s;
  gpu::blasHandle_t handle;
  gpu::blasCreate(&handle);
  gpu::blasSetStream(handle, s);
  fprintf(out.json, "{\"generator\": \"oracle/ref_driver.cpp golden (reference executor + cuBLAS)\", \"cases\": [");
  if (l3) {
    l3_cases<double>(out, handle, s, 20000);
    l3_cases<float>(out, handle, s, 30000);
  } else {
    all_cases<double>(out, handle, s, 1000);
    all_cases<float>(out, handle, s, 5000);
  }
  fprintf(out.json, "\n]}\n");
  fclose(out.bin);
  fclose(out.json);
  gpu::blasDestroy(handle);
  printf("wrote %zu bytes\n", out.offset);
  return 0;
}

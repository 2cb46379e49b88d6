// librtat_host.so — C entry points over the C++ host layer (planner + executor + plan algebra), so
// that non-C++ callers (the Python tests and bench.py via ctypes) drive exactly the code path a C++
// user of rtat.h gets.  Declared in include/rtat_b200_host.h.
#include <cstring>
#include <string>
#include <rtat/rtat.h>
#include "../../include/rtat_b200_host.h"

using namespace rtat;

namespace {

struct Planner_Base {
  virtual ~Planner_Base() = default;
  virtual int gemm(rtat_b200_handle_t h, int ta, int tb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B,
                   int ldb, const void *beta, void *C, int ldc, void *ws, size_t ws_bytes, const char *forced, char *plan_out) = 0;
  virtual long long workspace(int ta, int tb, int m, int n, int k, int lda, int ldb, int ldc, const char *plan) = 0;
  virtual int syrk(rtat_b200_handle_t, int, int, int, int, const void *, const void *, int, const void *, void *, int, void *, size_t,
                   const char *, char *) {
    throw std::runtime_error("this planner is not a syrk planner");
  }
  virtual long long syrk_workspace(int, int, int, int, int, int, const char *) { throw std::runtime_error("this planner is not a syrk planner"); }
  virtual int trsm(rtat_b200_handle_t, int, int, int, int, int, int, const void *, const void *, int, void *, int, void *, size_t, const char *,
                   char *) {
    throw std::runtime_error("this planner is not a trsm planner");
  }
  virtual long long trsm_workspace(int, int, int, int, int, int, int, int, const char *) {
    throw std::runtime_error("this planner is not a trsm planner");
  }
  virtual std::string statistics() = 0;
  virtual std::string save_plans() = 0;
  virtual long long load_plans(const char *text) = 0;
  virtual void set_converge(int n) = 0;
  virtual void set_preference(double margin) = 0;
  virtual void save_plans_file(const char *path) = 0;
  virtual long long load_plans_file(const char *path) = 0;
  virtual std::string statistics_rates() = 0;
  virtual int num_plans() = 0;
  virtual std::string plan_name(int i) = 0;
};

// what every method's planner shares: plan parsing, statistics, the plan cache
template <typename Exec, typename Opts>
struct Planner_Common : Planner_Base {
  Planning_System<Exec> planner;

  static Opts parse(const char *plan) {
    Opts o;
    std::istringstream in(plan);
    if (!(in >> o)) throw std::runtime_error(std::string("bad plan string ") + plan);
    return o;
  }
  template <typename Inputs, typename Key>
  int run(rtat_b200_handle_t h, Inputs in, Key key, void *ws, size_t ws_bytes, const char *forced, char *plan_out) {
    Opts plan = forced && *forced ? parse(forced) : planner.create_plan(key);
    gpu::Stream_t raw = nullptr;
    gpu::blasGetStream(h, &raw);
    planner.execute(in, plan, Workspace((char *)ws, ws_bytes), Stream(raw));
    if (plan_out) std::strcpy(plan_out, std::string(plan).c_str());
    return 0;
  }
  int gemm(rtat_b200_handle_t, int, int, int, int, int, const void *, const void *, int, const void *, int, const void *, void *, int, void *,
           size_t, const char *, char *) override {
    throw std::runtime_error("this planner is not a gemm planner");
  }
  long long workspace(int, int, int, int, int, int, int, int, const char *) override { throw std::runtime_error("this planner is not a gemm planner"); }
  std::string statistics() override { return planner.make_statistics().json().dump(1); }
  std::string save_plans() override { return planner.save_plans().dump(1); }
  long long load_plans(const char *text) override { return (long long)planner.load_plans(Json::parse(text)); }
  void set_converge(int n) override { planner.set_tests_until_converge((size_t)n); }
  void set_preference(double margin) override { planner.set_default_preference((float)margin); }
  void save_plans_file(const char *path) override { planner.save_plans_file(path); }
  long long load_plans_file(const char *path) override { return (long long)planner.load_plans_file(path); }
  std::string statistics_rates() override { return planner.make_statistics().json(sizeof(typename Exec::Params_T::Scalar)).dump(1); }
  int num_plans() override { return (int)Opts::enumerate().size(); }
  std::string plan_name(int i) override { return std::string(Opts::enumerate().at((size_t)i)); }
};

template <typename T>
struct Syrk_Planner : Planner_Common<SYRK_Executor<T>, SYRK_Options> {
  static SYRK_Inputs<T> inputs(rtat_b200_handle_t h, int uplo, int trans, int n, int k, T alpha, const void *A, int lda, T beta, void *C, int ldc) {
    const size_t ar = trans ? k : n, ac = trans ? n : k;
    Matrix<T> Am(Workspace((T *)A, (size_t)lda * ac), ar, ac, lda), Cm(Workspace((T *)C, (size_t)ldc * n), n, n, ldc);
    return SYRK_Inputs<T>(h, uplo == RTAT_B200_FILL_LOWER ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER,
                          trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, Am, Cm, alpha, beta);
  }
  int syrk(rtat_b200_handle_t h, int uplo, int trans, int n, int k, const void *alpha, const void *A, int lda, const void *beta, void *C, int ldc,
           void *ws, size_t ws_bytes, const char *forced, char *plan_out) override {
    auto in = inputs(h, uplo, trans, n, k, *(const T *)alpha, A, lda, *(const T *)beta, C, ldc);
    return this->run(h, in, SYRK_Key(in), ws, ws_bytes, forced, plan_out);
  }
  long long syrk_workspace(int uplo, int trans, int n, int k, int lda, int ldc, const char *plan) override {
    auto in = inputs(nullptr, uplo, trans, n, k, T(1), nullptr, lda, T(0), nullptr, ldc);
    return (long long)this->planner.calculate_workspace(in, this->parse(plan));
  }
};

template <typename T>
struct Trsm_Planner : Planner_Common<TRSM_Executor<T>, TRSM_Options> {
  static TRSM_Inputs<T> inputs(rtat_b200_handle_t h, int side, int uplo, int trans, int diag, int m, int n, T alpha, const void *A, int lda,
                               void *B, int ldb) {
    const size_t na = side == RTAT_B200_SIDE_LEFT ? m : n;
    Matrix<T> Am(Workspace((T *)A, (size_t)lda * na), na, na, lda), Bm(Workspace((T *)B, (size_t)ldb * n), m, n, ldb);
    return TRSM_Inputs<T>(h, side == RTAT_B200_SIDE_LEFT ? gpu::BLAS_SIDE_LEFT : gpu::BLAS_SIDE_RIGHT,
                          uplo == RTAT_B200_FILL_LOWER ? gpu::BLAS_FILL_MODE_LOWER : gpu::BLAS_FILL_MODE_UPPER,
                          trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, diag == RTAT_B200_DIAG_UNIT ? gpu::BLAS_DIAG_UNIT : gpu::BLAS_DIAG_NON_UNIT, Am,
                          Bm, alpha);
  }
  int trsm(rtat_b200_handle_t h, int side, int uplo, int trans, int diag, int m, int n, const void *alpha, const void *A, int lda, void *B,
           int ldb, void *ws, size_t ws_bytes, const char *forced, char *plan_out) override {
    auto in = inputs(h, side, uplo, trans, diag, m, n, *(const T *)alpha, A, lda, B, ldb);
    return this->run(h, in, TRSM_Key(in), ws, ws_bytes, forced, plan_out);
  }
  long long trsm_workspace(int side, int uplo, int trans, int diag, int m, int n, int lda, int ldb, const char *plan) override {
    auto in = inputs(nullptr, side, uplo, trans, diag, m, n, T(1), nullptr, lda, nullptr, ldb);
    return (long long)this->planner.calculate_workspace(in, this->parse(plan));
  }
};

template <typename T, typename Exec, typename Opts>
struct Planner_Impl : Planner_Common<Exec, Opts> {
  using Planner_Common<Exec, Opts>::planner;
  using Planner_Common<Exec, Opts>::parse;
  static GEMM_Inputs<T> inputs(rtat_b200_handle_t h, int ta, int tb, int m, int n, int k, T alpha, const void *A, int lda, const void *B,
                               int ldb, T beta, void *C, int ldc) {
    const size_t ac = ta ? m : k, bc = tb ? k : n;
    Matrix<T> Am(Workspace((T *)A, (size_t)lda * ac), ta ? k : m, ac, lda);
    Matrix<T> Bm(Workspace((T *)B, (size_t)ldb * bc), tb ? n : k, bc, ldb);
    Matrix<T> Cm(Workspace((T *)C, (size_t)ldc * n), m, n, ldc);
    return GEMM_Inputs<T>(h, ta ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, tb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, Am, Bm, Cm, alpha, beta);
  }
  int gemm(rtat_b200_handle_t h, int ta, int tb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb,
           const void *beta, void *C, int ldc, void *ws, size_t ws_bytes, const char *forced, char *plan_out) override {
    auto in = inputs(h, ta, tb, m, n, k, *(const T *)alpha, A, lda, B, ldb, *(const T *)beta, C, ldc);
    return this->run(h, in, GEMM_Key(in), ws, ws_bytes, forced, plan_out);
  }
  long long workspace(int ta, int tb, int m, int n, int k, int lda, int ldb, int ldc, const char *plan) override {
    auto in = inputs(nullptr, ta, tb, m, n, k, T(1), nullptr, lda, nullptr, ldb, T(0), nullptr, ldc);
    return (long long)planner.calculate_workspace(in, parse(plan));
  }
};

thread_local std::string last_error;

template <typename F>
int guarded(F f) {
  try {
    return f();
  } catch (const std::exception &e) {
    last_error = e.what();
  } catch (const char *msg) {
    last_error = msg;
  } catch (...) {
    last_error = "unknown error";
  }
  return -1;
}

char *dup_string(const std::string &s) {
  char *out = (char *)std::malloc(s.size() + 1);
  std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

}  // namespace

extern "C" {

rtat_host_planner_t rtat_host_planner_create(const char *method, const char *data_type) {
  const std::string m = method ? method : "", d = data_type ? data_type : "";
  Planner_Base *p = nullptr;
  guarded([&] {
    if (m == "gemm" && d == "double") p = new Planner_Impl<double, GEMM_Executor<double>, GEMM_Options>();
    else if (m == "gemm" && d == "float") p = new Planner_Impl<float, GEMM_Executor<float>, GEMM_Options>();
    else if (m == "gemm_pad" && d == "double") p = new Planner_Impl<double, GEMM_Executor_Pad<double>, GEMM_Options_Pad>();
    else if (m == "gemm_pad" && d == "float") p = new Planner_Impl<float, GEMM_Executor_Pad<float>, GEMM_Options_Pad>();
    else if (m == "syrk" && d == "double") p = new Syrk_Planner<double>();
    else if (m == "syrk" && d == "float") p = new Syrk_Planner<float>();
    else if (m == "trsm" && d == "double") p = new Trsm_Planner<double>();
    else if (m == "trsm" && d == "float") p = new Trsm_Planner<float>();
    else throw std::runtime_error("unknown method/data_type: " + m + "/" + d);
    return 0;
  });
  return p;
}
void rtat_host_planner_destroy(rtat_host_planner_t p) { delete static_cast<Planner_Base *>(p); }

int rtat_host_planner_gemm(rtat_host_planner_t p, rtat_b200_handle_t h, int transa, int transb, int m, int n, int k, const void *alpha,
                           const void *A, int lda, const void *B, int ldb, const void *beta, void *C, int ldc, void *workspace,
                           size_t workspace_bytes, const char *forced_plan, char *plan_out) {
  if (!p || !h) return -1;
  return guarded([&] {
    return static_cast<Planner_Base *>(p)->gemm(h, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, workspace, workspace_bytes,
                                                forced_plan, plan_out);
  });
}
long long rtat_host_planner_workspace(rtat_host_planner_t p, int transa, int transb, int m, int n, int k, int lda, int ldb, int ldc,
                                      const char *plan) {
  if (!p) return -1;
  long long out = -1;
  guarded([&] {
    out = static_cast<Planner_Base *>(p)->workspace(transa, transb, m, n, k, lda, ldb, ldc, plan);
    return 0;
  });
  return out;
}
int rtat_host_planner_syrk(rtat_host_planner_t p, rtat_b200_handle_t h, int uplo, int trans, int n, int k, const void *alpha, const void *A,
                           int lda, const void *beta, void *C, int ldc, void *workspace, size_t workspace_bytes, const char *forced_plan,
                           char *plan_out) {
  if (!p || !h) return -1;
  return guarded([&] {
    return static_cast<Planner_Base *>(p)->syrk(h, uplo, trans, n, k, alpha, A, lda, beta, C, ldc, workspace, workspace_bytes, forced_plan, plan_out);
  });
}
long long rtat_host_planner_syrk_workspace(rtat_host_planner_t p, int uplo, int trans, int n, int k, int lda, int ldc, const char *plan) {
  long long out = -1;
  if (p) guarded([&] { out = static_cast<Planner_Base *>(p)->syrk_workspace(uplo, trans, n, k, lda, ldc, plan); return 0; });
  return out;
}
int rtat_host_planner_trsm(rtat_host_planner_t p, rtat_b200_handle_t h, int side, int uplo, int trans, int diag, int m, int n,
                           const void *alpha, const void *A, int lda, void *B, int ldb, void *workspace, size_t workspace_bytes,
                           const char *forced_plan, char *plan_out) {
  if (!p || !h) return -1;
  return guarded([&] {
    return static_cast<Planner_Base *>(p)->trsm(h, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, workspace, workspace_bytes, forced_plan,
                                                plan_out);
  });
}
long long rtat_host_planner_trsm_workspace(rtat_host_planner_t p, int side, int uplo, int trans, int diag, int m, int n, int lda, int ldb,
                                           const char *plan) {
  long long out = -1;
  if (p) guarded([&] { out = static_cast<Planner_Base *>(p)->trsm_workspace(side, uplo, trans, diag, m, n, lda, ldb, plan); return 0; });
  return out;
}
char *rtat_host_planner_statistics(rtat_host_planner_t p) {
  char *out = nullptr;
  if (p) guarded([&] { out = dup_string(static_cast<Planner_Base *>(p)->statistics()); return 0; });
  return out;
}
char *rtat_host_planner_save_plans(rtat_host_planner_t p) {
  char *out = nullptr;
  if (p) guarded([&] { out = dup_string(static_cast<Planner_Base *>(p)->save_plans()); return 0; });
  return out;
}
long long rtat_host_planner_load_plans(rtat_host_planner_t p, const char *json_text) {
  long long out = -1;
  if (p && json_text) guarded([&] { out = static_cast<Planner_Base *>(p)->load_plans(json_text); return 0; });
  return out;
}
int rtat_host_planner_set_tests_until_converge(rtat_host_planner_t p, int n) {
  if (!p) return -1;
  static_cast<Planner_Base *>(p)->set_converge(n);
  return 0;
}
int rtat_host_planner_save_plans_file(rtat_host_planner_t p, const char *path) {
  if (!p || !path) return -1;
  return guarded([&] { static_cast<Planner_Base *>(p)->save_plans_file(path); return 0; });
}
long long rtat_host_planner_load_plans_file(rtat_host_planner_t p, const char *path) {
  long long out = -1;
  if (p && path) guarded([&] { out = static_cast<Planner_Base *>(p)->load_plans_file(path); return 0; });
  return out;
}
int rtat_host_planner_set_default_preference(rtat_host_planner_t p, double margin) {
  if (!p) return -1;
  static_cast<Planner_Base *>(p)->set_preference(margin);
  return 0;
}
char *rtat_host_planner_statistics_rates(rtat_host_planner_t p) {
  char *out = nullptr;
  if (p) guarded([&] { out = dup_string(static_cast<Planner_Base *>(p)->statistics_rates()); return 0; });
  return out;
}
int rtat_host_planner_num_plans(rtat_host_planner_t p) { return p ? static_cast<Planner_Base *>(p)->num_plans() : -1; }
char *rtat_host_planner_plan_name(rtat_host_planner_t p, int i) {
  char *out = nullptr;
  if (p) guarded([&] { out = dup_string(static_cast<Planner_Base *>(p)->plan_name(i)); return 0; });
  return out;
}
void rtat_host_free(char *s) { std::free(s); }
const char *rtat_host_last_error(void) { return last_error.c_str(); }

}  // extern "C"

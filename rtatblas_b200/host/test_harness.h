// test_harness.h — JSON-driven benchmark harness (reference src/app/test_harness.h).
// Input  {"keywords": {"run_type": "exhaustive"|"autotune", "method": "gemm"|"gemm_pad"|"syrk"|"trsm"|"move",
//                      "data_type": "double"|"float", "repetitions": N [, "seed": S] [, "plan_cache": "<path>"]
//                      [, "rates": false]},
//         "problems": [ {"transA","transB","m","n","k"}, ... ]}      (per method: the key's JSON form)
// Output the same document with "problems" replaced by the planner statistics; unless "rates" is false every plan entry
// also carries "tflops" and (GEMM keys) "gbs" next to "times".
// "plan_cache": the planner's converged plans are loaded from that file before the run (a missing file is an empty
// cache) and written back after it, so a later process starts tuned (reference src/app/runner.h:81-82 keeps its
// statistics on disk the same way).
// Differences from the reference: inputs come from a seeded generator and are generated once per
// problem instead of before every repetition; method "move" (MatrixMove alone, GB/s) is new.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include <rtat/gemm.h>
#include <rtat/json_encoding.h>
#include <rtat/planning_system.h>

namespace rtat {

struct Run_Type {
  enum _Run_Type { EXHAUSTIVE, AUTOTUNE } val;
  Run_Type(const std::string &s) {
    if (s == "exhaustive") val = EXHAUSTIVE;
    else if (s == "autotune") val = AUTOTUNE;
    else throw std::runtime_error("Invalid run_type: " + s);
  }
  operator std::string() const { return val == EXHAUSTIVE ? "exhaustive" : "autotune"; }
};

struct Method {
  enum _Method { GEMM, GEMM_PAD, SYRK, TRSM, MOVE } val;
  Method(const std::string &s) {
    if (s == "gemm") val = GEMM;
    else if (s == "gemm_pad") val = GEMM_PAD;
    else if (s == "syrk") val = SYRK;
    else if (s == "trsm") val = TRSM;
    else if (s == "move") val = MOVE;
    else throw std::runtime_error("Invalid method: " + s);
  }
  operator std::string() const {
    return val == GEMM ? "gemm" : val == GEMM_PAD ? "gemm_pad" : val == SYRK ? "syrk" : val == TRSM ? "trsm" : "move";
  }
};

struct Data_Type {
  enum _Data_Type { FLOAT, DOUBLE } val;
  Data_Type(const std::string &s) {
    if (s == "double") val = DOUBLE;
    else if (s == "float") val = FLOAT;
    else throw std::runtime_error("Invalid data_type: " + s);
  }
  operator std::string() const { return val == DOUBLE ? "double" : "float"; }
};

template <typename Key>
class Problem_Set {
  std::vector<Key> problems;

 public:
  explicit Problem_Set(const Json &list) {
    for (auto &p : list.items()) problems.push_back(from_json<Key>(p));
  }
  const std::vector<Key> &get_problems() const { return problems; }
};

struct Input_File {
  const Json problem_json;
  const Run_Type run_type;
  const Method method;
  const Data_Type data_type;
  const int repetitions;
  const long long seed;
  const std::string plan_cache;  // empty: none
  const bool rates;
  explicit Input_File(const Json &in)
      : problem_json(in["problems"]),
        run_type(in["keywords"]["run_type"].get<std::string>()),
        method(in["keywords"]["method"].get<std::string>()),
        data_type(in["keywords"]["data_type"].get<std::string>()),
        repetitions(in["keywords"]["repetitions"].get<int>()),
        seed(in["keywords"].contains("seed") ? in["keywords"]["seed"].get<long long>() : 0xB200),
        plan_cache(in["keywords"].contains("plan_cache") ? in["keywords"]["plan_cache"].get<std::string>() : std::string()),
        rates(in["keywords"].contains("rates") ? in["keywords"]["rates"].get<bool>() : true) {}
};

// stream + handle + a grow-only input arena + a grow-only scratch arena
struct Device_Resources {
  Stream s;
  gpu::blasHandle_t handle = nullptr;
  ManagedWorkspace scratch_space;

 private:
  ManagedWorkspace space;
  Device_RNG rng;

 public:
  explicit Device_Resources(uint64_t seed = 0xB200) : s(), scratch_space(1024), space(1024), rng(s, seed) {
    blas_status_check(gpu::blasCreate(&handle), "blasCreate");
    gpu::blasSetStream(handle, s);
  }
  ~Device_Resources() { gpu::blasDestroy(handle); }
  void sync() { s.synchronize(); }

  // carve 512-byte-aligned slices for the given shapes and fill them with U(0,1]
  template <typename T>
  std::vector<Matrix<T>> allocate_matrices(const std::vector<MatrixDims> &shapes) {
    std::vector<size_t> elems;
    size_t total = 0;
    for (auto &d : shapes) {
      size_t e = ((d.footprint() * sizeof(T) + 511) / 512) * 512 / sizeof(T);
      elems.push_back(e);
      total += e;
    }
    space.grow_to_fit<T>(total);
    std::vector<Matrix<T>> out;
    size_t offset = 0;
    for (size_t i = 0; i < shapes.size(); i++) {
      Workspace slice(space, offset, elems[i] * sizeof(T));
      rng.uniform<T>((T *)slice, slice.size<T>());
      out.emplace_back(slice, shapes[i]);
      offset += elems[i] * sizeof(T);
    }
    return out;
  }
};

template <typename Planner_Type>
class Test_Harness {
  using Params = typename Planner_Type::Params;
  using Key = typename Planner_Type::Key;
  using Opts = typename Planner_Type::Opts;
  using Scalar = typename Params::Scalar;

  Problem_Set<Key> problems;
  Device_Resources resources;
  std::string plan_cache;
  bool rates;

  Json finish(Planner_Type &planner) {
    if (!plan_cache.empty()) planner.save_plans_file(plan_cache);
    return planner.make_statistics().json(rates ? sizeof(Scalar) : 0);
  }

 public:
  explicit Test_Harness(const Json &problem_json, uint64_t seed = 0xB200, std::string plan_cache = "", bool rates = true)
      : problems(problem_json), resources(seed), plan_cache(std::move(plan_cache)), rates(rates) {}

  // every problem x every plan x repetitions
  Json run_exhaustive(int repetitions) {
    Planner_Type planner;
    for (auto &problem : problems.get_problems()) {
      Params input = form_input<Scalar>(problem);
      for (auto &opts : Opts::enumerate()) {
        resources.scratch_space.template grow_to_fit<char>(planner.calculate_workspace(input, opts));
        for (int i = 0; i < repetitions; i++) {
          planner.execute(input, opts, resources.scratch_space, resources.s);
          resources.sync();
        }
      }
      // every plan has its samples now: converge, so that the cache written below holds this problem's winner
      if (!plan_cache.empty() && repetitions > 0) (void)planner.create_plan(problem);
    }
    return finish(planner);
  }

  // what a user of the tuner sees: explore, converge, then keep running the winner.  With a plan cache the exploration
  // is skipped for every problem a previous run already tuned.
  Json run_autotune(int repetitions) {
    Planner_Type planner;
    if (!plan_cache.empty()) planner.load_plans_file(plan_cache);
    for (auto &problem : problems.get_problems()) {
      Params input = form_input<Scalar>(problem);
      for (int i = 0; i < repetitions; i++) {
        auto opts = planner.create_plan(problem);
        resources.scratch_space.template grow_to_fit<char>(planner.calculate_workspace(input, opts));
        planner.execute(input, opts, resources.scratch_space, resources.s);
        resources.sync();
      }
    }
    return finish(planner);
  }

  // tight leading dimensions, alpha = 1, beta = 0 (reference test_harness.h:232-248)
  template <typename T>
  GEMM_Inputs<T> form_input(GEMM_Key key) {
    const bool an = key.transa == gpu::BLAS_OP_N, bn = key.transb == gpu::BLAS_OP_N;
    MatrixDims Ad(an ? key.m : key.k, an ? key.k : key.m, an ? key.m : key.k);
    MatrixDims Bd(bn ? key.k : key.n, bn ? key.n : key.k, bn ? key.k : key.n);
    MatrixDims Cd(key.m, key.n, key.m);
    auto ms = resources.allocate_matrices<T>({Ad, Bd, Cd});
    return GEMM_Inputs<T>(resources.handle, key.transa, key.transb, ms[0], ms[1], ms[2], T(1), T(0));
  }

  // tight leading dimensions, alpha = 1, beta = 0 (reference test_harness.h:264-278)
  template <typename T>
  SYRK_Inputs<T> form_input(SYRK_Key key) {
    const bool an = key.trans == gpu::BLAS_OP_N;
    MatrixDims Ad(an ? key.n : key.k, an ? key.k : key.n, an ? key.n : key.k);
    MatrixDims Cd(key.n, key.n, key.n);
    auto ms = resources.allocate_matrices<T>({Ad, Cd});
    return SYRK_Inputs<T>(resources.handle, key.uplo, key.trans, ms[0], ms[1], T(1), T(0));
  }

  // tight leading dimensions, alpha = 1 (reference test_harness.h:250-262).  The fill is U(0,1] like the
  // reference's; it does not condition A, so timings are meaningful and values are not.
  template <typename T>
  TRSM_Inputs<T> form_input(TRSM_Key key) {
    const size_t na = key.side == gpu::BLAS_SIDE_LEFT ? key.m : key.n;
    MatrixDims Ad(na, na, na), Bd(key.m, key.n, key.m);
    auto ms = resources.allocate_matrices<T>({Ad, Bd});
    return TRSM_Inputs<T>(resources.handle, key.side, key.uplo, key.trans, key.diag, ms[0], ms[1], T(1));
  }
};


// method "move": MatrixMove alone (reference src/matrix_ops/matrixop.h:307-344), the matrix_op half of the metric.
// Problems {"m","n","transpose": bool, "pad": elements}; output per problem: "times" (ms, event pairs like
// Executor::execute) and "gbs" = 2 m n sizeof(T) / time.
template <typename T>
inline Json run_moves(const Json &problem_json, int repetitions, uint64_t seed) {
  Device_Resources resources(seed);
  Json out = Json::array();
  for (auto &pj : problem_json.items()) {
    const size_t m = (size_t)pj["m"].get<int>(), n = (size_t)pj["n"].get<int>();
    const bool transpose = pj.contains("transpose") && pj["transpose"].get<bool>();
    const size_t pad = pj.contains("pad") ? (size_t)pj["pad"].get<int>() : 1;
    auto src = resources.allocate_matrices<T>({MatrixDims(m, n, m)});
    MatrixMove<T> move(std::make_unique<NoOp<T>>(src[0]), T(1), transpose, pad ? pad : 1);
    resources.scratch_space.template grow_to_fit<char>(move.output_space_req() * sizeof(T));
    Workspace target(resources.scratch_space, 0, move.output_space_req() * sizeof(T));
    Timer_Bank bank;
    for (int i = 0; i < repetitions; i++) {
      Device_Timer timer([&](const Stream &) { move.execute(resources.handle, target, Workspace()); }, resources.s, Device_Timer::SEMI_SYNCHRONOUS);
      bank.append(timer);
      resources.sync();
    }
    bank.synchronize();
    Json entry = Json::object(), opt = Json::object();
    entry["key"] = pj;
    opt["option"] = Json::object();
    opt["times"] = Json(bank.get_times());
    std::vector<double> gbs;
    for (float ms : bank.get_times()) gbs.push_back(2.0 * m * n * sizeof(T) / (double(ms) * 1e-3) * 1e-9);
    opt["gbs"] = Json(gbs);
    Json opts = Json::array();
    opts.push_back(opt);
    entry["options"] = opts;
    out.push_back(entry);
  }
  return out;
}

}  // namespace rtat

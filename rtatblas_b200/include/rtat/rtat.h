// rtat.h — facade: lazily constructed planners per routine and precision (reference src/rtat.h).
// The reference spells the accessors as in-class explicit specialisations, which GCC rejects
// (SURVEY §8 a1); here they are plain if-constexpr members with the same names.
#pragma once
#include <memory>
#include <type_traits>
#include "gemm.h"
#include "planning_system.h"
#include "syrk.h"
#include "trsm.h"

namespace rtat {

template <typename T>
class Lazy {
  std::unique_ptr<T> val;

 public:
  operator T &() {
    if (!val) val = std::make_unique<T>();
    return *val;
  }
  bool constructed() const { return bool(val); }
};

class rtat {
  Lazy<Planning_System<GEMM_Executor<double>>> dgemm_planner;
  Lazy<Planning_System<GEMM_Executor<float>>> sgemm_planner;
  Lazy<Planning_System<GEMM_Executor_Pad<double>>> dgemm_pad_planner;
  Lazy<Planning_System<GEMM_Executor_Pad<float>>> sgemm_pad_planner;
  Lazy<Planning_System<TRSM_Executor<double>>> dtrsm_planner;
  Lazy<Planning_System<TRSM_Executor<float>>> strsm_planner;
  Lazy<Planning_System<SYRK_Executor<double>>> dsyrk_planner;
  Lazy<Planning_System<SYRK_Executor<float>>> ssyrk_planner;

 public:
  template <typename T>
  Planning_System<GEMM_Executor<T>> &gemm_planner() {
    static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "GEMM planners exist for double and float");
    if constexpr (std::is_same_v<T, double>) return dgemm_planner;
    else return sgemm_planner;
  }
  // the 64-plan space (transposes x pads); the reference only reaches it through run_tests
  template <typename T>
  Planning_System<GEMM_Executor_Pad<T>> &gemm_pad_planner() {
    static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "GEMM planners exist for double and float");
    if constexpr (std::is_same_v<T, double>) return dgemm_pad_planner;
    else return sgemm_pad_planner;
  }

  template <typename T>
  Planning_System<TRSM_Executor<T>> &trsm_planner() {
    static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "TRSM planners exist for double and float");
    if constexpr (std::is_same_v<T, double>) return dtrsm_planner;
    else return strsm_planner;
  }
  template <typename T>
  Planning_System<SYRK_Executor<T>> &syrk_planner() {
    static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "SYRK planners exist for double and float");
    if constexpr (std::is_same_v<T, double>) return dsyrk_planner;
    else return ssyrk_planner;
  }

  // Tuned GEMM in one call: pick (or keep exploring) the plan for this shape, make sure the
  // workspace can hold it, enqueue it on the stream bound to the handle.
  template <typename T>
  GEMM_Options gemm(GEMM_Inputs<T> inputs, ManagedWorkspace &scratch, Stream s) {
    auto &planner = gemm_planner<T>();
    GEMM_Options plan = planner.create_plan(GEMM_Key(inputs));
    scratch.grow_to_fit<char>(planner.calculate_workspace(inputs, plan));
    planner.execute(inputs, plan, scratch, s);
    return plan;
  }
};

}  // namespace rtat

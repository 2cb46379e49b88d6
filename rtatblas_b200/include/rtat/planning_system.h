// planning_system.h — run-time plan selection (reference src/planning/planning_system.h).
//
// For each tuning key the planner hands out every candidate plan until each has
// `tests_until_converge` timed samples, then commits to the plan with the lowest mean time.
// New here: the converged choices can be saved to / pre-seeded from a JSON plan cache, so a
// process does not have to re-explore shapes a previous run already tuned.
#pragma once
#include <fstream>
#include <limits>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include "gemm.h"
#include "gpu-api.h"
#include "planner_statistics.h"
#include "predicates.h"

namespace rtat {

// Restricts the plan space per key: apply(key) lists the plans of Opts::enumerate() that the
// predicate accepts, in enumeration order.
template <typename Key, typename Opts>
class Option_Filter {
  Predicate<std::pair<Opts, Key>> accept;

 public:
  Option_Filter() : accept([](std::pair<Opts, Key>) { return true; }) {}
  Option_Filter(Predicate<std::pair<Opts, Key>> accept) : accept(std::move(accept)) {}
  std::vector<Opts> apply(Key key) const {
    std::vector<Opts> kept;
    for (auto &opts : Opts::enumerate())
      if (accept(std::make_pair(opts, key))) kept.push_back(opts);
    return kept;
  }
};

template <typename Executor_Type>
class Planning_System {
 public:
  using Params = typename Executor_Type::Params_T;
  using Key = typename Executor_Type::Key_T;
  using Opts = typename Executor_Type::Opts_T;

 protected:
  Executor_Type executor;
  const Option_Filter<Key, Opts> opt_filter;
  size_t tests_until_converge = 1;
  // relative margin within which the default plan is kept over a faster-looking one; 0 = the reference's strict minimum
  // of the means (planning_system.h:84-95).  Drivers that sample short problems opt in with set_default_preference.
  float default_preference = 0.f;
  std::map<Key, Opts> converged_plans;
  Device_Timer::Mode sync_mode = Device_Timer::ASYNCHRONOUS;

  // what to run when the caller's workspace cannot hold the chosen plan
  Opts degrade_plan(Params, Opts, Workspace) { return Opts::default_opts(); }

 public:
  Planning_System() = default;
  Planning_System(Option_Filter<Key, Opts> opt_filter) : opt_filter(opt_filter) {}
  virtual ~Planning_System() = default;

  virtual Opts create_plan(Key key) {
    auto hit = converged_plans.find(key);
    if (hit != converged_plans.end()) return hit->second;

    std::map<Opts, Timer_Bank> &timings = executor.get_timings(key);
    const std::vector<Opts> candidates = opt_filter.apply(key);
    for (auto &opts : candidates)  // exploration: first plan still short of samples
      if (timings[opts].size() < tests_until_converge) return opts;

    Opts best = Opts::default_opts();
    float best_ms = std::numeric_limits<float>::max();
    for (auto &opts : candidates) {
      Timer_Bank &bank = timings[opts];
      bank.synchronize();  // blocks until the exploration samples have finished on the device
      const std::vector<float> &ts = bank.get_times();
      const float mean = float(std::accumulate(ts.begin(), ts.end(), 0.0) / ts.size());
      if (mean < best_ms) { best = opts; best_ms = mean; }
    }
    // Optional (off by default): event timings of short problems jitter by a few per cent; with a margin set, a plan that
    // spends extra passes and workspace must win by more than that to displace the default (fully fused) plan.
    if (default_preference > 0.f) {
      const Opts dflt = Opts::default_opts();
      auto it = timings.find(dflt);
      if (it != timings.end() && it->second.size() > 0) {
        const std::vector<float> &ts = it->second.get_times();
        const float mean = float(std::accumulate(ts.begin(), ts.end(), 0.0) / ts.size());
        if (mean <= best_ms * (1.f + default_preference)) best = dflt;
      }
    }
    converged_plans.insert_or_assign(key, best);
    return best;
  }

  void set_sync_mode(Device_Timer::Mode mode) { sync_mode = mode; }
  void set_tests_until_converge(size_t n) { tests_until_converge = n ? n : 1; }
  void set_default_preference(float margin) { default_preference = margin > 0.f ? margin : 0.f; }

  // `s` must be the stream bound to params.handle (timing events are recorded on it).
  void execute(Params params, Opts opts, Workspace space, Stream s) {
    if (space.size<char>() < executor.calculate_workspace(params, opts)) opts = degrade_plan(params, opts, space);
    Device_Timer::Mode mode = sync_mode;
    // while a plan is still being sampled, never let earlier queued work leak into its timing
    if (mode == Device_Timer::ASYNCHRONOUS && executor.get_timings(Key(params))[opts].size() < tests_until_converge)
      mode = Device_Timer::SEMI_SYNCHRONOUS;
    executor.execute(params, opts, space, s, mode);
  }

  size_t calculate_workspace(Params params, Opts opts) { return executor.calculate_workspace(params, opts); }

  Planner_Statistics<Key, Opts> make_statistics() {
    std::map<Key, std::map<Opts, std::vector<float>>> samples;
    for (auto &[key, per_plan] : executor.get_timings())
      for (auto &[opt, bank] : per_plan) {
        bank.synchronize();
        samples[key][opt] = bank.get_times();
      }
    return Planner_Statistics<Key, Opts>(samples);
  }

  // ---- JSON plan cache: [ {"key": {...}, "option": {...}}, ... ] -----------------------------------
  const std::map<Key, Opts> &get_converged_plans() const { return converged_plans; }
  Json save_plans() const {
    Json out = Json::array();
    for (auto &[key, opt] : converged_plans) {
      Json e = Json::object();
      e["key"] = to_json(key);
      e["option"] = to_json(opt);
      out.push_back(e);
    }
    return out;
  }
  // On-disk form of the same cache (reference src/app/runner.h:81-82 writes its statistics the same way: one JSON
  // document per file).  save: overwrite `path`; load: missing file = empty cache (0), malformed file throws.
  void save_plans_file(const std::string &path) const {
    std::ofstream out(path, std::ios::trunc);
    if (!out) throw std::runtime_error("plan cache: cannot write " + path);
    out << save_plans().dump(1) << "\n";
  }
  size_t load_plans_file(const std::string &path) {
    std::ifstream in(path);
    if (!in) return 0;
    std::stringstream text;
    text << in.rdbuf();
    return load_plans(Json::parse(text.str()));
  }
  // returns the number of plans taken over; entries whose plan the filter rejects are skipped
  size_t load_plans(const Json &cache) {
    size_t taken = 0;
    for (auto &e : cache.items()) {
      Key key = from_json<Key>(e["key"]);
      Opts opt = from_json<Opts>(e["option"]);
      bool allowed = false;
      for (auto &cand : opt_filter.apply(key)) allowed = allowed || std::string(cand) == std::string(opt);
      if (!allowed) continue;
      converged_plans.insert_or_assign(key, opt);
      taken++;
    }
    return taken;
  }
};

using GEMM_Planner = Planning_System<GEMM_Executor<double>>;
using SGEMM_Planner = Planning_System<GEMM_Executor<float>>;

}  // namespace rtat

// planner_statistics.h — per-(problem, plan) timing summary and its JSON form
// (reference src/planning/planner_statistics.h).  JSON layout (run_tests output):
//   [ {"key": {...}, "options": [ {"option": {...}, "times": [ms, ...]} , ...] }, ... ]
// in map order, i.e. key-text order then plan-text order.
#pragma once
#include <algorithm>
#include <map>
#include <numeric>
#include <type_traits>
#include <utility>
#include <vector>
#include "json_encoding.h"

namespace rtat {

namespace detail {
template <typename Key, typename = void>
struct has_bytecount : std::false_type {};
template <typename Key>
struct has_bytecount<Key, std::void_t<decltype(std::declval<const Key &>().bytecount(size_t(8)))>> : std::true_type {};
}  // namespace detail

template <typename Key, typename Opts>
class Planner_Statistics {
  using Table = std::map<Key, std::map<Opts, std::vector<float>>>;
  Table times;
  std::map<Key, std::map<Opts, float>> means;
  std::map<Key, std::map<Opts, size_t>> counts;

 public:
  explicit Planner_Statistics(Table samples) : times(std::move(samples)) {
    for (auto &[key, per_plan] : times)
      for (auto &[opt, ts] : per_plan) {
        counts[key][opt] = ts.size();
        means[key][opt] = ts.empty() ? 0.f : float(std::accumulate(ts.begin(), ts.end(), 0.0) / ts.size());
      }
  }

  const Table &get_times() const { return times; }
  const std::map<Key, std::map<Opts, float>> &get_means() const { return means; }
  const std::map<Key, std::map<Opts, size_t>> &get_counts() const { return counts; }

  // FLOP/s per sample (needs Key::flopcount(); the reference's version of this cannot be
  // instantiated, planner_statistics.h:35-48).
  Table get_floprates() const {
    Table rates;
    for (auto &[key, per_plan] : times)
      for (auto &[opt, ts] : per_plan) {
        auto &out = rates[key][opt];
        for (float ms : ts) out.push_back(float(key.flopcount() / (double(ms) * 1e-3)));
      }
    return rates;
  }

  // fastest plan by mean time for every key that has samples
  std::map<Key, Opts> best_plans() const {
    std::map<Key, Opts> best;
    for (auto &[key, per_plan] : means) {
      const Opts *arg = nullptr;
      float lo = 0.f;
      for (auto &[opt, mean] : per_plan)
        if (counts.at(key).at(opt) > 0 && (!arg || mean < lo)) { arg = &opt; lo = mean; }
      if (arg) best.emplace(key, *arg);
    }
    return best;
  }

  // `elem_bytes` > 0 adds two derived columns per plan next to "times": "tflops" (Key::flopcount / time, what the
  // reference's get_floprates was meant to give, planner_statistics.h:35-48) and, for keys that define bytecount,
  // "gbs" (algorithmic bytes of the call / time).  Like nlohmann's default-constructed values in the reference, an empty
  // table prints as null and a key without plans gets "options": null.
  Json json(size_t elem_bytes = 0) const {
    Json out = times.empty() ? Json() : Json::array();
    for (auto &[key, per_plan] : times) {
      Json entry = Json::object();
      entry["key"] = to_json(key);
      Json plans = per_plan.empty() ? Json() : Json::array();
      for (auto &[opt, ts] : per_plan) {
        Json p = Json::object();
        p["option"] = to_json(opt);
        p["times"] = Json(ts);
        if (elem_bytes > 0) {
          std::vector<double> tf, gb;
          for (float ms : ts) {
            tf.push_back(key.flopcount() / (double(ms) * 1e-3) * 1e-12);
            if constexpr (detail::has_bytecount<Key>::value) gb.push_back(key.bytecount(elem_bytes) / (double(ms) * 1e-3) * 1e-9);
          }
          p["tflops"] = Json(tf);
          if constexpr (detail::has_bytecount<Key>::value) p["gbs"] = Json(gb);
        }
        plans.push_back(p);
      }
      entry["options"] = plans;
      out.push_back(entry);
    }
    return out;
  }
};

}  // namespace rtat

// the reference declares this template in the global namespace (planner_statistics.h:8)
using rtat::Planner_Statistics;

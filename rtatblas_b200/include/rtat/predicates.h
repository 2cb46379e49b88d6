// predicates.h — composable filters over (option, key) pairs (reference src/methods/predicates.h).
#pragma once
#include <algorithm>
#include <functional>
#include <vector>

namespace rtat {

template <typename T>
using Predicate = std::function<bool(T)>;

template <typename T>
Predicate<T> disjunction(std::vector<Predicate<T>> preds) {
  return [preds = std::move(preds)](T v) { return std::any_of(preds.begin(), preds.end(), [&](const Predicate<T> &p) { return p(v); }); };
}

template <typename T>
Predicate<T> conjunction(std::vector<Predicate<T>> preds) {
  return [preds = std::move(preds)](T v) { return std::all_of(preds.begin(), preds.end(), [&](const Predicate<T> &p) { return p(v); }); };
}

}  // namespace rtat

// timer_bank.h — FIFO of in-flight Device_Timers plus the samples already read
// (reference src/timing/timer_bank.{h,cpp}).
#pragma once
#include <deque>
#include <vector>
#include "device_timer.h"

namespace rtat {

class Timer_Bank {
  std::deque<Device_Timer> pending;
  std::vector<float> times;

  // move every finished timer (in submission order) from `pending` to `times`
  void drain_finished() {
    while (!pending.empty()) {
      auto t = pending.front().query_time();
      if (!t) break;
      times.push_back(*t);
      pending.pop_front();
    }
  }

 public:
  void append(Device_Timer &timer) { pending.push_back(std::move(timer)); }
  void synchronize() {
    for (auto &t : pending) times.push_back(t.time());
    pending.clear();
  }
  size_t size() { return pending.size() + times.size(); }
  size_t completed() {
    drain_finished();
    return times.size();
  }
  const std::vector<float> &get_times() {
    drain_finished();
    return times;
  }
};

}  // namespace rtat

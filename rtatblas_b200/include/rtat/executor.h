// executor.h — runs one (problem, plan) pair and keeps its timing (reference src/methods/executor.h).
#pragma once
#include <map>
#include "timer_bank.h"
#include "workspace.h"

namespace rtat {

template <typename Params, typename Key, typename Opts>
class Executor {
 public:
  typedef Params Params_T;
  typedef Key Key_T;
  typedef Opts Opts_T;

  Executor() = default;
  virtual ~Executor() = default;

  // Enqueue the plan on `s` (which must be the stream bound to params.handle: the timing events
  // are recorded there) and log the sample under (key, plan); at most log_size_limit samples per
  // pair are kept, later runs execute untimed-logged (reference executor.h:18-33).
  virtual void execute(Params params, Opts opts, Workspace space, Stream s, Device_Timer::Mode sync = Device_Timer::ASYNCHRONOUS) {
    if (!warm) {
      warmup(params, opts, s);
      warm = true;
    }
    Timer_Bank &bank = timer_log[Key(params)][opts];
    if (bank.size() >= log_size_limit) {
      // the log for this (key, plan) is full: the reference still brackets the run with events and drops the timer
      // (executor.h:31-32); nothing observable changes if the events are not recorded at all
      if (sync != Device_Timer::ASYNCHRONOUS) gpuAssert(gpu::DeviceSynchronize());
      internal_execute(params, opts, space, s);
      if (sync == Device_Timer::SYNCHRONOUS) s.synchronize();
      return;
    }
    Device_Timer timer([&](const Stream &on) { internal_execute(params, opts, space, on); }, s, sync);
    bank.append(timer);
  }

  std::map<Key, std::map<Opts, Timer_Bank>> &get_timings() { return timer_log; }
  std::map<Opts, Timer_Bank> &get_timings(Key key) { return timer_log[key]; }

  // bytes of scratch the plan needs for this problem
  virtual size_t calculate_workspace(Params params, Opts opts) { return opts.form_operation(params)->workspace_req_bytes(); }

 protected:
  virtual void internal_execute(Params params, Opts opts, Workspace space, Stream) {
    auto operation = opts.form_operation(params);
    if (operation->workspace_req_bytes() > space.size<char>()) throw "internal_execute: Insufficient workspace";
    operation->execute(params.handle, Workspace(), space);
  }
  virtual void warmup(Params, Opts, Stream) = 0;

  std::map<Key, std::map<Opts, Timer_Bank>> timer_log;
  const size_t log_size_limit = 100;
  bool warm = false;
};

}  // namespace rtat

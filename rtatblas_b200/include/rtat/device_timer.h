// device_timer.h — event-pair timing of work enqueued on a stream (reference src/timing/device_timer.{h,cpp}).
#pragma once
#include <memory>
#include <optional>
#include "gpu-api.h"

namespace rtat {

class Device_Timer {
 public:
  enum Mode {
    ASYNCHRONOUS = 0,      // record, enqueue, record
    SEMI_SYNCHRONOUS = 1,  // drain the device first, so nothing queued earlier pollutes the sample
    SYNCHRONOUS = 2        // ... and also wait for the end event
  };

  // Times f(s).  The events are recorded on `s`, which must be the stream the work inside f is
  // enqueued on (for BLAS calls: the stream bound to the handle).
  template <typename Func>
  Device_Timer(Func f, Stream s, Mode mode = ASYNCHRONOUS) : start(std::make_unique<Event>()), end(std::make_unique<Event>()) {
    if (mode != ASYNCHRONOUS) gpuAssert(gpu::DeviceSynchronize());
    start->record(s);
    f(s);
    end->record(s);
    if (mode == SYNCHRONOUS) end->synchronize();
  }

  // Non-blocking: elapsed ms once both events have completed; the events are released after the
  // first successful read.
  std::optional<float> query_time() {
    if (ms < 0.f) {
      if (!start->query() || !end->query()) return std::nullopt;
      ms = Event::elapsed_time(*start, *end);
      start.reset();
      end.reset();
    }
    return ms;
  }

  // Blocking.
  float time() {
    if (ms < 0.f) {
      start->synchronize();
      end->synchronize();
    }
    return *query_time();
  }

 private:
  std::unique_ptr<Event> start, end;
  float ms = -1.f;
};

}  // namespace rtat

// host_pool.hpp — small persistent thread pool for the host side of the plugin boundary:
// packing physim's 80-byte AoS Entity records into pinned {x,y,z,m} staging and adding the
// returned accelerations into the caller's array.  The reference does this work on its single
// simulation thread (physim-core/src/pipeline.rs:134); here it only feeds the PCIe copies.
#pragma once

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pb200 {

class HostPool {
 public:
  static HostPool& instance() {
    static HostPool* pool = new HostPool();  // leaked on purpose: plugin may be unloaded at exit
    return *pool;
  }

  int threads() const { return n_threads_; }

  // Runs fn(part, parts) once on each of `parts` <= threads() threads (part 0 on the caller); blocks.
  // One wake-up of the pool serves a whole multi-chunk pipeline: the parts walk the chunks together and
  // hand chunk boundaries to each other through atomics instead of one parallel_for per chunk.
  void parallel_parts(int parts, const std::function<void(int, int)>& fn) {
    parts = std::max(1, std::min(parts, n_threads_));
    if (parts <= 1 || workers_.empty()) {
      fn(0, 1);
      return;
    }
    std::unique_lock<std::mutex> call_lock(call_mu_);  // one parallel region at a time
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      parts_ = parts;
      pending_ = parts - 1;
      ++generation_;
    }
    cv_.notify_all();
    fn(0, parts);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

  // how many parts a loop over n items is worth
  int parts_for(size_t n, size_t min_per_thread) const {
    return static_cast<int>(std::min<size_t>(n_threads_, std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread))));
  }

  // Runs fn(begin, end) over [0, n) split into contiguous ranges, one per part; blocks.
  void parallel_for(size_t n, size_t min_per_thread, const std::function<void(size_t, size_t)>& fn) {
    parallel_parts(parts_for(n, min_per_thread), [&](int part, int parts) {
      const size_t per = (n + size_t(parts) - 1) / size_t(parts);
      const size_t b = std::min(n, per * size_t(part)), e = std::min(n, b + per);
      if (b < e) fn(b, e);
    });
  }

 private:
  HostPool() {
    int n = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* e = std::getenv("PB200_HOST_THREADS")) n = std::atoi(e);
    n_threads_ = std::max(1, std::min(n, 32));
    for (int i = 1; i < n_threads_; ++i) workers_.emplace_back([this, i] { worker(i); });
    for (auto& t : workers_) t.detach();
  }

  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return generation_ != seen; });
      seen = generation_;
      const bool mine = id < parts_;
      lk.unlock();
      if (mine) {
        (*fn_)(id, parts_);
        lk.lock();
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }

  int n_threads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int parts_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
};

}  // namespace pb200

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

  // Runs fn(begin, end) over [0, n) split into contiguous chunks, one per worker; blocks.
  void parallel_for(size_t n, size_t min_per_thread, const std::function<void(size_t, size_t)>& fn) {
    int use = static_cast<int>(std::min<size_t>(n_threads_, std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread))));
    if (use <= 1 || workers_.empty()) {
      fn(0, n);
      return;
    }
    std::unique_lock<std::mutex> call_lock(call_mu_);  // one parallel_for at a time
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      total_ = n;
      parts_ = use;
      pending_ = use - 1;
      ++generation_;
    }
    cv_.notify_all();
    run_part(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() {
    int n = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* e = std::getenv("PB200_HOST_THREADS")) n = std::atoi(e);
    n_threads_ = std::max(1, std::min(n, 32));
    for (int i = 1; i < n_threads_; ++i) workers_.emplace_back([this, i] { worker(i); });
    for (auto& t : workers_) t.detach();
  }

  void run_part(int part) {
    const size_t per = (total_ + parts_ - 1) / parts_;
    const size_t b = std::min(total_, per * part), e = std::min(total_, b + per);
    if (b < e) (*fn_)(b, e);
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
        run_part(id);
        lk.lock();
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }

  int n_threads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t, size_t)>* fn_ = nullptr;
  size_t total_ = 0;
  int parts_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
};

}  // namespace pb200

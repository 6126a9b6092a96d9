// Exercises physim_b200/csrc/host_pool.hpp without a GPU: parallel_for coverage and the chunk hand-off
// pattern engine.cu builds on parallel_parts (every part does its share of every chunk, part 0 waits for a
// chunk's shares before "sending" it, the other parts wait for "arrivals" published by part 0).
#include <atomic>
#include <cstdio>
#include <numeric>
#include <vector>

#include "../../physim_b200/csrc/host_pool.hpp"

using pb200::HostPool;

static int fail(const char* what) {
  std::printf("FAIL %s\n", what);
  return 1;
}

int main() {
  HostPool& pool = HostPool::instance();
  std::printf("threads %d\n", pool.threads());
  // 1. parallel_for touches every index exactly once, for awkward sizes
  for (size_t n : {size_t(0), size_t(1), size_t(7), size_t(4096), size_t(4097), size_t(100003), size_t(1) << 20}) {
    std::vector<unsigned char> hit(n, 0);
    pool.parallel_for(n, 1, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i) ++hit[i];
    });
    for (size_t i = 0; i < n; ++i)
      if (hit[i] != 1) return fail("parallel_for coverage");
  }
  // 2. parallel_parts: each part id in [0, parts) runs exactly once, parts <= threads, part 0 on the caller
  for (int want : {1, 2, 3, 64}) {
    std::vector<std::atomic<int>> seen(64);
    for (auto& s : seen) s.store(0);
    std::atomic<int> parts_seen{0};
    const auto caller = std::this_thread::get_id();
    std::atomic<int> bad{0};
    pool.parallel_parts(want, [&](int part, int parts) {
      parts_seen.store(parts);
      if (part < 0 || part >= parts) bad.store(1);
      else seen[part].fetch_add(1);
      if (part == 0 && std::this_thread::get_id() != caller) bad.store(2);
    });
    const int parts = parts_seen.load();
    if (bad.load() || parts < 1 || parts > std::min(want, pool.threads())) return fail("parallel_parts ids");
    for (int p = 0; p < parts; ++p)
      if (seen[p].load() != 1) return fail("parallel_parts each part once");
  }
  // 3. the upload pattern: chunks packed by all parts, "sent" by part 0 only when complete and in order
  {
    const int chunks = 16;
    const size_t n = 1000003;
    std::vector<int> data(n, 0);
    std::atomic<int> packed[chunks];
    for (auto& a : packed) a.store(0);
    std::vector<long long> sent_sum(chunks, -1);
    pool.parallel_parts(pool.parts_for(n, 4096), [&](int part, int parts) {
      for (int c = 0; c < chunks; ++c) {
        const size_t b = n * size_t(c) / chunks, e = n * size_t(c + 1) / chunks, m = e - b;
        const size_t sb = b + m * size_t(part) / size_t(parts), se = b + m * size_t(part + 1) / size_t(parts);
        for (size_t i = sb; i < se; ++i) data[i] = int(i % 1000);
        packed[c].fetch_add(1, std::memory_order_release);
        if (part != 0) continue;
        while (packed[c].load(std::memory_order_acquire) < parts) std::this_thread::yield();
        long long s = 0;
        for (size_t i = b; i < e; ++i) s += data[i];  // the "DMA" reads the finished chunk
        sent_sum[c] = s;
      }
    });
    for (int c = 0; c < chunks; ++c) {
      const size_t b = n * size_t(c) / chunks, e = n * size_t(c + 1) / chunks;
      long long s = 0;
      for (size_t i = b; i < e; ++i) s += long(i % 1000);
      if (sent_sum[c] != s) return fail("upload pattern: chunk sent before it was complete");
    }
  }
  // 4. the download pattern: part 0 publishes arrivals, every part consumes its share of each arrived chunk
  {
    const int chunks = 16;
    const size_t n = 500009;
    std::vector<int> src(n, -1), dst(n, 0);
    std::atomic<int> arrived{0};
    std::atomic<int> early{0};
    pool.parallel_parts(pool.parts_for(n, 4096), [&](int part, int parts) {
      for (int c = 0; c < chunks; ++c) {
        const size_t b = n * size_t(c) / chunks, e = n * size_t(c + 1) / chunks, m = e - b;
        if (part == 0) {
          for (size_t i = b; i < e; ++i) src[i] = int(i & 0xffff);  // the "copy" lands
          arrived.store(c + 1, std::memory_order_release);
        } else {
          while (arrived.load(std::memory_order_acquire) < c + 1) std::this_thread::yield();
        }
        const size_t sb = b + m * size_t(part) / size_t(parts), se = b + m * size_t(part + 1) / size_t(parts);
        for (size_t i = sb; i < se; ++i) {
          if (src[i] < 0) early.store(1);
          dst[i] = src[i] + 1;
        }
      }
    });
    if (early.load()) return fail("download pattern: a share was consumed before its chunk arrived");
    for (size_t i = 0; i < n; ++i)
      if (dst[i] != int(i & 0xffff) + 1) return fail("download pattern coverage");
  }
  // 5. many short regions back to back (generation counter / wake-up races)
  {
    std::atomic<long long> total{0};
    for (int r = 0; r < 2000; ++r)
      pool.parallel_parts(8, [&](int part, int) { total.fetch_add(part + 1); });
    const int parts = std::min(8, pool.threads());
    if (total.load() != 2000LL * parts * (parts + 1) / 2) return fail("back-to-back regions");
  }
  std::printf("host_pool ok\n");
  return 0;
}

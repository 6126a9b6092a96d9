// Host-side microbenchmark of the Entity pack / unpack loops of engine.cu (no GPU involved):
// chunked (8 parallel_for calls) against one call, to separate pool wake-up cost from memory bandwidth.
//   g++ -O3 -march=native -std=c++17 -pthread tools/scratch/host_pack_bench.cpp -o /tmp/host_pack_bench
#include <emmintrin.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../physim_b200/csrc/host_pool.hpp"
struct Entity { double x, y, z, vx, vy, vz, radius, mass; size_t id; bool fixed; };
struct d4 { double x, y, z, w; };
using namespace pb200;
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void pack(const Entity* state, size_t n, d4* pos, uint8_t* fixed) {
  HostPool::instance().parallel_for(n, 4096, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      const Entity& s = state[i];
      double* d = (double*)(pos + i);
      _mm_stream_pd(d, _mm_set_pd(s.y, s.x));
      _mm_stream_pd(d + 2, _mm_set_pd(s.mass, s.z));
      fixed[i] = s.fixed ? 1 : 0;
    }
    _mm_sfence();
  });
}
static void unpack(const Entity* entities, Entity* out, size_t n, const double* out6) {
  HostPool::instance().parallel_for(n, 4096, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      const double* o = out6 + 6 * i;
      const double* src = (const double*)(entities + i);
      double* dst = (double*)(out + i);
      _mm_stream_pd(dst, _mm_loadu_pd(o));
      _mm_stream_pd(dst + 2, _mm_loadu_pd(o + 2));
      _mm_stream_pd(dst + 4, _mm_loadu_pd(o + 4));
      _mm_stream_pd(dst + 6, _mm_loadu_pd(src + 6));
      _mm_stream_pd(dst + 8, _mm_loadu_pd(src + 8));
    }
    _mm_sfence();
  });
}
int main() {
  size_t n = 1000004;
  std::vector<Entity> st(n), out(n);
  for (size_t i = 0; i < n; ++i) { st[i].x = i; st[i].mass = 1; st[i].fixed = false; }
  d4* pos = (d4*)aligned_alloc(64, n * 32);
  uint8_t* fx = (uint8_t*)malloc(n);
  double* o6 = (double*)aligned_alloc(64, n * 48);
  memset(pos, 0, n * 32); memset(fx, 0, n); memset(o6, 0, n * 48);
  printf("threads %d\n", HostPool::instance().threads());
  for (int rep = 0; rep < 6; ++rep) {
    double t0 = now();
    for (int c = 0; c < 8; ++c) { size_t b = n * c / 8, e = n * (c + 1) / 8; pack(st.data() + b, e - b, pos + b, fx + b); }
    double t1 = now();
    pack(st.data(), n, pos, fx);
    double t2 = now();
    for (int c = 0; c < 8; ++c) { size_t b = n * c / 8, e = n * (c + 1) / 8; unpack(st.data() + b, out.data() + b, e - b, o6 + 6 * b); }
    double t3 = now();
    unpack(st.data(), out.data(), n, o6);
    double t4 = now();
    double t5 = now();
    for (int c = 0; c < 64; ++c) HostPool::instance().parallel_for(1 << 20, 4096, [&](size_t, size_t) {});
    double t6 = now();
    printf("pack chunked %.3f single %.3f | unpack chunked %.3f single %.3f ms | empty parallel_for %.1f us\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, (t6 - t5) / 64 * 1e3);
  }
}

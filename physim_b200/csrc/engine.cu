// engine.cu — host side of libphysim_b200.so: the handles behind include/physim_b200.h section 3.
//
//   TransformObj  ~ AstroElement / AstroOctreeElement / SimpleAstroElement (astro/src/transformers.rs)
//   Verlet     ~ integrators/src/verlet.rs
//   Sim        ~ the simulation thread's loop (physim-core/src/pipeline.rs:143-182) kept in HBM
//
// Nothing here touches the GPU until the first force evaluation: physim instantiates and drops
// every transform during plugin discovery (physim-core/src/plugin/discover.rs:376-387).
#include <emmintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include <atomic>
#include <functional>
#include <thread>

#include "common.cuh"
#include "host_pool.hpp"

namespace pb200 {

thread_local char g_error[512] = {0};
static thread_local int g_device = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
  if (std::getenv("PB200_VERBOSE")) std::fprintf(stderr, "[physim_b200] %s\n", g_error);
}

namespace {

constexpr size_t kParallelGrain = 1 << 12;

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

thread_local double g_pack_ms = 0.0, g_unpack_ms = 0.0;

struct GpuContext {
  bool ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t chunk_ev[16] = {};
  cudaError_t init(int dev) {
    if (ready) return cudaSetDevice(device);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
      set_error("no CUDA device available (%s); physim_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
      return e == cudaSuccess ? cudaErrorNoDevice : e;
    }
    device = dev;
    PB_CUDA(cudaSetDevice(device));
    PB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (auto& x : ev) PB_CUDA(cudaEventCreate(&x));
    for (auto& x : chunk_ev) PB_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    ready = true;
    return cudaSuccess;
  }
  void destroy() {
    if (!ready) return;
    cudaSetDevice(device);
    for (auto& x : ev)
      if (x) cudaEventDestroy(x);
    for (auto& x : chunk_ev)
      if (x) cudaEventDestroy(x);
    if (stream && own_stream) cudaStreamDestroy(stream);
    ready = false;
  }
};

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

struct TransformObj {
  GravityParams prm;
  std::mutex mu;
  GpuContext gpu;
  int device;
  GravityWorkspace ws;
  DevBuf d_pos, d_fixed;
  PinnedBuf h_pos, h_fixed, h_acc;
  LaunchStats ls;
  Pb200Stats stats;
  size_t last_n = 0;
  bool have_tree = false;

  ~TransformObj() {
    if (gpu.ready) {
      cudaSetDevice(gpu.device);
      ws.release_all();
      d_pos.release();
      d_fixed.release();
      h_pos.release();
      h_fixed.release();
      h_acc.release();
      gpu.destroy();
    }
  }
};

// pack Entity AoS -> pinned {x,y,z,m} (+ fixed bytes).  The staging buffer is written with
// non-temporal stores: it is read next by the DMA engine, not by a core.
inline void pack_positions_range(const Entity* state, size_t b, size_t e, double4* pos, uint8_t* fixed) {
  for (size_t i = b; i < e; ++i) {
    const Entity& s = state[i];
    double* d = reinterpret_cast<double*>(pos + i);  // 32-byte aligned (pinned base, 32 B records)
    _mm_stream_pd(d, _mm_set_pd(s.y, s.x));
    _mm_stream_pd(d + 2, _mm_set_pd(s.mass, s.z));
    if (fixed) fixed[i] = s.fixed ? 1 : 0;
  }
}

inline void pack_velocities_range(const Entity* state, size_t b, size_t e, double4* vel) {
  for (size_t i = b; i < e; ++i) {
    double* d = reinterpret_cast<double*>(vel + i);
    _mm_stream_pd(d, _mm_set_pd(state[i].vy, state[i].vx));
    _mm_stream_pd(d + 2, _mm_set_pd(0.0, state[i].vz));
  }
}

// Chunking of the host<->device pipeline: packing chunk c+1 overlaps the H2D copy of chunk c, and
// unpacking chunk c overlaps the D2H copy of chunk c+1.  The pool is woken ONCE per direction: its
// parts walk the chunks together, each doing its share of every chunk; part 0 (the calling thread, the
// only one that talks to CUDA) enqueues the copy of a chunk when all shares of it are packed, and
// publishes the arrival of a chunk for the parts that unpack it.  (One parallel_for per chunk costs
// ~20 us of wake-up each; measured on the 16-core host, tools/scratch/host_pack_bench.cpp.)
constexpr int kMaxChunks = 16;
inline int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
inline int chunk_count(size_t n) {
  static const int chunks = std::max(1, std::min(kMaxChunks, env_int("PB200_CHUNKS", 16)));  // (tuning runs)
  return n >= (size_t(1) << 18) ? chunks : 1;
}
inline bool per_chunk_regions() {  // PB200_HOST_PIPE=perchunk: one parallel_for per chunk (A/B runs)
  static const bool v = std::getenv("PB200_HOST_PIPE") && std::string(std::getenv("PB200_HOST_PIPE")) == "perchunk";
  return v;
}
inline size_t chunk_begin(size_t n, int chunks, int c) { return n * size_t(c) / size_t(chunks); }
inline void share_of(size_t b, size_t e, int part, int parts, size_t* sb, size_t* se) {
  const size_t m = e - b;
  *sb = b + m * size_t(part) / size_t(parts);
  *se = b + m * size_t(part + 1) / size_t(parts);
}
inline void spin_until(const std::atomic<int>& a, int at_least) {
  for (int spins = 0; a.load(std::memory_order_acquire) < at_least; ++spins) {
    if (spins < 4096) _mm_pause();
    else std::this_thread::yield();
  }
}

// pack + enqueue the upload of positions (+ fixed flags, + velocities when d_vel != nullptr)
cudaError_t upload_packed(const Entity* state, size_t n, double4* h_pos, uint8_t* h_fixed, double4* h_vel,
                          double4* d_pos, uint8_t* d_fixed, double4* d_vel, cudaStream_t st) {
  const int chunks = chunk_count(n);
  std::atomic<int> packed[kMaxChunks];
  for (auto& a : packed) a.store(0, std::memory_order_relaxed);
  std::atomic<int> failed{0};  // part 0 hit a CUDA error: the other parts stop packing
  cudaError_t err = cudaSuccess;
  const double t_pack = now_ms();
  HostPool& pool = HostPool::instance();
  if (per_chunk_regions()) {
    for (int c = 0; c < chunks; ++c) {
      const size_t b = chunk_begin(n, chunks, c), e = chunk_begin(n, chunks, c + 1);
      if (e == b) continue;
      pool.parallel_for(e - b, kParallelGrain, [&](size_t sb, size_t se) {
        pack_positions_range(state, b + sb, b + se, h_pos, h_fixed);
        if (d_vel) pack_velocities_range(state, b + sb, b + se, h_vel);
        _mm_sfence();
      });
      PB_CUDA(cudaMemcpyAsync(d_pos + b, h_pos + b, (e - b) * sizeof(double4), cudaMemcpyHostToDevice, st));
      if (d_vel) PB_CUDA(cudaMemcpyAsync(d_vel + b, h_vel + b, (e - b) * sizeof(double4), cudaMemcpyHostToDevice, st));
    }
    g_pack_ms += now_ms() - t_pack;
    if (h_fixed && n) PB_CUDA(cudaMemcpyAsync(d_fixed, h_fixed, n, cudaMemcpyHostToDevice, st));
    return cudaSuccess;
  }
  pool.parallel_parts(pool.parts_for(n, kParallelGrain), [&](int part, int parts) {
    for (int c = 0; c < chunks; ++c) {
      const size_t b = chunk_begin(n, chunks, c), e = chunk_begin(n, chunks, c + 1);
      if (failed.load(std::memory_order_relaxed)) return;
      size_t sb, se;
      share_of(b, e, part, parts, &sb, &se);
      pack_positions_range(state, sb, se, h_pos, h_fixed);
      if (d_vel) pack_velocities_range(state, sb, se, h_vel);
      _mm_sfence();
      packed[c].fetch_add(1, std::memory_order_release);
      if (part != 0 || e == b) continue;
      spin_until(packed[c], parts);
      err = cudaMemcpyAsync(d_pos + b, h_pos + b, (e - b) * sizeof(double4), cudaMemcpyHostToDevice, st);
      if (err == cudaSuccess && d_vel)
        err = cudaMemcpyAsync(d_vel + b, h_vel + b, (e - b) * sizeof(double4), cudaMemcpyHostToDevice, st);
      if (err != cudaSuccess) {
        failed.store(1, std::memory_order_relaxed);
        return;
      }
    }
  });
  g_pack_ms += now_ms() - t_pack;
  if (err != cudaSuccess) {
    set_error("CUDA error: %s (upload_packed)", cudaGetErrorString(err));
    return err;
  }
  if (h_fixed && n) PB_CUDA(cudaMemcpyAsync(d_fixed, h_fixed, n, cudaMemcpyHostToDevice, st));
  return cudaSuccess;
}

// D2H of `bytes_per_item`-byte records in chunks, fn(begin, end) applied to each chunk behind its copy
// by all parts of the pool (see the note on chunking above)
cudaError_t download_chunked(GpuContext& gpu, cudaStream_t st, void* host, const void* dev, size_t n,
                             size_t bytes_per_item, cudaEvent_t after_copies,
                             const std::function<void(size_t, size_t)>& fn) {
  const int chunks = chunk_count(n);
  for (int c = 0; c < chunks; ++c) {
    const size_t b = chunk_begin(n, chunks, c), e = chunk_begin(n, chunks, c + 1);
    if (e > b)
      PB_CUDA(cudaMemcpyAsync(static_cast<char*>(host) + b * bytes_per_item,
                              static_cast<const char*>(dev) + b * bytes_per_item, (e - b) * bytes_per_item,
                              cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaEventRecord(gpu.chunk_ev[c], st));
  }
  if (after_copies) PB_CUDA(cudaEventRecord(after_copies, st));
  std::atomic<int> arrived{0};  // chunks whose copy has completed (-1: CUDA error)
  cudaError_t err = cudaSuccess;
  const double t_un = now_ms();
  HostPool& pool = HostPool::instance();
  if (per_chunk_regions()) {
    for (int c = 0; c < chunks; ++c) {
      const size_t b = chunk_begin(n, chunks, c), e = chunk_begin(n, chunks, c + 1);
      PB_CUDA(cudaEventSynchronize(gpu.chunk_ev[c]));
      if (e > b) pool.parallel_for(e - b, kParallelGrain, [&](size_t sb, size_t se) { fn(b + sb, b + se); });
    }
    g_unpack_ms += now_ms() - t_un;
    return cudaSuccess;
  }
  pool.parallel_parts(pool.parts_for(n, kParallelGrain), [&](int part, int parts) {
    for (int c = 0; c < chunks; ++c) {
      if (part == 0) {
        err = cudaEventSynchronize(gpu.chunk_ev[c]);
        arrived.store(err == cudaSuccess ? c + 1 : 1 << 20, std::memory_order_release);
        if (err != cudaSuccess) return;
      } else {
        spin_until(arrived, c + 1);
        if (arrived.load(std::memory_order_relaxed) >= (1 << 20)) return;
      }
      const size_t b = chunk_begin(n, chunks, c), e = chunk_begin(n, chunks, c + 1);
      size_t sb, se;
      share_of(b, e, part, parts, &sb, &se);
      if (se > sb) fn(sb, se);
    }
  });
  g_unpack_ms += now_ms() - t_un;
  if (err != cudaSuccess) {
    set_error("CUDA error: %s (download_chunked)", cudaGetErrorString(err));
    return err;
  }
  return cudaSuccess;
}

cudaError_t transform_forces(TransformObj& t, const Entity* state, size_t n, Acceleration* acc) {
  const double t_wall = now_ms();
  g_pack_ms = g_unpack_ms = 0.0;
  PB_PASS(t.gpu.init(t.device));
  cudaStream_t st = t.gpu.stream;
  PB_PASS(t.h_pos.ensure(n * sizeof(double4)));
  PB_PASS(t.h_fixed.ensure(n));
  PB_PASS(t.h_acc.ensure(n * sizeof(float4)));
  PB_PASS(t.d_pos.ensure(n * sizeof(double4)));
  PB_PASS(t.d_fixed.ensure(n));
  PB_CUDA(cudaEventRecord(t.gpu.ev[0], st));
  PB_PASS(upload_packed(state, n, t.h_pos.as<double4>(), t.h_fixed.as<uint8_t>(), nullptr,
                        t.d_pos.as<double4>(), t.d_fixed.as<uint8_t>(), nullptr, st));
  PB_CUDA(cudaEventRecord(t.gpu.ev[1], st));
  t.ws.pos64 = t.d_pos.as<double4>();
  t.ws.fixed = t.d_fixed.as<uint8_t>();
  t.ws.n = n;
  PB_PASS(gravity_evaluate(t.ws, t.prm, 0, n, st, t.ls, true));
  PB_CUDA(cudaEventRecord(t.gpu.ev[2], st));
  const float4* h = t.h_acc.as<float4>();
  // accelerations[i] += f / m_a for every non-fixed body (transformers.rs:139-141,154-158),
  // chunk by chunk behind the copies
  PB_PASS(download_chunked(t.gpu, st, t.h_acc.p, t.ws.acc.p, n, sizeof(float4), t.gpu.ev[3],
                           [&](size_t b, size_t e) {
                             for (size_t i = b; i < e; ++i) {
                               if (state[i].fixed) continue;
                               acc[i].x += double(h[i].x);
                               acc[i].y += double(h[i].y);
                               acc[i].z += double(h[i].z);
                             }
                           }));
  PB_CUDA(cudaStreamSynchronize(st));
  t.last_n = n;
  t.have_tree = t.ws.n_cells > 0;
  t.stats.n_bodies = n;
  t.stats.n_cells = t.ws.n_cells;
  t.stats.kernel_launches = t.ls.launches;
  t.stats.ms_h2d = elapsed(t.gpu.ev[0], t.gpu.ev[1]);
  t.stats.ms_build = 0.f;
  t.stats.ms_force = elapsed(t.gpu.ev[1], t.gpu.ev[2]);
  t.stats.ms_integrate = 0.f;
  t.stats.ms_d2h = elapsed(t.gpu.ev[2], t.gpu.ev[3]);
  t.stats.ms_host_pack = float(g_pack_ms);
  t.stats.ms_host_unpack = float(g_unpack_ms);
  t.stats.ms_wall = float(now_ms() - t_wall);
  return cudaSuccess;
}

struct VerletObj {
  std::mutex mu;
  GpuContext gpu;
  int device;
  int kind = PB200_VERLET;  // Pb200Integrator
  size_t n_prev = 0;  // length of the stored previous state (verlet.rs:102: len != n => first step)
  DevBuf cur, prev, vel, acc64, fixed, out6;
  DevBuf t_pos, t_vel, s_pos, s_vel;  // rk4: evaluation point and running sums
  PinnedBuf h_pos, h_vel, h_fixed, h_acc, h_out;
  std::vector<Entity> h_temp;         // rk4, generic form: the evaluation point as host entities
  // resident mode (pb200_verlet_set_resident): the state stays in HBM between calls as whole Entity records,
  // the caller's new_state buffer is page-locked once and receives them by one DMA
  bool resident = false, on_device = false;
  DevBuf ent80;
  void* reg_ptr = nullptr;
  size_t reg_bytes = 0;
  uint64_t resident_hits = 0, resident_misses = 0;
  int d2h_bytes_per_body = 48;  // of the last step through the host boundary (24: regular verlet step, positions only)
  LaunchStats ls;
  Pb200Stats stats;
  ~VerletObj() {
    if (reg_ptr) cudaHostUnregister(reg_ptr);
    if (gpu.ready) {
      cudaSetDevice(gpu.device);
      ent80.release();
      cur.release(); prev.release(); vel.release(); acc64.release(); fixed.release(); out6.release();
      t_pos.release(); t_vel.release(); s_pos.release(); s_vel.release();
      h_pos.release(); h_vel.release(); h_fixed.release(); h_acc.release(); h_out.release();
      gpu.destroy();
    }
  }
};

// new_state[i] = entities[i] with position and velocity replaced (verlet.rs:41-48 / :72-79).
// out6 holds {x,y,z,vx,vy,vz} per body.  new_state is written with non-temporal 16-byte stores when
// aligned (it is 80 MB the caller reads later; no read-for-ownership traffic).
void unpack_state_range(const Entity* entities, Entity* out, size_t b, size_t e, const double* out6) {
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  if (aligned) {
    for (size_t i = b; i < e; ++i) {
      const double* o = out6 + 6 * i;
      const double* src = reinterpret_cast<const double*>(entities + i);
      double* dst = reinterpret_cast<double*>(out + i);
      _mm_stream_pd(dst, _mm_loadu_pd(o));
      _mm_stream_pd(dst + 2, _mm_loadu_pd(o + 2));
      _mm_stream_pd(dst + 4, _mm_loadu_pd(o + 4));
      _mm_stream_pd(dst + 6, _mm_loadu_pd(src + 6));  // radius, mass
      _mm_stream_pd(dst + 8, _mm_loadu_pd(src + 8));  // id, fixed (+ padding)
    }
    _mm_sfence();
  } else {
    for (size_t i = b; i < e; ++i) {
      const double* o = out6 + 6 * i;
      Entity t = entities[i];
      t.x = o[0]; t.y = o[1]; t.z = o[2];
      t.vx = o[3]; t.vy = o[4]; t.vz = o[5];
      out[i] = t;
    }
  }
}

// the regular verlet step's form: the device returned {x,y,z} only, v = (x' - x_in) / dt (verlet.rs:68-70) is taken
// here from the input the caller handed in.  Everything of record i is read before anything of it is written, so
// `out` may be the input array itself.
void unpack_positions_range(const Entity* entities, Entity* out, size_t b, size_t e, const double* out3, double dt) {
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  for (size_t i = b; i < e; ++i) {
    const double* o = out3 + 3 * i;
    const double* src = reinterpret_cast<const double*>(entities + i);
    const double x = o[0], y = o[1], z = o[2];
    const double vx = (x - src[0]) / dt, vy = (y - src[1]) / dt, vz = (z - src[2]) / dt;
    const __m128d tail0 = _mm_loadu_pd(src + 6), tail1 = _mm_loadu_pd(src + 8);  // radius, mass | id, fixed (+ padding)
    double* dst = reinterpret_cast<double*>(out + i);
    if (aligned) {
      _mm_stream_pd(dst, _mm_set_pd(y, x));
      _mm_stream_pd(dst + 2, _mm_set_pd(vx, z));
      _mm_stream_pd(dst + 4, _mm_set_pd(vz, vy));
      _mm_stream_pd(dst + 6, tail0);
      _mm_stream_pd(dst + 8, tail1);
    } else {
      dst[0] = x; dst[1] = y; dst[2] = z; dst[3] = vx; dst[4] = vy; dst[5] = vz;
      _mm_storeu_pd(dst + 6, tail0);
      _mm_storeu_pd(dst + 8, tail1);
    }
  }
  if (aligned) _mm_sfence();
}

cudaError_t verlet_buffers(VerletObj& v, size_t n) {
  PB_PASS(v.gpu.init(v.device));
  PB_PASS(v.cur.ensure(n * sizeof(double4)));
  PB_PASS(v.prev.ensure(n * sizeof(double4)));
  PB_PASS(v.vel.ensure(n * sizeof(double4)));
  PB_PASS(v.fixed.ensure(n));
  PB_PASS(v.h_pos.ensure(n * sizeof(double4)));
  PB_PASS(v.h_vel.ensure(n * sizeof(double4)));
  PB_PASS(v.h_fixed.ensure(n));
  PB_PASS(v.out6.ensure(n * 48));
  PB_PASS(v.h_out.ensure(n * 48));
  return cudaSuccess;
}

// shared tail of both verlet entry points: chunked D2H of the packed result, unpack behind the copies
// (first-step formula: {x,y,z,vx,vy,vz} come back, 48 B/body; regular step: {x,y,z}, 24 B/body)
cudaError_t verlet_finish(VerletObj& v, cudaStream_t st, const Entity* entities, Entity* new_state, size_t n,
                          bool first = true, double dt = 0.0) {
  const double* h = v.h_out.as<double>();
  v.d2h_bytes_per_body = first ? 48 : 24;
  if (first) {
    PB_PASS(download_chunked(v.gpu, st, v.h_out.p, v.out6.p, n, 48, v.gpu.ev[4], [&](size_t b, size_t e) {
      unpack_state_range(entities, new_state, b, e, h);
    }));
  } else {
    PB_PASS(download_chunked(v.gpu, st, v.h_out.p, v.out6.p, n, 24, v.gpu.ev[4], [&](size_t b, size_t e) {
      unpack_positions_range(entities, new_state, b, e, h, dt);
    }));
  }
  PB_CUDA(cudaStreamSynchronize(st));
  v.n_prev = n;
  v.stats.n_bodies = n;
  v.stats.kernel_launches = v.ls.launches;
  return cudaSuccess;
}

cudaError_t rk4_buffers(VerletObj& v, size_t n) {
  PB_PASS(v.t_pos.ensure(n * sizeof(double4)));
  PB_PASS(v.t_vel.ensure(n * sizeof(double4)));
  PB_PASS(v.s_pos.ensure(n * sizeof(double4)));
  PB_PASS(v.s_vel.ensure(n * sizeof(double4)));
  return cudaSuccess;
}

// rk4 with a host acceleration callback (integrators/src/rk4.rs:23-183): four evaluations, the
// evaluation points 2-4 are materialised as host entities for the callback
cudaError_t rk4_step_generic(VerletObj& v, const Entity* entities, Entity* new_state, size_t n,
                             Pb200AccFn acc_fn, void* ctx, double dt) {
  if (n == 0) return cudaSuccess;
  PB_PASS(verlet_buffers(v, n));
  PB_PASS(rk4_buffers(v, n));
  PB_PASS(v.acc64.ensure(n * sizeof(Acceleration)));
  PB_PASS(v.h_acc.ensure(n * sizeof(Acceleration)));
  cudaStream_t st = v.gpu.stream;
  std::vector<Acceleration> acc(n);
  v.h_temp.resize(n);
  PB_PASS(upload_packed(entities, n, v.h_pos.as<double4>(), v.h_fixed.as<uint8_t>(), v.h_vel.as<double4>(),
                        v.cur.as<double4>(), v.fixed.as<uint8_t>(), v.vel.as<double4>(), st));
  const Entity* at = entities;
  for (int stage = 1; stage <= 4; ++stage) {
    std::fill(acc.begin(), acc.end(), Acceleration{0.0, 0.0, 0.0});
    acc_fn(ctx, at, n, acc.data());
    std::memcpy(v.h_acc.p, acc.data(), n * sizeof(Acceleration));
    PB_CUDA(cudaMemcpyAsync(v.acc64.p, v.h_acc.p, n * sizeof(Acceleration), cudaMemcpyHostToDevice, st));
    PB_PASS(rk4_stage(stage, v.cur.as<double4>(), v.vel.as<double4>(), v.fixed.as<uint8_t>(),
                      v.t_pos.as<double4>(), v.t_vel.as<double4>(), v.s_pos.as<double4>(),
                      v.s_vel.as<double4>(), nullptr, v.acc64.as<double>(), n, dt, v.cur.as<double4>(),
                      v.vel.as<double4>(), v.out6.as<double>(), st, v.ls));
    // stages 1-3: out6 = next evaluation point -> host entities; stage 4: out6 = new state
    PB_PASS(verlet_finish(v, st, entities, stage < 4 ? v.h_temp.data() : new_state, n));
    at = v.h_temp.data();
  }
  return cudaSuccess;
}

cudaError_t verlet_step_generic(VerletObj& v, const Entity* entities, Entity* new_state, size_t n,
                                Pb200AccFn acc_fn, void* ctx, double dt) {
  if (v.kind == PB200_RK4) return rk4_step_generic(v, entities, new_state, n, acc_fn, ctx, dt);
  if (n == 0) {
    acc_fn(ctx, entities, 0, nullptr);  // verlet.rs:93-94 (an empty vector)
    v.n_prev = 0;
    return cudaSuccess;
  }
  PB_PASS(verlet_buffers(v, n));
  PB_PASS(v.acc64.ensure(n * sizeof(Acceleration)));
  PB_PASS(v.h_acc.ensure(n * sizeof(Acceleration)));
  // verlet.rs:93-94 `vec![Acceleration::zero(); n]` + acc_fn: the vector is this handle's page-locked upload
  // buffer, zeroed by the pool (a fresh 24 MB std::vector per step costs milliseconds of page faults, and the
  // accelerations would have to be copied once more before they can be uploaded)
  Acceleration* acc = v.h_acc.as<Acceleration>();
  HostPool::instance().parallel_for(n, kParallelGrain, [&](size_t b, size_t e) {
    std::memset(static_cast<void*>(acc + b), 0, (e - b) * sizeof(Acceleration));
  });
  acc_fn(ctx, entities, n, acc);
  cudaStream_t st = v.gpu.stream;
  // euler (euler.rs:30-37) is verlet's first-step formula on every step
  const bool first = v.kind == PB200_EULER || v.n_prev != n;
  PB_CUDA(cudaEventRecord(v.gpu.ev[0], st));
  PB_PASS(upload_packed(entities, n, v.h_pos.as<double4>(), nullptr, v.h_vel.as<double4>(),
                        v.cur.as<double4>(), nullptr, first ? v.vel.as<double4>() : nullptr, st));
  PB_CUDA(cudaMemcpyAsync(v.acc64.p, v.h_acc.p, n * sizeof(Acceleration), cudaMemcpyHostToDevice, st));
  PB_CUDA(cudaEventRecord(v.gpu.ev[1], st));
  PB_PASS(verlet_update(v.cur.as<double4>(), v.prev.as<double4>(), v.vel.as<double4>(), nullptr,
                        v.acc64.as<double>(), n, dt, first ? 1 : 0, st, v.ls, v.out6.as<double>()));
  PB_CUDA(cudaEventRecord(v.gpu.ev[3], st));
  PB_PASS(verlet_finish(v, st, entities, new_state, n, first, dt));
  v.stats.ms_h2d = elapsed(v.gpu.ev[0], v.gpu.ev[1]);
  v.stats.ms_force = 0.f;
  v.stats.ms_integrate = elapsed(v.gpu.ev[1], v.gpu.ev[3]);
  v.stats.ms_d2h = elapsed(v.gpu.ev[3], v.gpu.ev[4]);
  return cudaSuccess;
}

// rk4 with the accelerations of `t` evaluated on the device at every stage
cudaError_t rk4_step_fused(VerletObj& v, TransformObj& t, const Entity* entities, Entity* new_state, size_t n,
                           double dt) {
  v.device = t.device;
  PB_PASS(verlet_buffers(v, n));
  PB_PASS(rk4_buffers(v, n));
  PB_PASS(t.gpu.init(t.device));
  cudaStream_t st = v.gpu.stream;
  PB_PASS(upload_packed(entities, n, v.h_pos.as<double4>(), v.h_fixed.as<uint8_t>(), v.h_vel.as<double4>(),
                        v.cur.as<double4>(), v.fixed.as<uint8_t>(), v.vel.as<double4>(), st));
  t.ws.fixed = v.fixed.as<uint8_t>();
  t.ws.n = n;
  for (int stage = 1; stage <= 4; ++stage) {
    t.ws.pos64 = stage == 1 ? v.cur.as<double4>() : v.t_pos.as<double4>();
    PB_PASS(gravity_evaluate(t.ws, t.prm, 0, n, st, t.ls, true));
    PB_PASS(rk4_stage(stage, v.cur.as<double4>(), v.vel.as<double4>(), v.fixed.as<uint8_t>(),
                      v.t_pos.as<double4>(), v.t_vel.as<double4>(), v.s_pos.as<double4>(),
                      v.s_vel.as<double4>(), t.ws.acc.as<float4>(), nullptr, n, dt, v.cur.as<double4>(),
                      v.vel.as<double4>(), stage == 4 ? v.out6.as<double>() : nullptr, st, v.ls));
  }
  PB_PASS(verlet_finish(v, st, entities, new_state, n));
  t.last_n = n;
  t.stats.n_bodies = n;
  t.stats.n_cells = t.ws.n_cells;
  t.stats.kernel_launches = t.ls.launches;
  v.stats.n_cells = t.ws.n_cells;
  v.stats.kernel_launches = v.ls.launches + t.ls.launches;
  return cudaSuccess;
}

// Fused step, resident form (opt-in).  physim's loop hands the integrator `state` = a clone of the new_state it
// returned one step earlier (pipeline.rs:166-173): when that is still true - checked on a sample of entities
// against what this handle wrote into new_state last time - the device already holds the input and the
// host->device copy is skipped.  The result leaves as whole 80-byte Entity records by ONE copy into the caller's
// new_state, page-locked on first use (it is a persistent Vec, pipeline.rs:100-103).  A different n, another
// buffer or a wholesale edit of the state is detected (pointer / size check, sample) and costs one full upload; an
// edit of a few bodies between two calls (a transmute element) can escape the sample - the opt-in's contract
// excludes it (include/physim_b200.h).
cudaError_t verlet_step_fused_resident(VerletObj& v, TransformObj& t, const Entity* entities, Entity* new_state, size_t n,
                                       double dt) {
  const double t_wall = now_ms();
  g_pack_ms = g_unpack_ms = 0.0;
  v.device = t.device;
  PB_PASS(verlet_buffers(v, n));
  PB_PASS(t.gpu.init(t.device));
  PB_PASS(v.ent80.ensure(n * sizeof(Entity)));
  cudaStream_t st = v.gpu.stream;
  const size_t bytes = n * sizeof(Entity);
  // is the device's state what the caller passes?  (the output buffer still holds the previous step's result)
  bool have = v.on_device && v.n_prev == n && v.reg_ptr == new_state && v.reg_bytes == bytes && v.kind != PB200_EULER;
  if (have) {
    const size_t samples = std::min<size_t>(n, 2048);
    const size_t stride = n / samples;
    for (size_t k = 0; k < samples && have; ++k) {
      const size_t i = k * stride + (k * 7919u) % stride;
      have = std::memcmp(&entities[i], &new_state[i], offsetof(Entity, fixed) + 1) == 0;
    }
  }
  if (v.reg_ptr != new_state || v.reg_bytes != bytes) {
    if (v.reg_ptr) cudaHostUnregister(v.reg_ptr);
    v.reg_ptr = nullptr;
    v.reg_bytes = 0;
    if (cudaHostRegister(new_state, bytes, cudaHostRegisterDefault) == cudaSuccess) {
      v.reg_ptr = new_state;
      v.reg_bytes = bytes;
    } else {
      cudaGetLastError();  // not fatal: the copy below is then staged by the driver
    }
  }
  PB_CUDA(cudaEventRecord(v.gpu.ev[0], st));
  bool first = v.kind == PB200_EULER || v.n_prev != n;
  if (!have) {
    v.resident_misses += 1;
    PB_CUDA(cudaMemcpyAsync(v.ent80.p, entities, bytes, cudaMemcpyHostToDevice, st));
    PB_PASS(entity_split(v.ent80.p, n, v.cur.as<double4>(), v.vel.as<double4>(), v.fixed.as<uint8_t>(), st, v.ls));
    // (same n: the regular step still uses the previous call's input as x_{n-1}, whatever this input is - exactly
    // what verlet.rs:52-82 does with its stored previous_state)
  } else {
    v.resident_hits += 1;
  }
  PB_CUDA(cudaEventRecord(v.gpu.ev[1], st));
  t.ws.pos64 = v.cur.as<double4>();
  t.ws.fixed = v.fixed.as<uint8_t>();
  t.ws.n = n;
  PB_PASS(gravity_evaluate(t.ws, t.prm, 0, n, st, t.ls, true));
  PB_CUDA(cudaEventRecord(v.gpu.ev[2], st));
  PB_PASS(verlet_update(v.cur.as<double4>(), v.prev.as<double4>(), v.vel.as<double4>(), t.ws.acc.as<float4>(), nullptr, n,
                        dt, first ? 1 : 0, st, v.ls));
  PB_PASS(entity_merge(v.ent80.p, n, v.cur.as<double4>(), v.vel.as<double4>(), st, v.ls));
  PB_CUDA(cudaEventRecord(v.gpu.ev[3], st));
  PB_CUDA(cudaMemcpyAsync(new_state, v.ent80.p, bytes, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaEventRecord(v.gpu.ev[4], st));
  PB_CUDA(cudaStreamSynchronize(st));
  v.on_device = true;
  v.n_prev = n;
  t.last_n = n;
  t.stats.n_bodies = n;
  t.stats.n_cells = t.ws.n_cells;
  t.stats.kernel_launches = t.ls.launches;
  v.stats.n_bodies = n;
  v.stats.n_cells = t.ws.n_cells;
  v.stats.kernel_launches = v.ls.launches + t.ls.launches;
  v.stats.ms_h2d = elapsed(v.gpu.ev[0], v.gpu.ev[1]);
  v.stats.ms_force = elapsed(v.gpu.ev[1], v.gpu.ev[2]);
  v.stats.ms_integrate = elapsed(v.gpu.ev[2], v.gpu.ev[3]);
  v.stats.ms_d2h = elapsed(v.gpu.ev[3], v.gpu.ev[4]);
  v.stats.ms_host_pack = 0.f;
  v.stats.ms_host_unpack = 0.f;
  v.stats.ms_wall = float(now_ms() - t_wall);
  return cudaSuccess;
}

cudaError_t verlet_step_fused(VerletObj& v, TransformObj& t, const Entity* entities, Entity* new_state, size_t n,
                              double dt) {
  if (n == 0) {
    v.n_prev = 0;
    v.on_device = false;
    return cudaSuccess;
  }
  if (v.kind == PB200_RK4) return rk4_step_fused(v, t, entities, new_state, n, dt);
  if (v.resident) return verlet_step_fused_resident(v, t, entities, new_state, n, dt);
  v.on_device = false;
  const double t_wall = now_ms();
  g_pack_ms = g_unpack_ms = 0.0;
  v.device = t.device;
  PB_PASS(verlet_buffers(v, n));
  PB_PASS(t.gpu.init(t.device));
  cudaStream_t st = v.gpu.stream;
  const bool first = v.kind == PB200_EULER || v.n_prev != n;
  PB_CUDA(cudaEventRecord(v.gpu.ev[0], st));
  PB_PASS(upload_packed(entities, n, v.h_pos.as<double4>(), v.h_fixed.as<uint8_t>(), v.h_vel.as<double4>(),
                        v.cur.as<double4>(), v.fixed.as<uint8_t>(), first ? v.vel.as<double4>() : nullptr, st));
  PB_CUDA(cudaEventRecord(v.gpu.ev[1], st));
  t.ws.pos64 = v.cur.as<double4>();
  t.ws.fixed = v.fixed.as<uint8_t>();
  t.ws.n = n;
  PB_PASS(gravity_evaluate(t.ws, t.prm, 0, n, st, t.ls, true));
  PB_CUDA(cudaEventRecord(v.gpu.ev[2], st));
  PB_PASS(verlet_update(v.cur.as<double4>(), v.prev.as<double4>(), v.vel.as<double4>(),
                        t.ws.acc.as<float4>(), nullptr, n, dt, first ? 1 : 0, st, v.ls, v.out6.as<double>()));
  PB_CUDA(cudaEventRecord(v.gpu.ev[3], st));
  PB_PASS(verlet_finish(v, st, entities, new_state, n, first, dt));
  t.last_n = n;
  t.stats.n_bodies = n;
  t.stats.n_cells = t.ws.n_cells;
  t.stats.kernel_launches = t.ls.launches;
  v.stats.n_cells = t.ws.n_cells;
  v.stats.kernel_launches = v.ls.launches + t.ls.launches;
  v.stats.ms_h2d = elapsed(v.gpu.ev[0], v.gpu.ev[1]);
  v.stats.ms_force = elapsed(v.gpu.ev[1], v.gpu.ev[2]);
  v.stats.ms_integrate = elapsed(v.gpu.ev[2], v.gpu.ev[3]);
  v.stats.ms_d2h = elapsed(v.gpu.ev[3], v.gpu.ev[4]);
  v.stats.ms_host_pack = float(g_pack_ms);
  v.stats.ms_host_unpack = float(g_unpack_ms);
  v.stats.ms_wall = float(now_ms() - t_wall);
  return cudaSuccess;
}

struct SimObj {
  GravityParams prm;
  double dt;
  int rank, world, device;
  std::mutex mu;
  GpuContext gpu;
  GravityWorkspace ws;
  DevBuf cur, prev, vel, fixed;  // cur doubles as the all-gather buffer ({x,y,z,m} of all n bodies)
  DevBuf ck_cur, ck_prev, ck_vel;  // checkpoint of the last verified state
  uint64_t replays = 0;
  PinnedBuf h_pos, h_vel, h_fixed, h_acc;
  size_t n = 0, t0 = 0, t1 = 0, slice = 0;
  int integrator = PB200_VERLET;
  DevBuf t_pos, t_vel, s_pos, s_vel;  // rk4
  bool first = true, checked = false;
  // lean verlet steps (world == 1, all bodies integrated): cur / prev swap roles every step, `vel` is
  // derived on demand (vel_stale), and the step leaves the extent of its new positions in ext[ext_slot]
  // for the next tree build (ext_ready); ext_dirty: the slots must be zeroed before the next lean step
  bool vel_stale = false, ext_ready = false, ext_dirty = true;
  bool pin_cur = false;  // pb200_sim_gather_buffer handed out cur's address
  bool ext_last_valid = false;  // ext[2] holds the extent the last build consumed (it came from a lean step's slot)
  int ext_slot = 0;
  DevBuf ext;
  LaunchStats ls;
  Pb200Stats stats;
  ~SimObj() {
    if (gpu.ready) {
      cudaSetDevice(gpu.device);
      ws.release_all();
      cur.release(); prev.release(); vel.release(); fixed.release(); ext.release();
      ck_cur.release(); ck_prev.release(); ck_vel.release();
      t_pos.release(); t_vel.release(); s_pos.release(); s_vel.release();
      h_pos.release(); h_vel.release(); h_fixed.release(); h_acc.release();
      gpu.destroy();
    }
  }
};

cudaError_t sim_upload(SimObj& s, const Entity* state, size_t n) {
  PB_PASS(s.gpu.init(s.device));
  cudaStream_t st = s.gpu.stream;
  s.n = n;
  // equal slices of S = ceil(n / world) bodies (the last rank's is shorter); the gather buffer holds
  // world * S records so that one in-place all-gather with equal counts serves every n
  s.slice = (n + size_t(s.world) - 1) / size_t(s.world);
  s.t0 = std::min(n, s.slice * size_t(s.rank));
  s.t1 = std::min(n, s.slice * size_t(s.rank + 1));
  s.first = true;
  s.checked = false;
  s.vel_stale = false;
  s.ext_ready = false;
  s.ext_dirty = true;
  s.ext_last_valid = false;
  if (n == 0) return cudaSuccess;
  PB_PASS(s.cur.ensure(s.slice * size_t(s.world) * sizeof(double4)));
  PB_PASS(s.prev.ensure(n * sizeof(double4)));
  PB_PASS(s.vel.ensure(n * sizeof(double4)));
  PB_PASS(s.fixed.ensure(n));
  PB_PASS(s.h_pos.ensure(n * sizeof(double4)));
  PB_PASS(s.h_vel.ensure(n * sizeof(double4)));
  PB_PASS(s.h_fixed.ensure(n));
  PB_PASS(upload_packed(state, n, s.h_pos.as<double4>(), s.h_fixed.as<uint8_t>(), s.h_vel.as<double4>(),
                        s.cur.as<double4>(), s.fixed.as<uint8_t>(), s.vel.as<double4>(), st));
  PB_CUDA(cudaStreamSynchronize(st));
  s.ws.pos64 = s.cur.as<double4>();
  s.ws.fixed = s.fixed.as<uint8_t>();
  s.ws.n = n;
  s.ws.n_cells = 0;
  return cudaSuccess;
}

// the same handle state as sim_upload, with the bodies made on the device (generate.cu)
cudaError_t sim_generate_cube(SimObj& s, size_t n, uint64_t seed, double spin, double mass, double size,
                              const double centre[3]) {
  PB_PASS(s.gpu.init(s.device));
  cudaStream_t st = s.gpu.stream;
  s.n = n;
  s.slice = (n + size_t(s.world) - 1) / size_t(s.world);
  s.t0 = std::min(n, s.slice * size_t(s.rank));
  s.t1 = std::min(n, s.slice * size_t(s.rank + 1));
  s.first = true;
  s.checked = false;
  s.vel_stale = false;
  s.ext_ready = false;
  s.ext_dirty = true;
  s.ext_last_valid = false;
  if (n == 0) return cudaSuccess;
  PB_PASS(s.cur.ensure(s.slice * size_t(s.world) * sizeof(double4)));
  PB_PASS(s.prev.ensure(n * sizeof(double4)));
  PB_PASS(s.vel.ensure(n * sizeof(double4)));
  PB_PASS(s.fixed.ensure(n));
  PB_PASS(s.h_pos.ensure(n * sizeof(double4)));  // (read-out staging)
  PB_PASS(s.h_vel.ensure(n * sizeof(double4)));
  PB_PASS(generate_cube(s.cur.as<double4>(), s.vel.as<double4>(), s.fixed.as<uint8_t>(), n, seed, spin, mass, size,
                        centre, st, s.ls));
  PB_CUDA(cudaStreamSynchronize(st));
  s.ws.pos64 = s.cur.as<double4>();
  s.ws.fixed = s.fixed.as<uint8_t>();
  s.ws.n = n;
  s.ws.n_cells = 0;
  return cudaSuccess;
}

// velocities of the current state, if the lean steps left them implicit
cudaError_t sim_materialise_velocities(SimObj& s) {
  if (!s.vel_stale) return cudaSuccess;
  PB_PASS(verlet_velocity(s.cur.as<double4>(), s.prev.as<double4>(), s.vel.as<double4>(), s.n, s.dt,
                          s.gpu.stream, s.ls));
  s.vel_stale = false;
  return cudaSuccess;
}

inline bool sim_lean_enabled() {  // PB200_VERLET_LEAN=0: general verlet kernel every step (A/B runs, tests)
  const char* e = std::getenv("PB200_VERLET_LEAN");  // (read per step: tests flip it inside one process)
  return !(e && std::atoi(e) == 0);
}

// one step: forces for targets [t0,t1) from all n positions, then verlet on the owned slice
cudaError_t sim_step(SimObj& s, bool force_check) {
  if (s.n == 0) return cudaSuccess;
  cudaStream_t st = s.gpu.stream;
  const bool check = force_check || !s.checked;  // size the cell table once, then stay asynchronous
  const bool lean = sim_lean_enabled() && s.integrator == PB200_VERLET && !s.first && s.world == 1 &&
                    !s.pin_cur && s.t0 == 0 && s.t1 == s.n;
  if (!lean) PB_PASS(sim_materialise_velocities(s));
  if (s.integrator == PB200_RK4) {
    // four force evaluations per step on the device-side evaluation points (world == 1)
    const size_t bytes = s.n * sizeof(double4);
    PB_PASS(s.t_pos.ensure(bytes));
    PB_PASS(s.t_vel.ensure(bytes));
    PB_PASS(s.s_pos.ensure(bytes));
    PB_PASS(s.s_vel.ensure(bytes));
    for (int stage = 1; stage <= 4; ++stage) {
      s.ws.pos64 = stage == 1 ? s.cur.as<double4>() : s.t_pos.as<double4>();
      PB_PASS(gravity_evaluate(s.ws, s.prm, 0, s.n, st, s.ls, check));
      PB_PASS(rk4_stage(stage, s.cur.as<double4>(), s.vel.as<double4>(), s.fixed.as<uint8_t>(),
                        s.t_pos.as<double4>(), s.t_vel.as<double4>(), s.s_pos.as<double4>(),
                        s.s_vel.as<double4>(), s.ws.acc.as<float4>(), nullptr, s.n, s.dt,
                        s.cur.as<double4>(), s.vel.as<double4>(), nullptr, st, s.ls));
    }
    s.ws.pos64 = s.cur.as<double4>();
    s.checked = true;
    s.first = false;
    s.ext_ready = false;
    s.ext_dirty = true;
    s.ext_last_valid = false;
    return cudaSuccess;
  }
  s.ws.pos64 = s.cur.as<double4>();
  s.ws.extent_pre = (lean && s.ext_ready) ? s.ext.as<unsigned long long>() + s.ext_slot : nullptr;
  PB_PASS(gravity_evaluate(s.ws, s.prm, s.t0, s.t1, st, s.ls, check));
  s.ws.extent_pre = nullptr;  // (the direct sum does not consume it)
  s.checked = true;
  if (lean) {
    PB_PASS(s.ext.ensure(32));  // two alternating slots + the extent the last build used (statistics)
    if (s.ext_dirty) {
      PB_CUDA(cudaMemsetAsync(s.ext.p, 0, 16, st));
      s.ext_dirty = false;
    }
    // the extent just consumed (if any) sits in ext_slot: reduce the new one into the other slot, and
    // let the kernel zero the consumed one for the step after
    const int out = s.ext_slot ^ 1;
    PB_PASS(verlet_update_lean(s.cur.as<double4>(), s.prev.as<double4>(), s.ws.acc.as<float4>(), s.n, s.dt,
                               s.ext.as<unsigned long long>() + out, s.ext.as<unsigned long long>() + s.ext_slot,
                               s.ext.as<unsigned long long>() + 2, st, s.ls));
    s.ext_last_valid = s.ws.extent_cur == s.ext.as<unsigned long long>() + s.ext_slot;
    std::swap(s.cur, s.prev);  // x_{n+1} was written over x_{n-1}
    s.ws.pos64 = s.cur.as<double4>();
    s.ext_slot = out;
    s.ext_ready = true;
    s.vel_stale = true;
    return cudaSuccess;
  }
  const size_t nl = s.t1 - s.t0;
  PB_PASS(verlet_update(s.cur.as<double4>() + s.t0, s.prev.as<double4>() + s.t0,
                        s.vel.as<double4>() + s.t0, s.ws.acc.as<float4>() + s.t0, nullptr, nl, s.dt,
                        (s.first || s.integrator == PB200_EULER) ? 1 : 0, st, s.ls));
  s.first = false;
  s.ext_ready = false;
  s.ext_dirty = true;
  s.ext_last_valid = false;
  return cudaSuccess;
}

bool sim_is_direct(const SimObj& s) {
  return s.prm.kind == PB200_SIMPLE_ASTRO || !(s.prm.theta > 0.0);
}

// `steps` steps back to back.  Tree builds run unchecked (no host sync) in chunks of up to 32 steps;
// each chunk starts from a device-side checkpoint and is verified at its end (cell-table capacity,
// truncated-sort validity).  A chunk that fails verification is restored and replayed with a
// host check after every build, so the result never depends on an unverified tree.
typedef void (*ExchangeFn)(void*);

cudaError_t sim_run_steps(SimObj& s, size_t steps, ExchangeFn exchange = nullptr, void* xctx = nullptr) {
  if (s.n == 0) return cudaSuccess;
  cudaStream_t st = s.gpu.stream;
  if (sim_is_direct(s)) {
    for (size_t i = 0; i < steps; ++i) {
      PB_PASS(sim_step(s, false));
      if (exchange) exchange(xctx);
    }
    return cudaSuccess;
  }
  const size_t bytes = s.n * sizeof(double4);  // every rank holds all n positions
  while (steps) {
    const size_t chunk = steps < 32 ? steps : 32;
    PB_PASS(s.ck_cur.ensure(bytes));
    PB_PASS(s.ck_prev.ensure(bytes));
    PB_PASS(s.ck_vel.ensure(bytes));
    PB_CUDA(cudaMemcpyAsync(s.ck_cur.p, s.cur.p, bytes, cudaMemcpyDeviceToDevice, st));
    PB_CUDA(cudaMemcpyAsync(s.ck_prev.p, s.prev.p, bytes, cudaMemcpyDeviceToDevice, st));
    PB_CUDA(cudaMemcpyAsync(s.ck_vel.p, s.vel.p, bytes, cudaMemcpyDeviceToDevice, st));
    const bool first_at_ck = s.first, vel_stale_at_ck = s.vel_stale;
    for (size_t i = 0; i < chunk; ++i) {
      PB_PASS(sim_step(s, false));
      if (exchange) exchange(xctx);
    }
    TreeCheck chk;
    PB_PASS(gravity_check(s.ws, st, &chk));
    if (!chk.ok()) {
      if (chk.sort_error) {
        set_error("radix sort look-back did not complete");
        return cudaErrorUnknown;
      }
      PB_CUDA(cudaMemcpyAsync(s.cur.p, s.ck_cur.p, bytes, cudaMemcpyDeviceToDevice, st));
      PB_CUDA(cudaMemcpyAsync(s.prev.p, s.ck_prev.p, bytes, cudaMemcpyDeviceToDevice, st));
      PB_CUDA(cudaMemcpyAsync(s.vel.p, s.ck_vel.p, bytes, cudaMemcpyDeviceToDevice, st));
      s.first = first_at_ck;
      s.vel_stale = vel_stale_at_ck;
      s.ext_ready = false;  // (the extent slots belong to the abandoned steps)
      s.ext_dirty = true;
      s.ws.pos64 = s.cur.as<double4>();
      s.replays += 1;
      for (size_t i = 0; i < chunk; ++i) {
        PB_PASS(sim_step(s, true));
        if (exchange) exchange(xctx);
      }
    }
    steps -= chunk;
  }
  return cudaSuccess;
}

// ---- tiny flat-JSON reader for the plugin's init(): {"theta": 1.5, "e": 0.5, ...} --------------
struct JsonProps {
  bool ok = true;
  bool has_theta = false, has_e = false;
  double theta = 0.0, e = 0.0;
};

struct JsonCursor {
  const char* p;
  const char* end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  bool string(std::string* out) {
    ws();
    if (p >= end || *p != '"') return false;
    ++p;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        ++p;
        if (p >= end) return false;
        if (*p == 'u') {
          if (end - p < 5) return false;
          p += 4;
          if (out) out->push_back('?');
        } else if (out) {
          out->push_back(*p);
        }
        ++p;
      } else {
        if (out) out->push_back(*p);
        ++p;
      }
    }
    if (p >= end) return false;
    ++p;
    return true;
  }
  // skips any JSON value; *num/*is_num report a number (serde's as_f64 accepts ints and floats)
  bool value(double* num, bool* is_num, int depth = 0) {
    ws();
    *is_num = false;
    if (p >= end || depth > 32) return false;
    const char c = *p;
    if (c == '"') return string(nullptr);
    if (c == '{' || c == '[') {
      const char close = (c == '{') ? '}' : ']';
      ++p;
      ws();
      if (eat(close)) return true;
      for (;;) {
        if (c == '{') {
          if (!string(nullptr) || !eat(':')) return false;
        }
        double d;
        bool b;
        if (!value(&d, &b, depth + 1)) return false;
        if (eat(',')) continue;
        return eat(close);
      }
    }
    if (end - p >= 4 && !std::strncmp(p, "true", 4)) { p += 4; return true; }
    if (end - p >= 5 && !std::strncmp(p, "false", 5)) { p += 5; return true; }
    if (end - p >= 4 && !std::strncmp(p, "null", 4)) { p += 4; return true; }
    if (c == '-' || (c >= '0' && c <= '9')) {
      char buf[64];
      size_t k = 0;
      while (p < end && k + 1 < sizeof buf &&
             ((*p >= '0' && *p <= '9') || *p == '-' || *p == '+' || *p == '.' || *p == 'e' || *p == 'E'))
        buf[k++] = *p++;
      buf[k] = 0;
      char* stop = nullptr;
      *num = std::strtod(buf, &stop);
      if (stop == buf || *stop != 0) return false;
      *is_num = true;
      return true;
    }
    return false;
  }
};

JsonProps parse_props(const uint8_t* json, size_t len) {
  JsonProps out;
  JsonCursor c{reinterpret_cast<const char*>(json), reinterpret_cast<const char*>(json) + len};
  if (!json) { out.ok = false; return out; }
  if (!c.eat('{')) { out.ok = false; return out; }
  if (c.eat('}')) { c.ws(); out.ok = (c.p == c.end); return out; }
  for (;;) {
    std::string key;
    double d = 0.0;
    bool is_num = false;
    if (!c.string(&key) || !c.eat(':') || !c.value(&d, &is_num)) { out.ok = false; return out; }
    if (is_num && key == "theta") { out.has_theta = true; out.theta = d; }
    if (is_num && key == "e") { out.has_e = true; out.e = d; }
    if (c.eat(',')) continue;
    if (!c.eat('}')) { out.ok = false; return out; }
    break;
  }
  c.ws();
  out.ok = (c.p == c.end);
  return out;
}

template <class T>
cudaError_t copy_out(T* host, const void* dev, size_t count, cudaStream_t st) {
  if (!host || !count) return cudaSuccess;
  PB_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
  return cudaSuccess;
}

}  // namespace
}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pb200_set_device(int device) {
  g_device = device;
  return 0;
}

const char* pb200_last_error(void) { return g_error; }

void* pb200_transform_create(int kind, double theta, double e) {
  if (kind < PB200_ASTRO || kind > PB200_SIMPLE_ASTRO) {
    set_error("unknown element kind %d", kind);
    return nullptr;
  }
  auto* t = new TransformObj();
  t->prm.kind = kind;
  t->prm.theta = std::isnan(theta) ? 1.0 : theta;        // transformers.rs:72-75
  t->prm.easing = std::isnan(e) ? 1.0 : std::fabs(e);    // transformers.rs:77-81 (.abs())
  t->device = g_device;
  std::memset(&t->stats, 0, sizeof t->stats);
  return t;
}

void* pb200_transform_create_json(int kind, const uint8_t* json, size_t len) {
  const JsonProps p = parse_props(json, len);
  if (!p.ok) {
    set_error("malformed property JSON");
    return nullptr;
  }
  return pb200_transform_create(kind, p.has_theta ? p.theta : NAN, p.has_e ? p.e : NAN);
}

void pb200_transform_destroy(void* obj) { delete static_cast<TransformObj*>(obj); }

double pb200_transform_theta(void* obj) { return static_cast<TransformObj*>(obj)->prm.theta; }
double pb200_transform_easing(void* obj) { return static_cast<TransformObj*>(obj)->prm.easing; }

int pb200_transform_apply(void* obj, const Entity* state, size_t n, Acceleration* acc, size_t n_acc) {
  if (!obj) {
    set_error("null transform");
    return -1;
  }
  auto& t = *static_cast<TransformObj*>(obj);
  std::lock_guard<std::mutex> lk(t.mu);
  if (n_acc < n) n = n_acc;
  if (n == 0) return 0;
  if (!state || !acc) {
    set_error("null state/acceleration pointer");
    return -1;
  }
  if (transform_forces(t, state, n, acc) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] transform failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_transform_stats(void* obj, Pb200Stats* out) {
  if (!obj || !out) return -1;
  auto& t = *static_cast<TransformObj*>(obj);
  std::lock_guard<std::mutex> lk(t.mu);
  if (t.gpu.ready && t.last_n) {
    cudaSetDevice(t.gpu.device);
    uint64_t inter = 0;
    if (gravity_count_interactions(t.ws, t.gpu.stream, t.ls, &inter) != cudaSuccess) return -1;
    t.stats.interactions = inter;
    if (t.ws.extent_bits.p && t.have_tree) {
      unsigned long long bits = 0;
      cudaMemcpy(&bits, t.ws.extent_bits.p, 8, cudaMemcpyDeviceToHost);
      std::memcpy(&t.stats.extent, &bits, 8);
    }
    t.stats.kernel_launches = t.ls.launches;
  }
  t.stats.sort_mode = uint32_t(t.ws.last_mode);
  t.stats.max_bucket = t.ws.last_max_bucket;
  *out = t.stats;
  return 0;
}

int pb200_transform_debug_tree(void* obj, uint64_t* key, uint32_t* perm, uint32_t* cell_start,
                               uint8_t* level, uint32_t* head, uint32_t* count, uint32_t* skip,
                               uint32_t* parent, double* centre_ext, double* com_mass,
                               uint32_t* counts) {
  if (!obj) return -1;
  auto& t = *static_cast<TransformObj*>(obj);
  std::lock_guard<std::mutex> lk(t.mu);
  if (!t.gpu.ready || !t.last_n) {
    set_error("no evaluation to inspect");
    return -1;
  }
  cudaSetDevice(t.gpu.device);
  cudaStream_t st = t.gpu.stream;
  const size_t n = t.last_n, c = t.ws.n_cells;
  auto run = [&]() -> cudaError_t {
    if (t.have_tree) {
      PB_PASS(gravity_fill_parents(t.ws, st, t.ls));
      PB_PASS(copy_out(key, t.ws.sorted_key, n, st));
      PB_PASS(copy_out(perm, t.ws.perm, n, st));
      PB_PASS(copy_out(cell_start, t.ws.cell_start.p, n + 1, st));
      PB_PASS(copy_out(level, t.ws.c_level.p, c, st));
      PB_PASS(copy_out(head, t.ws.c_head.p, c, st));
      PB_PASS(copy_out(count, t.ws.c_count.p, c, st));
      PB_PASS(copy_out(skip, t.ws.c_skip.p, c, st));
      PB_PASS(copy_out(parent, t.ws.c_parent.p, c, st));
      PB_PASS(copy_out(centre_ext, t.ws.c_centre_ext.p, c * 4, st));
      PB_PASS(copy_out(com_mass, t.ws.c_com.p, c * 4, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    if (t.have_tree && centre_ext && level) {
      // the device keeps the walk's link (skip | level) in the fourth word of a centre record; the inspected table
      // carries the half-width there: extent / 2^level, the same double the build's halving reaches
      unsigned long long bits = 0;
      PB_CUDA(cudaMemcpy(&bits, t.ws.extent_bits.p, 8, cudaMemcpyDeviceToHost));
      double ext = 0.0;
      std::memcpy(&ext, &bits, 8);
      for (size_t i = 0; i < c; ++i) centre_ext[4 * i + 3] = std::ldexp(ext, -int(level[i]));
    }
    if (counts) {
      std::vector<float4> a(n);
      PB_CUDA(cudaMemcpy(a.data(), t.ws.acc.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n; ++i) std::memcpy(&counts[i], &a[i].w, 4);
    }
    return cudaSuccess;
  };
  return run() == cudaSuccess ? 0 : -1;
}

int pb200_transform_debug_hint(void* obj, int sort_lo, size_t n_cells_hint) {
  if (!obj) return -1;
  auto& t = *static_cast<TransformObj*>(obj);
  std::lock_guard<std::mutex> lk(t.mu);
  t.ws.sort_lo = sort_lo;
  t.ws.tree_dim = t.prm.kind == PB200_ASTRO ? 2 : 3;
  t.ws.n_cells = n_cells_hint;
  return 0;
}

int pb200_transform_debug_sort_mode(void* obj, int mode) {
  if (!obj || mode < 0 || mode > 3) return -1;
  auto& t = *static_cast<TransformObj*>(obj);
  std::lock_guard<std::mutex> lk(t.mu);
  t.ws.sort_mode = mode;
  t.ws.tree_dim = t.prm.kind == PB200_ASTRO ? 2 : 3;
  return 0;
}

// ---- verlet -----------------------------------------------------------------------------------

void* pb200_integrator_create(int kind) {
  if (kind < PB200_VERLET || kind > PB200_RK4) {
    set_error("unknown integrator kind %d", kind);
    return nullptr;
  }
  auto* v = new VerletObj();
  v->kind = kind;
  v->device = g_device;
  std::memset(&v->stats, 0, sizeof v->stats);
  return v;
}
void* pb200_verlet_create(void) { return pb200_integrator_create(PB200_VERLET); }
void pb200_integrator_destroy(void* v) { delete static_cast<VerletObj*>(v); }
int pb200_integrator_step(void* g, const Entity* entities, Entity* new_state, size_t n, Pb200AccFn acc_fn,
                          void* ctx, double dt) {
  return pb200_verlet_step(g, entities, new_state, n, acc_fn, ctx, dt);
}
int pb200_integrator_step_fused(void* g, void* transform, const Entity* entities, Entity* new_state, size_t n,
                                double dt) {
  return pb200_verlet_step_fused(g, transform, entities, new_state, n, dt);
}
void pb200_verlet_destroy(void* v) { delete static_cast<VerletObj*>(v); }

int pb200_verlet_step(void* vp, const Entity* entities, Entity* new_state, size_t n, Pb200AccFn acc_fn,
                      void* ctx, double dt) {
  if (!vp || !acc_fn || (n && (!entities || !new_state))) {
    set_error("null argument");
    return -1;
  }
  auto& v = *static_cast<VerletObj*>(vp);
  std::lock_guard<std::mutex> lk(v.mu);
  if (verlet_step_generic(v, entities, new_state, n, acc_fn, ctx, dt) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] verlet step failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_verlet_step_fused(void* vp, void* transform, const Entity* entities, Entity* new_state, size_t n,
                            double dt) {
  if (!vp || !transform || (n && (!entities || !new_state))) {
    set_error("null argument");
    return -1;
  }
  auto& v = *static_cast<VerletObj*>(vp);
  auto& t = *static_cast<TransformObj*>(transform);
  std::lock_guard<std::mutex> lk(v.mu);
  std::lock_guard<std::mutex> lk2(t.mu);
  if (verlet_step_fused(v, t, entities, new_state, n, dt) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] fused verlet step failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_verlet_set_resident(void* vp, int on) {
  if (!vp) return -1;
  auto& v = *static_cast<VerletObj*>(vp);
  std::lock_guard<std::mutex> lk(v.mu);
  v.resident = on != 0;
  if (!v.resident) v.on_device = false;
  return 0;
}

int pb200_verlet_resident_counts(void* vp, uint64_t* hits, uint64_t* misses) {
  if (!vp) return -1;
  auto& v = *static_cast<VerletObj*>(vp);
  std::lock_guard<std::mutex> lk(v.mu);
  if (hits) *hits = v.resident_hits;
  if (misses) *misses = v.resident_misses;
  return 0;
}

/* Pb200AccFn adapter over a transform element loaded through the plugin ABI: ctx points to a Pb200TransformRef.
 * This is the composition stock physim runs - the integrator's acc_fn closure calls every transform through its
 * vtable (pipeline.rs:137-141, plugin/transform.rs:85-104) - as plain C, for hosts and benchmarks without Rust. */
void pb200_acc_from_transform(void* ctx, const Entity* state, size_t n, Acceleration* acc) {
  const auto* ref = static_cast<const Pb200TransformRef*>(ctx);
  ref->api->transform(ref->obj, state, n, acc, n);
}

int pb200_verlet_stats(void* vp, Pb200Stats* out) {
  if (!vp || !out) return -1;
  auto& v = *static_cast<VerletObj*>(vp);
  std::lock_guard<std::mutex> lk(v.mu);
  *out = v.stats;
  return 0;
}

// ---- device-resident simulation ----------------------------------------------------------------

void* pb200_sim_create(int kind, double theta, double e, double dt, int rank, int world) {
  if (kind < PB200_ASTRO || kind > PB200_SIMPLE_ASTRO || world < 1 || rank < 0 || rank >= world) {
    set_error("bad arguments to pb200_sim_create");
    return nullptr;
  }
  auto* s = new SimObj();
  s->prm.kind = kind;
  s->prm.theta = std::isnan(theta) ? 1.0 : theta;
  s->prm.easing = std::isnan(e) ? 1.0 : std::fabs(e);
  s->dt = dt;
  s->rank = rank;
  s->world = world;
  s->device = g_device;
  std::memset(&s->stats, 0, sizeof s->stats);
  return s;
}

void pb200_sim_destroy(void* sim) { delete static_cast<SimObj*>(sim); }

int pb200_sim_upload(void* sim, const Entity* state, size_t n) {
  if (!sim || (n && !state)) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (sim_upload(s, state, n) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] sim upload failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_sim_generate_cube(void* sim, size_t n, uint64_t seed, double spin, double mass, double size,
                            const double* centre3) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  const double zero[3] = {0.0, 0.0, 0.0};
  if (sim_generate_cube(s, n, seed, spin, mass, size, centre3 ? centre3 : zero) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] sim generate failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_sim_run(void* sim, size_t steps) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.gpu.ready) {
    set_error("pb200_sim_upload first");
    return -1;
  }
  cudaSetDevice(s.gpu.device);
  if (sim_run_steps(s, steps) != cudaSuccess || cudaStreamSynchronize(s.gpu.stream) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] sim run failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_sim_run_csvsink(void* sim, size_t steps, void* sink) {
  if (!sim || !sink) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.gpu.ready) {
    set_error("pb200_sim_upload first");
    return -1;
  }
  if (s.world != 1) {
    set_error("pb200_sim_run_csvsink: single-rank simulations only");
    return -1;
  }
  cudaSetDevice(s.gpu.device);
  cudaStream_t st = s.gpu.stream;
  std::vector<Entity> host;
  // what the pipeline sends its renderer (pipeline.rs:129-131,179): a copy of the state
  auto emit = [&]() -> int {
    if (!pb200_csvsink_next_is_printed(sink)) return pb200_csvsink_skip(sink);
    if (host.size() != s.n) host.assign(s.n, Entity{});
    if (cudaMemcpyAsync(s.h_pos.p, s.cur.p, s.n * sizeof(double4), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      set_error("pb200_sim_run_csvsink: copy back failed");
      return -1;
    }
    const double4* p = s.h_pos.as<double4>();
    for (size_t i = 0; i < s.n; ++i) {
      host[i].x = p[i].x; host[i].y = p[i].y; host[i].z = p[i].z;
    }
    return pb200_csvsink_push(sink, host.data(), s.n);
  };
  if (pb200_csvsink_count(sink) == 0 && emit() != 0) return -1;
  for (size_t i = 0; i < steps; ++i) {
    if (sim_run_steps(s, 1) != cudaSuccess) {
      std::fprintf(stderr, "[physim_b200] sim run failed: %s\n", g_error);
      return -1;
    }
    if (emit() != 0) return -1;
  }
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -1;
}

int pb200_sim_run_timed(void* sim, size_t steps, float* ms) {
  if (!sim || !ms) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  {
    std::lock_guard<std::mutex> lk(s.mu);
    if (!s.gpu.ready) {
      set_error("pb200_sim_upload first");
      return -1;
    }
    cudaSetDevice(s.gpu.device);
    if (cudaStreamSynchronize(s.gpu.stream) != cudaSuccess ||
        cudaEventRecord(s.gpu.ev[0], s.gpu.stream) != cudaSuccess)
      return -1;
  }
  const int rc = pb200_sim_run(sim, steps);
  std::lock_guard<std::mutex> lk(s.mu);
  if (rc != 0) return rc;
  if (cudaEventRecord(s.gpu.ev[1], s.gpu.stream) != cudaSuccess ||
      cudaEventSynchronize(s.gpu.ev[1]) != cudaSuccess)
    return -1;
  *ms = elapsed(s.gpu.ev[0], s.gpu.ev[1]);
  return 0;
}

int pb200_sim_profile(void* sim, int enable) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (s.gpu.ready) {
    cudaSetDevice(s.gpu.device);
    cudaStreamSynchronize(s.gpu.stream);
  }
  s.ls.reset_profile();
  s.ls.profiling = enable != 0;
  return 0;
}

int pb200_sim_profile_report(void* sim, char* buf, size_t cap) {
  if (!sim || !buf || cap < 3) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (s.gpu.ready) {
    cudaSetDevice(s.gpu.device);
    cudaStreamSynchronize(s.gpu.stream);
  }
  s.ls.collect();
  std::string out = "[";
  for (size_t i = 0; i < s.ls.totals.size(); ++i) {
    char row[256];
    const KernelTime& k = s.ls.totals[i];
    std::snprintf(row, sizeof row, "%s{\"kernel\":\"%s\",\"launches\":%llu,\"ms\":%.6f}", i ? "," : "",
                  k.name, static_cast<unsigned long long>(k.launches), k.ms);
    out += row;
  }
  out += "]";
  if (out.size() + 1 > cap) return -1;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int pb200_sim_step_local(void* sim) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.gpu.ready) {
    set_error("pb200_sim_upload first");
    return -1;
  }
  cudaSetDevice(s.gpu.device);
  if (sim_step(s, true) != cudaSuccess) {  // multi-rank steps are host-checked every time
    std::fprintf(stderr, "[physim_b200] sim step failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_sim_gather_buffer(void* sim, void** dev_ptr, size_t* total_bytes, size_t* slice_offset,
                            size_t* slice_bytes) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  s.pin_cur = true;  // the caller holds the address: the lean steps' buffer swap is off from now on
  if (dev_ptr) *dev_ptr = s.cur.p;
  if (total_bytes) *total_bytes = s.slice * size_t(s.world) * sizeof(double4);
  if (slice_offset) *slice_offset = s.slice * size_t(s.rank) * sizeof(double4);
  if (slice_bytes) *slice_bytes = s.slice * sizeof(double4);
  return 0;
}

int pb200_sim_download(void* sim, Entity* state, size_t n) {
  if (!sim || (n && !state)) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (n != s.n) {
    set_error("pb200_sim_download: n = %zu but the simulation holds %zu bodies", n, s.n);
    return -1;
  }
  if (n == 0) return 0;
  cudaSetDevice(s.gpu.device);
  cudaStream_t st = s.gpu.stream;
  auto run = [&]() -> cudaError_t {
    PB_PASS(sim_materialise_velocities(s));
    PB_CUDA(cudaMemcpyAsync(s.h_pos.p, s.cur.p, n * sizeof(double4), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(s.h_vel.p, s.vel.p, n * sizeof(double4), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return cudaSuccess;
  };
  if (run() != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] sim download failed: %s\n", g_error);
    return -1;
  }
  const double4* p = s.h_pos.as<double4>();
  const double4* v = s.h_vel.as<double4>();
  const size_t b0 = s.world == 1 ? 0 : s.t0, b1 = s.world == 1 ? n : s.t1;
  HostPool::instance().parallel_for(b1 - b0, kParallelGrain, [&](size_t b, size_t e) {
    for (size_t i = b0 + b; i < b0 + e; ++i) {
      state[i].x = p[i].x; state[i].y = p[i].y; state[i].z = p[i].z;
      state[i].vx = v[i].x; state[i].vy = v[i].y; state[i].vz = v[i].z;
    }
  });
  return 0;
}

int pb200_sim_last_accelerations(void* sim, Acceleration* acc, size_t n) {
  if (!sim || (n && !acc)) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (n != s.n || !s.ws.acc.p) {
    set_error("pb200_sim_last_accelerations: nothing evaluated or size mismatch");
    return -1;
  }
  cudaSetDevice(s.gpu.device);
  std::vector<float4> a(n);
  if (cudaStreamSynchronize(s.gpu.stream) != cudaSuccess ||
      cudaMemcpy(a.data(), s.ws.acc.p, n * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  for (size_t i = 0; i < n; ++i) acc[i] = Acceleration{double(a[i].x), double(a[i].y), double(a[i].z)};
  return 0;
}

int pb200_sim_stats(void* sim, Pb200Stats* out) {
  if (!sim || !out) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (s.gpu.ready && s.n && s.ws.acc.p) {
    cudaSetDevice(s.gpu.device);
    uint64_t inter = 0;
    if (gravity_count_interactions(s.ws, s.gpu.stream, s.ls, &inter) != cudaSuccess) return -1;
    s.stats.interactions = inter;
    uint32_t total = 0;
    const bool direct = s.prm.kind == PB200_SIMPLE_ASTRO || !(s.prm.theta > 0.0);
    if (!direct) {
      gravity_cell_total(s.ws, s.gpu.stream, &total);
      // the extent the last build consumed: its own reduction, or the slot the lean verlet step left
      unsigned long long bits = 0;
      const void* src = s.ext_last_valid ? static_cast<const void*>(s.ext.as<unsigned long long>() + 2) : s.ws.extent_bits.p;
      cudaMemcpy(&bits, src, 8, cudaMemcpyDeviceToHost);
      std::memcpy(&s.stats.extent, &bits, 8);
    }
    s.stats.n_cells = total;
  }
  s.stats.n_bodies = s.n;
  s.stats.kernel_launches = s.ls.launches;
  s.stats.replays = static_cast<uint32_t>(s.replays);
  s.stats.sort_mode = uint32_t(s.ws.last_mode);
  s.stats.max_bucket = s.ws.last_max_bucket;
  s.stats.sort_bits = s.ws.tree_dim ? uint32_t(s.ws.tree_dim * (s.ws.tree_dim == 3 ? 21 : 31) - s.ws.last_lo) : 0u;
  *out = s.stats;
  return 0;
}

int pb200_sim_set_integrator(void* sim, int kind) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (kind < PB200_VERLET || kind > PB200_RK4 || (kind == PB200_RK4 && s.world != 1)) {
    set_error("pb200_sim_set_integrator: bad kind (rk4 needs world == 1: every stage would need an exchange)");
    return -1;
  }
  s.integrator = kind;
  s.first = true;
  return 0;
}

int pb200_sim_set_targets(void* sim, size_t t0, size_t t1) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (t0 > t1 || t1 > s.n) {
    set_error("pb200_sim_set_targets: bad range");
    return -1;
  }
  s.t0 = t0;
  s.t1 = t1;
  return 0;
}

int pb200_sim_set_stream(void* sim, void* stream) {
  if (!sim) return -1;
  auto& s = *static_cast<SimObj*>(sim);
  std::lock_guard<std::mutex> lk(s.mu);
  if (s.gpu.init(s.device) != cudaSuccess) return -1;
  cudaStreamSynchronize(s.gpu.stream);
  if (s.gpu.own_stream && s.gpu.stream) cudaStreamDestroy(s.gpu.stream);
  s.gpu.stream = static_cast<cudaStream_t>(stream);
  s.gpu.own_stream = false;
  return 0;
}

void* pb200_sim_stream(void* sim) {
  if (!sim) return nullptr;
  return static_cast<SimObj*>(sim)->gpu.stream;
}

double pb200_probe_fp32_tflops(void) {
  double t = 0.0;
  if (cudaSetDevice(g_device) != cudaSuccess || probe_fp32(&t) != cudaSuccess) return -1.0;
  return t;
}

}  // extern "C"

// multi.cu — the device-resident simulation on several GPUs of one box.
//
// One RANK per GPU.  The ranks of a run may live in one process (the plugin's case: physim is one process;
// `pb200_msim_create` with n_local == world) or in one process each (bench.py under torchrun: n_local == 1);
// the code below is the same, a single host thread walks its local ranks phase by phase and every collective
// is issued inside an NCCL group.
//
// What is replicated and what is sharded (reference path: astro/src/transformers.rs:123-160 + verlet.rs:52-82):
//   * the integrator state (x_n, x_{n-1}: 64 B/body) is replicated and advanced identically on every rank -
//     all operations are deterministic, so the copies stay bit-identical without ever being compared;
//   * Barnes-Hut: the tree build and the walk are sharded by key range (gravity.cu, "Sharded Barnes-Hut"):
//     two collectives per step - the level-K cell records (197 KB per rank) and the accelerations in sorted
//     order with the permutation (20 B/body) - and peer-memory (NVLink) loads inside the walk for the cells
//     of other ranks that a target near a range boundary opens;
//   * direct sum (theta <= 0, simple_astro): targets sharded by body index, one all-gather of the
//     accelerations (16 B/body) per step.
// The first step of a run (verlet's first-step formula) and every step of a chunk that failed verification
// run REPLICATED (the whole single-GPU step on every rank, no communication): that is also where the cuts
// between the ranks' key ranges are (re)planned from.
//
// NCCL is resolved with dlopen at the first multi-rank use (libnccl.so.2: torch's when torch is loaded in the
// process, else the system's), so the library still links nothing but cudart and a single-GPU user never
// touches it.  Reference for the collective's role: SURVEY.md §8e "ncclAllGather ... each force evaluation".
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_pool.hpp"

namespace pb200 {
namespace {

// ---- NCCL through dlopen -------------------------------------------------------------------------
struct NcclId {
  char b[128];
};
struct NcclApi {
  void* so = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool load() {
    if (so) return true;
    const char* names[] = {std::getenv("PB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm) continue;
      so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (so) break;
    }
    if (!so) {
      set_error("cannot load NCCL (libnccl.so.2): %s", dlerror());
      return false;
    }
#define PB_SYM(field, name)                                        \
  field = reinterpret_cast<decltype(field)>(dlsym(so, name));      \
  if (!field) {                                                    \
    set_error("NCCL symbol %s missing", name);                     \
    so = nullptr;                                                  \
    return false;                                                  \
  }
    PB_SYM(GetUniqueId, "ncclGetUniqueId")
    PB_SYM(CommInitRank, "ncclCommInitRank")
    PB_SYM(CommDestroy, "ncclCommDestroy")
    PB_SYM(AllGather, "ncclAllGather")
    PB_SYM(AllReduce, "ncclAllReduce")
    PB_SYM(GroupStart, "ncclGroupStart")
    PB_SYM(GroupEnd, "ncclGroupEnd")
    PB_SYM(GetErrorString, "ncclGetErrorString")
    PB_SYM(GetVersion, "ncclGetVersion")
#undef PB_SYM
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclChar = 0, kNcclUint32 = 3, kNcclMax = 2;

#define PB_NCCL(expr)                                                                              \
  do {                                                                                             \
    int _r = (expr);                                                                               \
    if (_r != 0) {                                                                                 \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));          \
      return cudaErrorUnknown;                                                                     \
    }                                                                                              \
  } while (0)

float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// what a rank tells the others about the buffers they read or write (exchanged once per plan):
// cell table (centre_ext, com, skip), dense top tree (info, com), per-rank counts, exchange buffer, flags
constexpr int kPeerBufs = 8;
struct PeerRecord {
  uint64_t pid;
  uint64_t capacity;     // cells
  uint64_t device;
  uint64_t pad;
  uint64_t ptr[kPeerBufs];              // valid inside process `pid`
  cudaIpcMemHandle_t handle[kPeerBufs];
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
static_assert(sizeof(PeerRecord) == 32 + 8 * kPeerBufs + 64 * kPeerBufs, "PeerRecord layout");

struct RankCtx {
  int rank = 0, device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void* comm = nullptr;
  GravityWorkspace ws;
  DevBuf cur, prev, vel, fixed, ext, ck_cur, ck_prev, ck_vel, peer_stage, cnt;
  PinnedBuf h_pos, h_vel, h_fixed;
  LaunchStats ls;
  bool vel_stale = false, ext_ready = false, ext_dirty = true;
  int ext_slot = 0;
  PeerRecord opened[8];                  // what the currently mapped peer pointers were opened from
  void* mapped[8][kPeerBufs] = {};       // cudaIpcOpenMemHandle results to close
  bool have_peers = false;
};

struct MultiSim {
  GravityParams prm;
  double dt = 0.0;
  int world = 1;
  std::vector<RankCtx*> local;
  std::mutex mu;
  size_t n = 0, slice = 0;
  bool first = true, checked = false;
  bool shard_ok = true;       // false: sharding given up for this state (a rank's range cannot fit); every step replicated
  uint64_t replays = 0, sharded_steps = 0, replicated_steps = 0, replans = 0;
  bool direct() const { return prm.kind == PB200_SIMPLE_ASTRO || !(prm.theta > 0.0); }
  ~MultiSim() {
    for (RankCtx* c : local) {
      cudaSetDevice(c->device);
      if (c->stream) cudaStreamSynchronize(c->stream);
      for (int r = 0; r < 8; ++r)
        for (int k = 0; k < kPeerBufs; ++k)
          if (c->mapped[r][k]) cudaIpcCloseMemHandle(c->mapped[r][k]);
      if (c->comm && g_nccl.so) g_nccl.CommDestroy(c->comm);
      c->ws.release_all();
      DevBuf* bufs[] = {&c->cur, &c->prev, &c->vel, &c->fixed, &c->ext, &c->ck_cur, &c->ck_prev, &c->ck_vel,
                        &c->peer_stage, &c->cnt};
      for (DevBuf* b : bufs) b->release();
      c->h_pos.release();
      c->h_vel.release();
      c->h_fixed.release();
      if (c->ev0) cudaEventDestroy(c->ev0);
      if (c->ev1) cudaEventDestroy(c->ev1);
      if (c->stream) cudaStreamDestroy(c->stream);
      delete c;
    }
  }
};

#define FOR_LOCAL(ms, c)          \
  for (RankCtx * c : (ms).local)  \
    if (cudaSetDevice(c->device) != cudaSuccess) { \
      set_error("cudaSetDevice(%d) failed", c->device); \
      return cudaErrorInvalidDevice; \
    } else

cudaError_t group_start(MultiSim& m) {
  if (m.local.size() > 1) PB_NCCL(g_nccl.GroupStart());
  return cudaSuccess;
}
cudaError_t group_end(MultiSim& m) {
  if (m.local.size() > 1) PB_NCCL(g_nccl.GroupEnd());
  return cudaSuccess;
}

// in-place all-gather of `block` bytes per rank on every local rank's stream
cudaError_t all_gather_blocks(MultiSim& m, DevBuf RankCtx::*unused, size_t block, void* (*buf_of)(RankCtx*)) {
  (void)unused;
  PB_PASS(group_start(m));
  FOR_LOCAL(m, c) {
    char* base = static_cast<char*>(buf_of(c));
    PB_NCCL(g_nccl.AllGather(base + size_t(c->rank) * block, base, block, kNcclChar, c->comm, c->stream));
  }
  PB_PASS(group_end(m));
  return cudaSuccess;
}

void* buf_acc(RankCtx* c) { return c->ws.acc.p; }
void* buf_peer(RankCtx* c) { return c->peer_stage.p; }

cudaError_t sync_all(MultiSim& m) {
  FOR_LOCAL(m, c) PB_CUDA(cudaStreamSynchronize(c->stream));
  return cudaSuccess;
}

// (re)map every rank's buffers into every local rank: direct pointers inside a process (peer access enabled),
// CUDA IPC handles across processes.  Collective (one small ncclAllGather of the records); synchronises.
cudaError_t exchange_peers(MultiSim& m) {
  const uint64_t pid = uint64_t(getpid());
  std::vector<std::vector<PeerRecord>> all(m.local.size(), std::vector<PeerRecord>(m.world));
  FOR_LOCAL(m, c) {
    PeerRecord rec;
    std::memset(&rec, 0, sizeof rec);
    rec.pid = pid;
    rec.device = uint64_t(c->device);
    rec.capacity = c->ws.cell_cap;
    const ShardState& sh = c->ws.shard;
    const void* p[kPeerBufs] = {c->ws.c_centre_ext.p, c->ws.c_com.p, c->ws.c_skip.p, sh.top_info.p,
                                sh.top_com.p,         sh.top_meta.p, sh.xacc.p,      sh.flags.p};
    for (int k = 0; k < kPeerBufs; ++k) {
      rec.ptr[k] = reinterpret_cast<uint64_t>(p[k]);
      if (m.local.size() < size_t(m.world)) PB_CUDA(cudaIpcGetMemHandle(&rec.handle[k], const_cast<void*>(p[k])));
    }
    PB_PASS(c->peer_stage.ensure(size_t(m.world) * sizeof(PeerRecord)));
    PB_CUDA(cudaStreamSynchronize(c->stream));
    PB_CUDA(cudaMemcpy(static_cast<char*>(c->peer_stage.p) + size_t(c->rank) * sizeof(PeerRecord), &rec, sizeof rec,
                       cudaMemcpyHostToDevice));
  }
  PB_PASS(all_gather_blocks(m, nullptr, sizeof(PeerRecord), buf_peer));
  size_t li = 0;
  FOR_LOCAL(m, c) {
    PB_CUDA(cudaMemcpyAsync(all[li].data(), c->peer_stage.p, size_t(m.world) * sizeof(PeerRecord), cudaMemcpyDeviceToHost,
                            c->stream));
    ++li;
  }
  PB_PASS(sync_all(m));
  li = 0;
  FOR_LOCAL(m, c) {
    ShardPeers& sp = c->ws.shard.peers;
    uint64_t cap = ~0ull;
    for (int r = 0; r < m.world; ++r) {
      const PeerRecord& rec = all[li][r];
      cap = std::min(cap, rec.capacity);
      void* q[kPeerBufs];
      if (rec.pid == pid) {
        if (int(rec.device) != c->device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(int(rec.device), 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (e != cudaSuccess) {
            set_error("no peer access from device %d to device %d: %s", c->device, int(rec.device), cudaGetErrorString(e));
            return e;
          }
        }
        for (int k = 0; k < kPeerBufs; ++k) q[k] = reinterpret_cast<void*>(rec.ptr[k]);
      } else {
        for (int k = 0; k < kPeerBufs; ++k) {
          const bool same = c->have_peers && c->mapped[r][k] && c->opened[r].pid == rec.pid &&
                            !std::memcmp(&c->opened[r].handle[k], &rec.handle[k], sizeof rec.handle[k]);
          if (!same) {
            if (c->mapped[r][k]) {
              cudaIpcCloseMemHandle(c->mapped[r][k]);
              c->mapped[r][k] = nullptr;
            }
            cudaError_t e = cudaIpcOpenMemHandle(&c->mapped[r][k], rec.handle[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
              set_error("cudaIpcOpenMemHandle (buffer %d of rank %d on rank %d): %s", k, r, c->rank, cudaGetErrorString(e));
              return e;
            }
          }
          q[k] = c->mapped[r][k];
        }
      }
      c->opened[r] = rec;
      sp.centre_ext[r] = q[0];
      sp.com[r] = q[1];
      sp.skip[r] = q[2];
      sp.top_info[r] = q[3];
      sp.top_com[r] = q[4];
      sp.top_meta[r] = q[5];
      sp.xacc[r] = q[6];
      sp.flags[r] = q[7];
    }
    sp.capacity = uint32_t(std::min<uint64_t>(cap, 0xfffffff0ull));
    c->have_peers = true;
    ++li;
  }
  // nobody may start storing into a peer before that peer has finished this exchange (its buffers may be new)
  PB_PASS(all_gather_blocks(m, nullptr, sizeof(PeerRecord), buf_peer));
  PB_PASS(sync_all(m));
  return cudaSuccess;
}

cudaError_t materialise_velocities(MultiSim& m) {
  FOR_LOCAL(m, c) {
    if (!c->vel_stale) continue;
    PB_PASS(verlet_velocity(c->cur.as<double4>(), c->prev.as<double4>(), c->vel.as<double4>(), m.n, m.dt, c->stream, c->ls));
    c->vel_stale = false;
  }
  return cudaSuccess;
}

// the lean verlet bookkeeping shared by the three kinds of step (same protocol as engine.cu's resident loop)
struct LeanSlots {
  unsigned long long *out, *zero, *last;
};
cudaError_t lean_prepare(RankCtx* c, LeanSlots* s) {
  PB_PASS(c->ext.ensure(32));
  if (c->ext_dirty) {
    PB_CUDA(cudaMemsetAsync(c->ext.p, 0, 16, c->stream));
    c->ext_dirty = false;
  }
  const int out = c->ext_slot ^ 1;
  s->out = c->ext.as<unsigned long long>() + out;
  s->zero = c->ext.as<unsigned long long>() + c->ext_slot;
  s->last = c->ext.as<unsigned long long>() + 2;
  return cudaSuccess;
}
void lean_done(RankCtx* c) {
  std::swap(c->cur, c->prev);  // x_{n+1} was written over x_{n-1}
  c->ws.pos64 = c->cur.as<double4>();
  c->ext_slot ^= 1;
  c->ext_ready = true;
  c->vel_stale = true;
}

// One step with the whole single-GPU evaluation on every rank (no communication).  first / euler-like steps
// use the general kernel with the stored velocities.
cudaError_t step_replicated(MultiSim& m, bool host_check) {
  const bool lean = !m.first;
  if (!lean) PB_PASS(materialise_velocities(m));
  FOR_LOCAL(m, c) {
    c->ws.pos64 = c->cur.as<double4>();
    c->ws.extent_pre = (lean && c->ext_ready) ? c->ext.as<unsigned long long>() + c->ext_slot : nullptr;
    PB_PASS(gravity_evaluate(c->ws, m.prm, 0, m.n, c->stream, c->ls, host_check));
    c->ws.extent_pre = nullptr;
    if (lean) {
      LeanSlots s;
      PB_PASS(lean_prepare(c, &s));
      PB_PASS(verlet_update_lean(c->cur.as<double4>(), c->prev.as<double4>(), c->ws.acc.as<float4>(), m.n, m.dt, s.out,
                                 s.zero, s.last, c->stream, c->ls));
      lean_done(c);
    } else {
      PB_PASS(verlet_update(c->cur.as<double4>(), c->prev.as<double4>(), c->vel.as<double4>(), c->ws.acc.as<float4>(),
                            nullptr, m.n, m.dt, 1, c->stream, c->ls));
      c->ext_ready = false;
      c->ext_dirty = true;
    }
  }
  m.first = false;
  m.checked = true;
  m.replicated_steps += 1;
  return cudaSuccess;
}

// after a replicated step: cuts + splitters for the sharded builds, peer tables
cudaError_t plan_shards(MultiSim& m) {
  FOR_LOCAL(m, c) {
    PB_PASS(gravity_shard_setup(c->ws, m.prm.kind, c->rank, m.world, m.n));
    if (!gravity_shard_fits(c->ws)) {  // more bodies per rank than the shared-memory bucket sort takes: stay replicated
      m.shard_ok = false;
      return cudaSuccess;
    }
    PB_PASS(gravity_shard_plan(c->ws, c->stream, c->ls));
  }
  PB_PASS(exchange_peers(m));
  return cudaSuccess;
}

cudaError_t step_sharded(MultiSim& m) {
  // three phases per rank, no collective call: the kernels store into the peers' buffers and signal / wait on
  // epoch flags (gravity.cu).  With several ranks in this process the phases are enqueued rank by rank so that
  // no device waits for work this thread has not enqueued yet.
  FOR_LOCAL(m, c) {
    c->ws.pos64 = c->cur.as<double4>();
    c->ws.extent_pre = c->ext_ready ? c->ext.as<unsigned long long>() + c->ext_slot : nullptr;
    PB_PASS(gravity_shard_build(c->ws, m.prm, c->stream, c->ls));
    c->ws.extent_pre = nullptr;
  }
  FOR_LOCAL(m, c) PB_PASS(gravity_shard_walk(c->ws, m.prm, c->stream, c->ls));
  FOR_LOCAL(m, c) {
    PB_PASS(gravity_shard_scatter(c->ws, c->stream, c->ls));
    LeanSlots s;
    PB_PASS(lean_prepare(c, &s));
    PB_PASS(verlet_update_lean(c->cur.as<double4>(), c->prev.as<double4>(), c->ws.acc.as<float4>(), m.n, m.dt, s.out,
                               s.zero, s.last, c->stream, c->ls, SHARD_ACC_STRIDE));
    lean_done(c);
  }
  m.sharded_steps += 1;
  return cudaSuccess;
}

// direct sum: targets sharded by body index, accelerations all-gathered (equal slices of S = ceil(n / world))
cudaError_t step_direct(MultiSim& m) {
  const bool lean = !m.first;
  if (!lean) PB_PASS(materialise_velocities(m));
  FOR_LOCAL(m, c) {
    c->ws.pos64 = c->cur.as<double4>();
    PB_PASS(c->ws.acc.ensure(m.slice * size_t(m.world) * sizeof(float4)));
    const size_t t0 = std::min(m.n, m.slice * size_t(c->rank)), t1 = std::min(m.n, m.slice * size_t(c->rank + 1));
    PB_PASS(gravity_evaluate(c->ws, m.prm, t0, t1, c->stream, c->ls, false));
  }
  if (m.world > 1) PB_PASS(all_gather_blocks(m, nullptr, m.slice * sizeof(float4), buf_acc));
  FOR_LOCAL(m, c) {
    if (lean) {
      LeanSlots s;
      PB_PASS(lean_prepare(c, &s));
      PB_PASS(verlet_update_lean(c->cur.as<double4>(), c->prev.as<double4>(), c->ws.acc.as<float4>(), m.n, m.dt, s.out,
                                 s.zero, s.last, c->stream, c->ls));
      lean_done(c);
    } else {
      PB_PASS(verlet_update(c->cur.as<double4>(), c->prev.as<double4>(), c->vel.as<double4>(), c->ws.acc.as<float4>(),
                            nullptr, m.n, m.dt, 1, c->stream, c->ls));
    }
  }
  m.first = false;
  return cudaSuccess;
}

// the ranks' verdicts on the builds since the last check, made one verdict (max over ranks), then applied
// to every workspace - all ranks take the same decision (restore + replay, or go on)
cudaError_t collective_check(MultiSim& m, TreeCheck* out) {
  if (m.world > 1) {
    PB_PASS(group_start(m));
    FOR_LOCAL(m, c) {
      if (!c->ws.sticky.p) continue;
      PB_NCCL(g_nccl.AllReduce(c->ws.sticky.p, c->ws.sticky.p, 8, kNcclUint32, kNcclMax, c->comm, c->stream));
    }
    PB_PASS(group_end(m));
  }
  *out = TreeCheck();
  FOR_LOCAL(m, c) {
    TreeCheck chk;
    PB_PASS(gravity_check(c->ws, c->stream, &chk));
    if (c == m.local[0]) *out = chk;
    if (c->ws.shard.flags.p) {
      uint32_t timed_out = 0;
      PB_CUDA(cudaMemcpy(&timed_out, c->ws.shard.flags.as<uint32_t>() + SHARD_TIMEOUT, 4, cudaMemcpyDeviceToHost));
      if (timed_out) {
        set_error("sharded step: rank %d waited in vain for a peer's signal (a rank stopped or lost its mapping)", c->rank);
        return cudaErrorUnknown;
      }
    }
  }
  return cudaSuccess;
}

// The cuts between the ranks' key ranges are fixed by the plan; bodies drift across them.  Every rank holds every
// rank's body count of the last sharded step (meta): when the fullest rank has used up half of its headroom over
// n / world, the next step runs replicated and the shards are planned afresh (same verdict on every rank).
cudaError_t rebalance_if_drifted(MultiSim& m) {
  if (m.world < 2 || !m.shard_ok || m.sharded_steps == 0) return cudaSuccess;
  bool replan = false;
  FOR_LOCAL(m, c) {
    const ShardState& sh = c->ws.shard;
    if (!sh.planned || !sh.top_meta.p) continue;
    uint32_t meta[32];
    PB_CUDA(cudaMemcpy(meta, sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * 32u, sizeof meta, cudaMemcpyDeviceToHost));
    const size_t fair = m.n / size_t(m.world);
    for (int r = 0; r < m.world; ++r)
      if (size_t(meta[1 + r]) > fair + (sh.n_cap - fair) / 2) replan = true;
  }
  if (replan) {
    for (RankCtx* c : m.local) c->ws.shard.planned = false;
    m.replans += 1;
  }
  return cudaSuccess;
}

cudaError_t run_steps(MultiSim& m, size_t steps) {
  if (m.n == 0) return cudaSuccess;
  if (m.direct()) {
    for (size_t i = 0; i < steps; ++i) PB_PASS(step_direct(m));
    return cudaSuccess;
  }
  const size_t bytes = m.n * sizeof(double4);
  while (steps) {
    const size_t chunk = steps < 32 ? steps : 32;
    FOR_LOCAL(m, c) {
      PB_PASS(c->ck_cur.ensure(bytes));
      PB_PASS(c->ck_prev.ensure(bytes));
      PB_PASS(c->ck_vel.ensure(bytes));
      PB_CUDA(cudaMemcpyAsync(c->ck_cur.p, c->cur.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
      PB_CUDA(cudaMemcpyAsync(c->ck_prev.p, c->prev.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
      PB_CUDA(cudaMemcpyAsync(c->ck_vel.p, c->vel.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    const bool first_at_ck = m.first;
    std::vector<bool> stale_at_ck;
    for (RankCtx* c : m.local) stale_at_ck.push_back(c->vel_stale);
    for (size_t i = 0; i < chunk; ++i) {
      const bool sharded = m.world > 1 && m.shard_ok && !m.first && m.local[0]->ws.shard.planned;
      if (sharded) {
        PB_PASS(step_sharded(m));
      } else {
        PB_PASS(step_replicated(m, !m.checked));
        if (m.world > 1 && m.shard_ok) PB_PASS(plan_shards(m));
      }
    }
    TreeCheck chk;
    PB_PASS(collective_check(m, &chk));
    if (chk.ok()) PB_PASS(rebalance_if_drifted(m));
    if (!chk.ok()) {
      if (chk.sort_error) {
        set_error("radix sort look-back did not complete");
        return cudaErrorUnknown;
      }
      size_t li = 0;
      FOR_LOCAL(m, c) {
        PB_CUDA(cudaMemcpyAsync(c->cur.p, c->ck_cur.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
        PB_CUDA(cudaMemcpyAsync(c->prev.p, c->ck_prev.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
        PB_CUDA(cudaMemcpyAsync(c->vel.p, c->ck_vel.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
        c->vel_stale = stale_at_ck[li++];
        c->ext_ready = false;
        c->ext_dirty = true;
        c->ws.pos64 = c->cur.as<double4>();
        c->ws.shard.planned = false;
      }
      m.first = first_at_ck;
      m.replays += 1;
      if (chk.shard_overflow && m.replays > 4) m.shard_ok = false;  // this state does not balance at level-K granularity
      for (size_t i = 0; i < chunk; ++i) PB_PASS(step_replicated(m, true));
      if (m.world > 1 && m.shard_ok) PB_PASS(plan_shards(m));
    }
    steps -= chunk;
  }
  return cudaSuccess;
}

__global__ void __launch_bounds__(256) count_sharded_kernel(const char* __restrict__ xacc, size_t n_cap,
                                                            const uint32_t* __restrict__ n_locals,
                                                            unsigned long long* __restrict__ out) {
  const unsigned r = blockIdx.y;
  const size_t n_r = n_locals[r];
  const float4* a = reinterpret_cast<const float4*>(xacc + size_t(r) * (n_cap * 20));
  unsigned long long s = 0;
  for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < n_r; j += size_t(gridDim.x) * blockDim.x)
    s += __float_as_uint(a[j].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

}  // namespace
}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_comm_unique_id(uint8_t* out128) {
  if (!out128) return -1;
  if (!g_nccl.load()) return -1;
  NcclId id;
  if (g_nccl.GetUniqueId(&id) != 0) {
    set_error("ncclGetUniqueId failed");
    return -1;
  }
  std::memcpy(out128, id.b, 128);
  return 0;
}

void* pb200_msim_create(int kind, double theta, double e, double dt, int world, int n_local, const int* local_ranks,
                        const int* devices, const uint8_t* nccl_id128) {
  if (kind < PB200_ASTRO || kind > PB200_SIMPLE_ASTRO || world < 1 || world > 8 || n_local < 1 || n_local > world ||
      !local_ranks || !devices) {
    set_error("bad arguments to pb200_msim_create (world 1..8, 1 <= n_local <= world)");
    return nullptr;
  }
  if (n_local != world && !nccl_id128) {
    set_error("pb200_msim_create: ranks spread over several processes need a shared id (pb200_comm_unique_id)");
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    set_error("no CUDA device available; physim_b200 has no CPU fallback");
    return nullptr;
  }
  auto* m = new MultiSim();
  m->prm.kind = kind;
  m->prm.theta = std::isnan(theta) ? 1.0 : theta;
  m->prm.easing = std::isnan(e) ? 1.0 : std::fabs(e);
  m->dt = dt;
  m->world = world;
  auto fail = [&](const char* what) -> void* {
    std::fprintf(stderr, "[physim_b200] msim create failed (%s): %s\n", what, g_error);
    delete m;
    return nullptr;
  };
  for (int i = 0; i < n_local; ++i) {
    if (devices[i] < 0 || devices[i] >= count || local_ranks[i] < 0 || local_ranks[i] >= world) {
      set_error("rank %d / device %d out of range (%d devices visible)", local_ranks[i], devices[i], count);
      return fail("arguments");
    }
    auto* c = new RankCtx();
    c->rank = local_ranks[i];
    c->device = devices[i];
    m->local.push_back(c);
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
      set_error("device %d: %s", c->device, cudaGetErrorString(cudaGetLastError()));
      return fail("stream");
    }
  }
  if (world > 1) {
    if (!g_nccl.load()) return fail("nccl");
    NcclId id;
    if (nccl_id128) std::memcpy(id.b, nccl_id128, 128);
    else if (g_nccl.GetUniqueId(&id) != 0) {
      set_error("ncclGetUniqueId failed");
      return fail("nccl id");
    }
    int rc = 0;
    if (n_local > 1) rc |= g_nccl.GroupStart();
    for (RankCtx* c : m->local) {
      cudaSetDevice(c->device);
      rc |= g_nccl.CommInitRank(&c->comm, world, id, c->rank);
    }
    if (n_local > 1) rc |= g_nccl.GroupEnd();
    if (rc != 0) {
      set_error("ncclCommInitRank failed (%s)", g_nccl.GetErrorString(rc));
      return fail("nccl init");
    }
  }
  return m;
}

void pb200_msim_destroy(void* h) { delete static_cast<MultiSim*>(h); }

int pb200_msim_upload(void* h, const Entity* state, size_t n) {
  if (!h || (n && !state)) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  m.n = n;
  m.slice = (n + size_t(m.world) - 1) / size_t(m.world);
  m.first = true;
  m.checked = false;
  m.shard_ok = true;
  if (n == 0) return 0;
  auto run = [&]() -> cudaError_t {
    FOR_LOCAL(m, c) {
      PB_PASS(c->cur.ensure(n * sizeof(double4)));
      PB_PASS(c->prev.ensure(n * sizeof(double4)));
      PB_PASS(c->vel.ensure(n * sizeof(double4)));
      PB_PASS(c->fixed.ensure(n));
      PB_PASS(c->h_pos.ensure(n * sizeof(double4)));
      PB_PASS(c->h_vel.ensure(n * sizeof(double4)));
      PB_PASS(c->h_fixed.ensure(n));
      double4* hp = c->h_pos.as<double4>();
      double4* hv = c->h_vel.as<double4>();
      uint8_t* hf = c->h_fixed.as<uint8_t>();
      HostPool::instance().parallel_for(n, size_t(1) << 12, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
          hp[i] = make_double4(state[i].x, state[i].y, state[i].z, state[i].mass);
          hv[i] = make_double4(state[i].vx, state[i].vy, state[i].vz, 0.0);
          hf[i] = state[i].fixed ? 1 : 0;
        }
      });
      PB_CUDA(cudaMemcpyAsync(c->cur.p, hp, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
      PB_CUDA(cudaMemcpyAsync(c->vel.p, hv, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
      PB_CUDA(cudaMemcpyAsync(c->fixed.p, hf, n, cudaMemcpyHostToDevice, c->stream));
      c->ws.pos64 = c->cur.as<double4>();
      c->ws.fixed = c->fixed.as<uint8_t>();
      c->ws.n = n;
      c->ws.n_cells = 0;
      c->ws.shard.planned = false;
      c->vel_stale = false;
      c->ext_ready = false;
      c->ext_dirty = true;
    }
    return sync_all(m);
  };
  if (run() != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] msim upload failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

/* `cube n seed spin mass size centre` (astro/src/initialisers.rs:82-106) generated on every rank's device */
int pb200_msim_generate_cube(void* h, size_t n, uint64_t seed, double spin, double mass, double size,
                             const double* centre3) {
  if (!h) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  const double zero[3] = {0.0, 0.0, 0.0};
  const double* centre = centre3 ? centre3 : zero;
  m.n = n;
  m.slice = (n + size_t(m.world) - 1) / size_t(m.world);
  m.first = true;
  m.checked = false;
  m.shard_ok = true;
  if (n == 0) return 0;
  auto run = [&]() -> cudaError_t {
    FOR_LOCAL(m, c) {
      PB_PASS(c->cur.ensure(n * sizeof(double4)));
      PB_PASS(c->prev.ensure(n * sizeof(double4)));
      PB_PASS(c->vel.ensure(n * sizeof(double4)));
      PB_PASS(c->fixed.ensure(n));
      PB_PASS(generate_cube(c->cur.as<double4>(), c->vel.as<double4>(), c->fixed.as<uint8_t>(), n, seed, spin, mass,
                            size, centre, c->stream, c->ls));
      c->ws.pos64 = c->cur.as<double4>();
      c->ws.fixed = c->fixed.as<uint8_t>();
      c->ws.n = n;
      c->ws.n_cells = 0;
      c->ws.shard.planned = false;
      c->vel_stale = false;
      c->ext_ready = false;
      c->ext_dirty = true;
    }
    return sync_all(m);
  };
  if (run() != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] msim generate failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

int pb200_msim_run(void* h, size_t steps) {
  if (!h) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  if (run_steps(m, steps) != cudaSuccess || sync_all(m) != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] msim run failed: %s\n", g_error);
    return -1;
  }
  return 0;
}

/* device time of `steps` steps: CUDA events on every local rank's stream, the maximum over them */
int pb200_msim_run_timed(void* h, size_t steps, float* ms) {
  if (!h || !ms) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  auto run = [&]() -> cudaError_t {
    PB_PASS(sync_all(m));
    FOR_LOCAL(m, c) PB_CUDA(cudaEventRecord(c->ev0, c->stream));
    PB_PASS(run_steps(m, steps));
    FOR_LOCAL(m, c) PB_CUDA(cudaEventRecord(c->ev1, c->stream));
    PB_PASS(sync_all(m));
    return cudaSuccess;
  };
  if (run() != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] msim run failed: %s\n", g_error);
    return -1;
  }
  float worst = 0.f;
  for (RankCtx* c : m.local) worst = std::max(worst, elapsed_ms(c->ev0, c->ev1));
  *ms = worst;
  return 0;
}

int pb200_msim_download(void* h, Entity* state, size_t n) {
  if (!h || (n && !state)) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  if (n != m.n) {
    set_error("pb200_msim_download: n = %zu but the simulation holds %zu bodies", n, m.n);
    return -1;
  }
  if (n == 0) return 0;
  RankCtx* c = m.local[0];  // the state is replicated: any rank's copy is the state
  auto run = [&]() -> cudaError_t {
    PB_PASS(materialise_velocities(m));
    PB_CUDA(cudaSetDevice(c->device));
    PB_PASS(c->h_pos.ensure(n * sizeof(double4)));
    PB_PASS(c->h_vel.ensure(n * sizeof(double4)));
    PB_CUDA(cudaMemcpyAsync(c->h_pos.p, c->cur.p, n * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(cudaMemcpyAsync(c->h_vel.p, c->vel.p, n * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    return sync_all(m);
  };
  if (run() != cudaSuccess) {
    std::fprintf(stderr, "[physim_b200] msim download failed: %s\n", g_error);
    return -1;
  }
  const double4* p = c->h_pos.as<double4>();
  const double4* v = c->h_vel.as<double4>();
  HostPool::instance().parallel_for(n, size_t(1) << 12, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      state[i].x = p[i].x; state[i].y = p[i].y; state[i].z = p[i].z;
      state[i].vx = v[i].x; state[i].vy = v[i].y; state[i].vz = v[i].z;
    }
  });
  return 0;
}

/* copies of the replicated state on two local ranks, compared on the host: 0 identical, 1 different, -1 error (tests) */
int pb200_msim_replicas_identical(void* h) {
  if (!h) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  if (m.local.size() < 2 || m.n == 0) return 0;
  std::vector<double4> a(m.n), b(m.n);
  if (sync_all(m) != cudaSuccess) return -1;
  cudaSetDevice(m.local[0]->device);
  if (cudaMemcpy(a.data(), m.local[0]->cur.p, m.n * sizeof(double4), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  for (size_t i = 1; i < m.local.size(); ++i) {
    cudaSetDevice(m.local[i]->device);
    if (cudaMemcpy(b.data(), m.local[i]->cur.p, m.n * sizeof(double4), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (std::memcmp(a.data(), b.data(), m.n * sizeof(double4)) != 0) return 1;
  }
  return 0;
}

int pb200_msim_stats(void* h, Pb200Stats* out, uint64_t* sharded_steps, uint64_t* replicated_steps) {
  if (!h || !out) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  std::memset(out, 0, sizeof *out);
  out->n_bodies = m.n;
  out->replays = uint32_t(m.replays);
  if (sharded_steps) *sharded_steps = m.sharded_steps;
  if (replicated_steps) *replicated_steps = m.replicated_steps;
  uint64_t launches = 0;
  for (RankCtx* c : m.local) launches += c->ls.launches;
  out->kernel_launches = launches;
  if (m.n == 0 || m.local.empty()) return 0;
  RankCtx* c = m.local[0];
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  out->sort_mode = uint32_t(c->ws.last_mode);
  out->max_bucket = c->ws.last_max_bucket;
  const ShardState& sh = c->ws.shard;
  if (!m.direct() && m.world > 1 && sh.planned && m.sharded_steps > 0 && sh.top_meta.p) {
    uint32_t meta[32];
    if (cudaMemcpy(meta, sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * 32u, sizeof meta, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    uint64_t cells = 0;
    for (int r = 0; r < m.world; ++r) cells += meta[9 + r];
    out->n_cells = cells;  // (the ranks' tables; the cells above level K exist once more in each)
    if (c->cnt.ensure(8) != cudaSuccess) return -1;
    cudaMemset(c->cnt.p, 0, 8);
    count_sharded_kernel<<<dim3(148 * 2, m.world), 256, 0, c->stream>>>(static_cast<const char*>(sh.xacc.p), sh.n_cap,
                                                                        sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * 32u + 1u,
                                                                        c->cnt.as<unsigned long long>());
    unsigned long long inter = 0;
    if (cudaMemcpyAsync(&inter, c->cnt.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      return -1;
    out->interactions = inter;
  } else if (c->ws.acc.p) {
    uint64_t inter = 0;
    if (gravity_count_interactions(c->ws, c->stream, c->ls, &inter) != cudaSuccess) return -1;
    out->interactions = inter;
    out->n_cells = c->ws.n_cells;
  }
  return 0;
}

/* per-rank bodies of the last sharded step (balance of the cuts); returns world, or -1 */
int pb200_msim_rank_counts(void* h, uint32_t* bodies, uint32_t* cells) {
  if (!h) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  RankCtx* c = m.local[0];
  const ShardState& sh = c->ws.shard;
  if (!sh.top_meta.p || m.sharded_steps == 0) return -1;
  uint32_t meta[32];
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
      cudaMemcpy(meta, sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * 32u, sizeof meta, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  for (int r = 0; r < m.world; ++r) {
    if (bodies) bodies[r] = meta[1 + r];
    if (cells) cells[r] = meta[9 + r];
  }
  return m.world;
}

/* diagnostics of local rank 0's shard plan: out[0..8] the key cuts the next build will use, out[9] bodies kept by the
   last sharded build, out[10] epoch, out[11] capacity in bodies, out[12] "a wait gave up" flag */
int pb200_msim_debug_shard(void* h, uint64_t* out13) {
  if (!h || !out13) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  RankCtx* c = m.local[0];
  const ShardState& sh = c->ws.shard;
  std::memset(out13, 0, 13 * 8);
  if (!sh.cuts.p || !sh.flags.p) return -1;
  uint32_t nl = 0, to = 0;
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
      cudaMemcpy(out13, sh.cuts.p, 9 * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(&nl, sh.n_local.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(&to, sh.flags.as<uint32_t>() + SHARD_TIMEOUT, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  out13[9] = nl;
  out13[10] = sh.epoch;
  out13[11] = sh.n_cap;
  out13[12] = to;
  return 0;
}

/* accelerations of the last step's force evaluation, original order (tests; gathers on the host) */
int pb200_msim_last_accelerations(void* h, Acceleration* acc, size_t n) {
  if (!h || (n && !acc)) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  if (n != m.n || n == 0) return -1;
  RankCtx* c = m.local[0];
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  const ShardState& sh = c->ws.shard;
  if (!m.direct() && m.world > 1 && sh.planned && m.sharded_steps > 0) {
    uint32_t meta[32];
    if (cudaMemcpy(meta, sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * 32u, sizeof meta, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    std::vector<float4> a(sh.n_cap);
    std::vector<uint32_t> p(sh.n_cap);
    for (int r = 0; r < m.world; ++r) {
      const char* block = static_cast<const char*>(sh.xacc.p) + size_t(r) * sh.xacc_block_bytes();
      if (cudaMemcpy(a.data(), block, sh.n_cap * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(p.data(), block + sh.n_cap * sizeof(float4), sh.n_cap * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
      for (uint32_t j = 0; j < meta[1 + r]; ++j)
        if (p[j] < n) acc[p[j]] = Acceleration{double(a[j].x), double(a[j].y), double(a[j].z)};
    }
    return 0;
  }
  std::vector<float4> a(n);
  if (!c->ws.acc.p || cudaMemcpy(a.data(), c->ws.acc.p, n * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  for (size_t i = 0; i < n; ++i) acc[i] = Acceleration{double(a[i].x), double(a[i].y), double(a[i].z)};
  return 0;
}

int pb200_msim_profile(void* h, int enable) {
  if (!h) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  for (RankCtx* c : m.local) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->ls.reset_profile();
    c->ls.profiling = enable != 0;
  }
  return 0;
}

/* per-kernel device times of local rank 0 (JSON array, as pb200_sim_profile_report) */
int pb200_msim_profile_report(void* h, char* buf, size_t cap) {
  if (!h || !buf || cap < 3) return -1;
  auto& m = *static_cast<MultiSim*>(h);
  std::lock_guard<std::mutex> lk(m.mu);
  RankCtx* c = m.local[0];
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->ls.collect();
  std::string out = "[";
  for (size_t i = 0; i < c->ls.totals.size(); ++i) {
    char row[256];
    const KernelTime& k = c->ls.totals[i];
    std::snprintf(row, sizeof row, "%s{\"kernel\":\"%s\",\"launches\":%llu,\"ms\":%.6f}", i ? "," : "", k.name,
                  static_cast<unsigned long long>(k.launches), k.ms);
    out += row;
  }
  out += "]";
  if (out.size() + 1 > cap) return -1;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

}  // extern "C"

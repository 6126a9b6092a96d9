// common.cuh — shared declarations for the physim_b200 CUDA engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/physim_b200.h"

static_assert(sizeof(Entity) == 80, "physim Entity is 80 bytes (physim-core/src/lib.rs:16-30)");
static_assert(sizeof(Acceleration) == 24, "physim Acceleration is 24 bytes (lib.rs:32-37)");
static_assert(offsetof(Entity, mass) == 56 && offsetof(Entity, id) == 64 && offsetof(Entity, fixed) == 72,
              "Entity field offsets");

namespace pb200 {

void set_error(const char* fmt, ...);
extern thread_local char g_error[512];

#define PB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::pb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return _e;                                                                              \
    }                                                                                         \
  } while (0)

// propagate an error whose message was already recorded by the callee
#define PB_PASS(expr)                  \
  do {                                 \
    cudaError_t _e = (expr);           \
    if (_e != cudaSuccess) return _e;  \
  } while (0)

// Grow-only device / pinned-host buffers.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
      return e;
    }
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e));
      return e;
    }
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// ---- sharded Barnes-Hut: what one rank keeps besides its GravityWorkspace (see gravity.cu, "Sharded Barnes-Hut") ----
struct ShardPeers {  // every rank's buffers as this device sees them (own rank: local pointers; others: NVLink peer mappings)
  const void* centre_ext[8];  // cell tables, read by the walk
  const void* com[8];
  const void* skip[8];
  void* top_info[8];          // dense top tree + per-rank counts, exchange buffer, flags: WRITTEN by the peers
  void* top_com[8];
  void* top_meta[8];
  void* xacc[8];
  void* flags[8];
  uint32_t capacity;          // cells (the smallest table of any rank)
};
struct ShardState {
  int rank = 0, world = 1;
  bool planned = false;   // cuts and splitters for a sharded build exist (left by gravity_shard_plan / the last sharded step)
  size_t n_cap = 0;       // capacity of the per-rank arrays in bodies (the same on every rank)
  uint32_t epoch = 0;     // sharded steps taken: the value the ranks signal each other with (same sequence on every rank)
  DevBuf cuts;            // u64[world + 1]: key cuts of the NEXT build
  DevBuf n_local;         // u32: bodies in this rank's range (left by the sort of the current build)
  DevBuf slot_cell;       // u32[level-K prefixes]
  DevBuf top_ce, top_com, top_info;  // dense top tree (levels 0..K); its level-K entries are written by their owners
  DevBuf top_meta;        // u32: per-rank bodies / cells / abandoned-build flags, double-buffered by epoch parity
  DevBuf xacc;            // [world] x { float4 acc[n_cap], u32 perm[n_cap] }: block r written by rank r's walk
  DevBuf flags;           // u32 epochs per phase and rank (SHARD_FLAG_*), stored by the peers
  ShardPeers peers;
  size_t xacc_block_bytes() const { return n_cap * 20; }
  void release();
};

// ---- device workspace of one gravity evaluation ------------------------------------------------
// Layout in HBM (N bodies, C cells ≈ 1.5 N):
//   pos64   double4[N]  {x,y,z,m}, original order            32 B/body   (input of every kernel)
//   fixed   u8[N]                                              1 B/body
//   src4    float4[N]   {x,y,z,m} fp32 (direct sum sources)   16 B/body
//   key[2]  u64[N], idx[2] u32[N]  (radix sort ping-pong)      24 B/body
//   spos64  double4[N]  sorted {x,y,z,m}                       32 B/body
//   ab      uchar2[N], cell_start u32[N+1]                      6 B/body
//   nsv1    u8[N + N/16], nsv2 u8[9][N/256]  shared-level byte per body, minima per window / block
//                                                             1.1 B/body
//   cells:  level u8, head/count/skip/parent/arrived u32, centre_ext double4, com double4
//                                                              85 B/cell
//   c_kids  u32[2^DIM][C] child tables of the cells summed bottom-up (sparse: ~5 % of the cells touched),
//   c_ready u32 lists of climb starts                          16-32 + 4 B/cell
//   acc     float4[N]   {ax,ay,az, bits(interactions)}, original order   16 B/body
struct GravityWorkspace {
  // inputs (owned elsewhere when running device-resident)
  const double4* pos64 = nullptr;
  const uint8_t* fixed = nullptr;
  size_t n = 0;
  // extent of pos64 (u64 bits of the double) already reduced on the device by whoever wrote pos64 (the
  // resident verlet step); consumed by the next tree build instead of running extent_kernel
  const unsigned long long* extent_pre = nullptr;
  const unsigned long long* extent_cur = nullptr;  // where the running build reads its extent
  // owned
  DevBuf src4, key0, key1, idx0, idx1, bucket_key, bucket_idx, splitters, nsv1, nsv2, spos64, ab, cell_start, scan_tmp, tile_counts, digit_base,
      extent_bits, tgt_list, tgt_flags;
  DevBuf c_level, c_head, c_count, c_skip, c_parent, c_arrived, c_centre_ext, c_com;
  DevBuf c_kids, c_ready;       // child tables / climb starts of the cells summed bottom-up
  bool parents_filled = false;  // c_parent holds every cell's parent (else only those the bottom-up sums needed)
  DevBuf acc, acc_part, counters, sticky;
  size_t n_cells = 0;   // cells of the last checked evaluation
  size_t cell_cap = 0;  // capacity of the cell arrays
  int tree_dim = 0;     // 2 / 3 after a tree build, 0 otherwise
  int sort_lo = 0;      // lowest key bit the next sort will include (0 = all bits)
  int last_lo = 0;      // ... that the last sort included
  int sort_extra_levels = 0;  // safety margin, grown whenever a truncated sort proved too short
  int sort_mode = 0;    // next sort: 0 global LSD passes, 1..3 bucket sort (chosen by gravity_check)
  int last_mode = 0;    // ... that the last sort used
  int splitter_cur = 0; // which of the two splitter sets the next evaluation reads
  unsigned spl_nb[2] = {0, 0};  // how many buckets each splitter set was written for (0: never written)
  int bucket_ban = 0;   // checks left during which the bucket sort stays off (a bucket's bodies were too alike)
  int bucket_min_mode = 1;  // smallest bucket capacity class still trusted (raised when a bucket overflowed its tile)
  uint32_t last_max_bucket = 0;  // fullest top-8-bit bin seen at the last check
  int unchecked_builds = 0;   // tree builds since the last gravity_check()
  uint32_t last_total = 0;    // verdict of the last gravity_check(), returned again when nothing was built since
  int last_deepest = -1;
  unsigned* sort_err_flag = nullptr;
  // where the sorted keys / permutation ended up after the last sort
  const uint64_t* sorted_key = nullptr;
  const uint32_t* perm = nullptr;
  ShardState shard;
  void release_all();
};

struct GravityParams {
  int kind;       // Pb200Kind
  double theta;   // reference semantics: accept iff half_width / |p - centre| < theta
  double easing;  // added to r^2
};

// Launch bookkeeping.  `launches` always counts; with `profiling` on, every launch is bracketed by a
// CUDA event pair on the launching stream and collect() aggregates device time per kernel name
// (what bench.py reports as the live per-kernel durations behind `roofline`).
struct KernelTime {
  const char* name;
  uint64_t launches;
  double ms;
};
struct LaunchStats {
  static inline thread_local const char* current = "";  // the kernel being launched (read by pb_launch_pdl)
  uint64_t launches = 0;
  bool profiling = false;
  struct Pending {
    const char* name;
    cudaEvent_t e0, e1;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::vector<KernelTime> totals;
  cudaEvent_t get_event() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void begin(const char* name, cudaStream_t st) {
    ++launches;
    current = name;
    if (!profiling) return;
    Pending p{name, get_event(), get_event()};
    cudaEventRecord(p.e0, st);
    pending.push_back(p);
  }
  void end(cudaStream_t st) {
    if (profiling && !pending.empty()) cudaEventRecord(pending.back().e1, st);
  }
  // requires the stream to be idle (caller synchronises)
  void collect() {
    for (const Pending& p : pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.e0, p.e1) != cudaSuccess) {
        cudaGetLastError();
        ms = 0.f;
      }
      bool found = false;
      for (KernelTime& k : totals)
        if (k.name == p.name || !std::strcmp(k.name, p.name)) {
          k.launches += 1;
          k.ms += ms;
          found = true;
          break;
        }
      if (!found) totals.push_back(KernelTime{p.name, 1, ms});
      pool.push_back(p.e0);
      pool.push_back(p.e1);
    }
    pending.clear();
  }
  void reset_profile() {
    collect();
    totals.clear();
  }
  ~LaunchStats() {
    for (const Pending& p : pending) {
      cudaEventDestroy(p.e0);
      cudaEventDestroy(p.e1);
    }
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
  }
};

// Programmatic dependent launch for the kernels of the step's critical chain: the next kernel's CTAs are
// scheduled while this one's last CTAs drain (its launch latency and ramp-up hide behind the tail), and
// wait in pb_pdl_sync() until everything before them in the stream is complete and visible.  A kernel
// launched this way must call pb_pdl_sync() before it touches global memory.  PB200_PDL=0: plain launches.
#ifdef __CUDACC__
__device__ __forceinline__ void pb_pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
#endif

#ifdef __CUDACC__
// Cross-GPU signalling of the sharded step: flags are u32 epochs in the CONSUMER's memory, stored by the producers
// through peer mappings.  Bounded spin: a rank that never signals (a crashed peer) leaves flags[18] set and the
// wait returns - the results are then garbage and the host's next check reports the error instead of a hung GPU.
// [0..7] "rank r's level-K records of epoch e are in place", [8..15] "... accelerations ...", [24] a wait gave up
constexpr int SHARD_FLAG_EXPORT = 0, SHARD_FLAG_WALK = 8, SHARD_TIMEOUT = 24;
constexpr size_t SHARD_ACC_STRIDE = 2;  // gravity_shard_scatter leaves the accelerations as 32-byte records (float4 + padding)
__device__ __forceinline__ void shard_wait_flag(uint32_t* flags, int slot, uint32_t epoch) {
  volatile uint32_t* f = flags + slot;
  unsigned spins = 0;
  while (int32_t(*f - epoch) < 0) {
    __nanosleep(64);
    if (++spins > (1u << 25)) {  // ~ seconds
      flags[SHARD_TIMEOUT] = 1u;
      break;
    }
  }
  __threadfence_system();
}
#endif

inline bool pb_pdl_enabled() {
  static const bool v = !(std::getenv("PB200_PDL") && std::atoi(std::getenv("PB200_PDL")) == 0);
  return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t pb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // PB200_PDL_OFF=kids_kernel,climb_kernel: plain launches for the named kernels only (tuning runs)
  static const char* off = std::getenv("PB200_PDL_OFF");
  cfg.numAttrs = (pb_pdl_enabled() && !(off && *LaunchStats::current && std::strstr(off, LaunchStats::current))) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// PB_LAUNCH(ls, stream, "name", kernel<<<grid, block, smem, stream>>>(args...));
#define PB_LAUNCH(ls, st, name, ...) \
  do {                               \
    (ls).begin(name, st);            \
    __VA_ARGS__;                     \
    (ls).end(st);                    \
  } while (0)

// Computes acc[i] (original order) for targets with original index in [t0, t1) from all n bodies.
// Fixed targets and targets outside the range are left at zero.  *n_cells_out receives the cell count.
// host_check: read the cell total back (one stream sync) and re-run once if the cell table was too
// small; without it the caller must poll gravity_cell_total() before trusting the result.
cudaError_t gravity_evaluate(GravityWorkspace& ws, const GravityParams& prm, size_t t0, size_t t1,
                             cudaStream_t stream, LaunchStats& ls, bool host_check = true);
// Sharded Barnes-Hut, one call per phase.  The exchanges are peer-memory stores issued by the kernels
// themselves (no collective call per step): every phase ends by signalling the step's epoch to all ranks, the
// consumer kernels wait for every rank's signal.
//   gravity_shard_plan    after a full build on every rank (gravity_evaluate over all bodies): first cuts + splitters
//   gravity_shard_build   tree of this rank's key range; its level-K records stored into EVERY rank's dense top tree
//   gravity_shard_walk    cells above level K, walk of this rank's bodies; accelerations (sorted order) and the
//                         permutation stored into block `rank` of EVERY rank's shard.xacc
//   gravity_shard_scatter waits for every rank's block, then ws.acc[i] (original order, all bodies) for the
//                         integrator - every rank advances the whole replicated state
// shard.peers must hold every rank's buffers (peer-mapped) before gravity_shard_build; shard.epoch is advanced
// by gravity_shard_build.
cudaError_t gravity_shard_setup(GravityWorkspace& ws, int kind, int rank, int world, size_t n);
bool gravity_shard_fits(const GravityWorkspace& ws);  // the per-rank capacity is within what the sharded build's sort takes
cudaError_t gravity_shard_plan(GravityWorkspace& ws, cudaStream_t stream, LaunchStats& ls);
cudaError_t gravity_shard_build(GravityWorkspace& ws, const GravityParams& prm, cudaStream_t stream, LaunchStats& ls);
cudaError_t gravity_shard_walk(GravityWorkspace& ws, const GravityParams& prm, cudaStream_t stream, LaunchStats& ls);
cudaError_t gravity_shard_scatter(GravityWorkspace& ws, cudaStream_t stream, LaunchStats& ls);
// Host-side verdict on the last tree build (one small D2H + stream sync).  Also feeds the next
// evaluation: cell-table capacity follows the observed total, the sort drops the key bits below
// the observed tree depth (+2 levels) and is re-validated every time.
// (worst case over every build since the previous check: builds may run unverified in between)
struct TreeCheck {
  uint32_t total = 0;       // cells
  int deepest_shared = -1;  // deepest level shared by two sorted neighbours with different keys
  bool overflow = false;    // total > capacity: the tree kernels bailed out
  bool sort_short = false;  // truncated sort left ties that the dropped bits would have ordered
  bool sort_error = false;  // look-back spin limit hit (never expected)
  uint32_t max_bucket = 0;  // bodies in the fullest bin of the keys' top 8 bits
  bool bucket_overflow = false;  // a bucket-local sort met a bucket larger than its shared-memory tile
  bool shard_overflow = false;   // sharded build: a rank's key range held more bodies than its arrays
  bool ok() const { return !overflow && !sort_short && !sort_error && !bucket_overflow && !shard_overflow; }
};
cudaError_t gravity_check(GravityWorkspace& ws, cudaStream_t stream, TreeCheck* out);
cudaError_t gravity_cell_total(GravityWorkspace& ws, cudaStream_t stream, uint32_t* total);
cudaError_t gravity_fill_parents(GravityWorkspace& ws, cudaStream_t stream, LaunchStats& ls);
// Sum of per-target interaction counters of the last evaluation (synchronises the stream).
cudaError_t gravity_count_interactions(GravityWorkspace& ws, cudaStream_t stream, LaunchStats& ls,
                                       uint64_t* out);

// ---- packing / unpacking at the host boundary ---------------------------------------------------
// AoS Entity (device copy not needed): host packs into {x,y,z,m} + fixed flags (see host_pack.cpp).

// ---- verlet --------------------------------------------------------------------------------------
// In-place update of cur (double4 {x,y,z,m}), prev (double3-as-double4 without mass use), vel.
// first != 0: x1 = x0 + v0 dt + ½ a dt², v1 = v0 + a dt, prev = x0      (verlet.rs:24-50)
// else      : x' = 2x − prev + a dt², v' = (x' − x)/dt, prev = x         (verlet.rs:52-82)
// acc32 (float4, index i - acc_offset... see verlet.cu) or acc64 (3 doubles per body) is used.
// device-resident verlet step (no first-step form): reads x_n (cur), x_{n-1} (prev_inout) and the
// accelerations, writes x_{n+1} over x_{n-1} (the caller swaps the two buffers) and nothing else -
// v_{n+1} = (x_{n+1} - x_n) / dt is derived on demand by verlet_velocity().  Also reduces the extent
// max(|x|,|y|,|z|) of the new positions into *extent_out (atomicMax on the double's bits) and zeroes
// *extent_zero (the slot the next step will reduce into) after saving its value - the extent the build of
// this step used - in *extent_last.
cudaError_t verlet_update_lean(const double4* cur, double4* prev_inout, const float4* acc32, size_t n, double dt,
                               unsigned long long* extent_out, unsigned long long* extent_zero,
                               unsigned long long* extent_last, cudaStream_t st, LaunchStats& ls,
                               size_t acc_stride = 1 /* float4 records between consecutive bodies' accelerations */);
// the same step on every rank of a sharded run: body perm[r][j] takes acc[r][j] for j < n_locals[r] (the gathered
// blocks of ShardState::xacc, n_cap records each)
cudaError_t verlet_velocity(const double4* cur, const double4* prev, double4* vel, size_t n, double dt,
                            cudaStream_t st, LaunchStats& ls);
cudaError_t verlet_update(double4* cur, double4* prev, double4* vel, const float4* acc32,
                          const double* acc64, size_t n, double dt, int first, cudaStream_t stream,
                          LaunchStats& ls, double* out6 = nullptr);

// Entity AoS (80-byte records) on the device <-> the kernels' state arrays (resident host boundary)
cudaError_t entity_split(const void* ent80, size_t n, double4* pos, double4* vel, uint8_t* fixed, cudaStream_t st,
                         LaunchStats& ls);
cudaError_t entity_merge(void* ent80, size_t n, const double4* pos, const double4* vel, cudaStream_t st, LaunchStats& ls);

// One rk4 stage (1..4): folds k_stage into the running sums, writes the next evaluation point
// (stages 1-3; also packed into out6 when given) or the new state (stage 4: out_pos/out_vel/out6).
cudaError_t rk4_stage(int stage, const double4* e_pos, const double4* e_vel, const uint8_t* fixed,
                      double4* t_pos, double4* t_vel, double4* s_pos, double4* s_vel, const float4* acc32,
                      const double* acc64, size_t n, double dt, double4* out_pos, double4* out_vel,
                      double* out6, cudaStream_t stream, LaunchStats& ls);

// `cube` initial conditions generated in place on the device (generate.cu)
cudaError_t generate_cube(double4* pos, double4* vel, uint8_t* fixed, size_t n, uint64_t seed, double spin, double mass,
                          double size, const double centre[3], cudaStream_t st, LaunchStats& ls);

// fp32 FFMA probe
cudaError_t probe_fp32(double* tflops);

}  // namespace pb200

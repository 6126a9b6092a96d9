/* physim_b200.h — C ABI of libphysim_b200.so
 *
 * A drop-in physim plugin (elements `astro`, `astro2`, `simple_astro`) whose transforms run as
 * hand-written sm_100a CUDA kernels, plus the C entry points a thin Rust shim needs to provide
 * the `verlet` integrator (physim's integrator elements have no C ABI).
 *
 * Section 1 restates the host's plugin contract (what cbindgen emits as c_plugin/physim.h).
 * Section 2 is the set of symbols physim's loader resolves in a plugin library.
 * Section 3 is the engine-level API (same kernels, explicit handles; used by the Rust shim,
 *           by tests/ and by bench.py through ctypes).
 *
 * Plain C types only: pointers, sizes, doubles.  Every entry point that needs the GPU fails
 * loudly (message on stderr, then abort() or an error code as documented) when no CUDA device
 * or kernel image is available.  There is no CPU fallback.
 */
#ifndef PHYSIM_B200_H
#define PHYSIM_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. Host contract (reference: c_plugin/physim.h:1-94)
 * ---------------------------------------------------------------------------------------- */

#ifndef PHYSIM_B200_NO_HOST_TYPES

/* c_plugin/physim.h:6-13 ; physim-core/src/plugin/mod.rs (ElementKind) */
typedef enum ElementKind {
  Initialiser,
  Transform,
  Render,
  Synth,
  Transmute,
  Integrator,
} ElementKind;

/* c_plugin/physim.h:15-18 */
typedef enum MessageOrigin {
  Rust = 0,
  C = 1,
} MessageOrigin;

/* c_plugin/physim.h:20-27 */
typedef enum MessagePriority {
  Background,
  Low,
  Normal,
  High,
  RealTime,
  Critical,
} MessagePriority;

/* c_plugin/physim.h:29-35 ; physim-core/src/messages.rs */
typedef struct CMessage {
  enum MessagePriority priority;
  const char *topic;
  const char *message;
  uintptr_t sender_id;
  enum MessageOrigin origin;
} CMessage;

/* c_plugin/physim.h:37-48 ; physim-core/src/lib.rs:16-30 — 80 bytes, id @64, fixed @72 */
typedef struct Entity {
  double x;
  double y;
  double z;
  double vx;
  double vy;
  double vz;
  double radius;
  double mass;
  uintptr_t id;
  bool fixed;
} Entity;

/* c_plugin/physim.h:50-54 ; physim-core/src/lib.rs:32-37 — 24 bytes */
typedef struct Acceleration {
  double x;
  double y;
  double z;
} Acceleration;

/* c_plugin/physim.h:56 */
typedef char *(*RustStringAllocFn)(const char *);

/* c_plugin/physim.h:58-65 ; physim-core/src/plugin/transform.rs:32-50 — field order is ABI */
typedef struct TransformElementAPI {
  void *(*init)(const uint8_t *, uintptr_t);
  void (*transform)(const void *, const struct Entity *, uintptr_t, struct Acceleration *, uintptr_t);
  void (*destroy)(void *);
  char *(*get_property_descriptions)(void *, RustStringAllocFn);
  void (*recv_message)(void *obj, const struct CMessage *msg);
  void (*post_configuration_messages)(void *obj);
} TransformElementAPI;

/* c_plugin/physim.h:70-79 ; physim-core/src/plugin/meta.rs:95-106 */
typedef struct ElementMetaFFI {
  enum ElementKind kind;
  char *name;
  char *plugin;
  char *version;
  char *license;
  char *author;
  char *blurb;
  char *repo;
} ElementMetaFFI;

#endif /* PHYSIM_B200_NO_HOST_TYPES */

/* ------------------------------------------------------------------------------------------
 * 2. Symbols physim's loader resolves (reference call sites in parentheses)
 * ---------------------------------------------------------------------------------------- */

/* "C" — physim-core/src/plugin/discover.rs:327-351 ; c_plugin/plugin.c:28,39-41 */
const char *get_plugin_abi_info(void);
/* "astro,astro2,simple_astro" — discover.rs:353-362 ; replaces astro/src/lib.rs:14-23 for the
 * three transform elements (the generators cube/star/plummer/solar/bar stay in libastro) */
const char *register_plugin(void);
/* physim-core/src/plugin/mod.rs:234-244 ; c_plugin/plugin.c:47-53 */
void set_callback_target(void *target);

/* `astro` — quadtree Barnes-Hut; replaces AstroElement, astro/src/transformers.rs:15-113 */
ElementMetaFFI astro_register(RustStringAllocFn alloc);
const TransformElementAPI *astro_get_api(void);
/* `astro2` — octree Barnes-Hut; replaces AstroOctreeElement, transformers.rs:114-204 */
ElementMetaFFI astro2_register(RustStringAllocFn alloc);
const TransformElementAPI *astro2_get_api(void);
/* `simple_astro` — direct sum; replaces SimpleAstroElement, transformers.rs:206-271 */
ElementMetaFFI simple_astro_register(RustStringAllocFn alloc);
const TransformElementAPI *simple_astro_get_api(void);

/* ------------------------------------------------------------------------------------------
 * 3. Engine-level API
 * ---------------------------------------------------------------------------------------- */

typedef enum Pb200Kind {
  PB200_ASTRO = 0,        /* quadtree on (x,y), 3-D forces  (astro/src/quadtree.rs) */
  PB200_ASTRO2 = 1,       /* octree                         (astro/src/octree.rs)   */
  PB200_SIMPLE_ASTRO = 2, /* all pairs                      (transformers.rs:220-244) */
} Pb200Kind;

/* Counters of the most recent force evaluation / step of a handle. */
typedef struct Pb200Stats {
  uint64_t n_bodies;
  uint64_t n_cells;        /* cells in the tree (0 for direct sum) */
  uint64_t interactions;   /* (target, source-or-cell) pairs evaluated */
  uint64_t kernel_launches;/* kernels launched by this library since the handle was created */
  double extent;           /* root half-width used */
  float ms_h2d, ms_build, ms_force, ms_integrate, ms_d2h; /* CUDA-event times of the last call */
  float ms_host_pack, ms_host_unpack, ms_wall;             /* host wall-clock parts of the last call */
  uint32_t replays;        /* chunks of unverified steps that failed verification and were replayed */
  uint32_t sort_bits;      /* key bits the last sort covered (global passes: tree depth seen + margin; bucket sort: all) */
  uint32_t sort_mode;      /* form of the last sort: 0 global LSD radix passes, 1 / 2 / 3 bucket sort in shared memory (4608 / 8192 / 16384-body buckets) */
  uint32_t max_bucket;     /* bodies in the fullest bin of the keys' top 8 bits at the last check */
} Pb200Stats;

/* Number of CUDA devices visible; <= 0 means the GPU entry points will fail. No context is made. */
int pb200_device_count(void);
/* Select the device used by handles created afterwards on this thread (default 0). */
int pb200_set_device(int device);
const char *pb200_last_error(void);

/* --- gravity transform (same object the plugin's `init` returns) -------------------------
 * pb200_transform_create mirrors TransformElement::new (transformers.rs:71-89,162-180,246-256):
 * theta default 1.0, e -> |e| default 1.0; pass NAN to take a default.  Creation does not touch
 * the GPU (discover.rs:376-387 instantiates and drops every transform at start-up). */
void *pb200_transform_create(int kind, double theta, double e);
/* JSON form used by the plugin `init` (not NUL-terminated; `{}` allowed). NULL on malformed JSON. */
void *pb200_transform_create_json(int kind, const uint8_t *json, size_t len);
void pb200_transform_destroy(void *obj);
/* AstroElement::transform / AstroOctreeElement::transform / SimpleAstroElement::transform
 * (transformers.rs:32-69,123-160,220-244): acc[i] += a_i for every non-fixed body.
 * Host pointers.  Returns 0, or nonzero after printing the CUDA error to stderr. */
int pb200_transform_apply(void *obj, const Entity *state, size_t n, Acceleration *acc, size_t n_acc);
int pb200_transform_stats(void *obj, Pb200Stats *out);
double pb200_transform_theta(void *obj);
double pb200_transform_easing(void *obj);

/* Inspection of the tree built by the last pb200_transform_apply (parity tests).  Any pointer
 * may be NULL.  Sizes: key/perm n; cell_start n+1; per-cell arrays n_cells; centre_ext and
 * com_mass 4 doubles per cell; counts n (interactions per target, original order). */
int pb200_transform_debug_tree(void *obj, uint64_t *key, uint32_t *perm, uint32_t *cell_start,
                               uint8_t *level, uint32_t *head, uint32_t *count, uint32_t *skip,
                               uint32_t *parent, double *centre_ext, double *com_mass,
                               uint32_t *counts);

/* Test hook: pretend the previous evaluation saw a tree that needs only key bits >= sort_lo and
 * n_cells_hint cells, so that the next one exercises the validate-and-retry paths (truncated sort
 * too short, cell table too small). */
int pb200_transform_debug_hint(void *obj, int sort_lo, size_t n_cells_hint);
/* Test hook: form of the next sort (Pb200Stats.sort_mode); a bucket-local form that meets an
 * oversized bucket must be detected and the build re-run with the global passes. */
int pb200_transform_debug_sort_mode(void *obj, int mode);

/* --- verlet (integrators/src/verlet.rs:86-107) -------------------------------------------
 * The Rust shim's IntegratorElement::integrate forwards here.  acc_fn has the meaning of the
 * `&dyn Fn(&[Entity], &mut [Acceleration])` closure built at pipeline.rs:137-141. */
typedef void (*Pb200AccFn)(void *ctx, const Entity *state, size_t n, Acceleration *acc);
void *pb200_verlet_create(void);
void pb200_verlet_destroy(void *v);
/* Generic step: zeroed accelerations -> acc_fn (host) -> first-step or regular update on the GPU.
 * new_state[i] = entities[i] with x,y,z,vx,vy,vz replaced (verlet.rs:24-50,52-82). */
int pb200_verlet_step(void *v, const Entity *entities, Entity *new_state, size_t n,
                      Pb200AccFn acc_fn, void *ctx, double dt);
/* Fused step: the accelerations come from `transform` (a pb200_transform_create handle) and never
 * leave the device: H2D packed state -> tree/force kernels -> verlet kernel -> D2H x,v. */
int pb200_verlet_step_fused(void *v, void *transform, const Entity *entities, Entity *new_state,
                            size_t n, double dt);
/* Opt-in for the fused step: keep the state in HBM between calls and return whole Entity records by one DMA
 * into the caller's (persistent, page-locked on first use) new_state.  CONTRACT: between two calls nothing edits
 * the state - physim's loop then passes a clone of what the previous call returned (pipeline.rs:166-173); a
 * pipeline with transmute elements (pipeline.rs:169-171) must not set it.  Each call compares 2048 sampled
 * entities with the previous output and falls back to a full upload on any difference (another run, a reset,
 * a wholesale edit); an edit of a few bodies can escape the sample, hence the contract.  The Rust shim's
 * `resident` property. */
int pb200_verlet_set_resident(void *v, int on);
int pb200_verlet_resident_counts(void *v, uint64_t *steps_without_upload, uint64_t *steps_with_upload);
int pb200_verlet_stats(void *v, Pb200Stats *out);

/* The composition stock physim runs, in C: an acc_fn that calls a transform element through its plugin vtable
 * (pipeline.rs:137-141 builds that closure; plugin/transform.rs:85-104 is the call).  ctx = &Pb200TransformRef. */
#ifndef PHYSIM_B200_NO_HOST_TYPES
typedef struct Pb200TransformRef {
  const TransformElementAPI *api;
  void *obj;
} Pb200TransformRef;
#endif
void pb200_acc_from_transform(void *ctx, const Entity *state, size_t n, Acceleration *acc);

/* --- euler and rk4 (integrators/src/euler.rs:17-50, rk4.rs:23-183) ------
 * Same handle type and calling convention as verlet; a Rust shim's `euler` / `rk4` elements forward
 * to these.  euler applies x + v dt + a dt²/2, v + a dt every step; rk4 evaluates acc_fn four times
 * (on temporaries, as the reference does) and honours `fixed` (position kept, velocity zeroed). */
typedef enum Pb200Integrator {
  PB200_VERLET = 0,
  PB200_EULER = 1,
  PB200_RK4 = 2,
} Pb200Integrator;
void *pb200_integrator_create(int kind);
void pb200_integrator_destroy(void *g);
int pb200_integrator_step(void *g, const Entity *entities, Entity *new_state, size_t n,
                          Pb200AccFn acc_fn, void *ctx, double dt);
int pb200_integrator_step_fused(void *g, void *transform, const Entity *entities, Entity *new_state,
                                size_t n, double dt);

/* --- device-resident simulation (bench `value`, multi-GPU sharding) ----------------------
 * State lives in HBM across steps: fp64 {x,y,z,m}, previous positions, velocities.
 * rank/world: this handle owns bodies [rank*S, min(n, (rank+1)*S)), S = ceil(n/world), as force targets and
 * integrator state (sampled timing of a target slice; pb200_sim_step_local + the caller's all-gather on the
 * pointers below for a step-by-step multi-rank run).  The multi-GPU simulation loop proper is pb200_msim_* below:
 * it takes its replay decisions collectively, which a per-handle loop with a caller-side exchange cannot. */
void *pb200_sim_create(int kind, double theta, double e, double dt, int rank, int world);
void pb200_sim_destroy(void *sim);
int pb200_sim_upload(void *sim, const Entity *state, size_t n);
/* Instead of pb200_sim_upload: `cube n seed spin mass size centre` (astro/src/initialisers.rs:82-106 over
 * Entity::random, physim-core/src/lib.rs:115-128) generated in place on the device from the reference's
 * ChaCha8 stream; radius 0.02, id 0, fixed false.  centre3 NULL = origin. */
int pb200_sim_generate_cube(void *sim, size_t n, uint64_t seed, double spin, double mass, double size,
                            const double *centre3);
/* Run `steps` steps of the owned slice back to back.  With world > 1 no positions are exchanged:
 * use pb200_sim_step_local + an all-gather for a real multi-rank run (this form serves sampled
 * timing of a target slice). */
int pb200_sim_run(void *sim, size_t steps);
/* pb200_sim_run bracketed by CUDA events on the handle's stream: *ms = device time of the `steps` steps. */
int pb200_sim_run_timed(void *sim, size_t steps, float *ms);
/* Per-kernel device times: enable (resets the table), run steps, then read a JSON array
 * [{"kernel": name, "launches": k, "ms": total}, ...] into buf. */
int pb200_sim_profile(void *sim, int enable);
int pb200_sim_profile_report(void *sim, char *buf, size_t cap);
/* One step for world > 1: forces for the owned targets from the gathered positions, then the
 * verlet update of the owned slice, written into the owned slice of the gather buffer. */
int pb200_sim_step_local(void *sim);
/* Device pointer and byte counts of the gathered {x,y,z,m} fp64 buffer and of this rank's slice. */
int pb200_sim_gather_buffer(void *sim, void **dev_ptr, size_t *total_bytes, size_t *slice_offset,
                            size_t *slice_bytes);
/* Copy the current state back: positions/velocities of all n bodies (world == 1) or of the owned
 * slice (world > 1; `state` still indexes all n). */
int pb200_sim_download(void *sim, Entity *state, size_t n);
/* Accelerations of the last force evaluation, original order, for the owned targets. */
int pb200_sim_last_accelerations(void *sim, Acceleration *acc, size_t n);
int pb200_sim_stats(void *sim, Pb200Stats *out);
/* Integrator of the device-resident loop (default verlet; rk4 only with world == 1). */
int pb200_sim_set_integrator(void *sim, int kind);
/* Override the owned target range [t0, t1) chosen by rank/world at upload (sampled timing of a
 * target slice of a large all-pairs problem). */
int pb200_sim_set_targets(void *sim, size_t t0, size_t t1);
/* Launch on a caller-owned CUDA stream (cudaStream_t; e.g. the stream a collective library uses)
 * instead of the handle's own.  NULL selects the legacy default stream. */
int pb200_sim_set_stream(void *sim, void *stream);
/* CUDA stream (cudaStream_t) the handle launches on, for event timing by the caller. */
void *pb200_sim_stream(void *sim);

/* --- the same simulation on several GPUs of one box (multi.cu) --------------
 * One rank per GPU; the ranks of a run live in ONE process (n_local == world: the plugin's case, physim is
 * one process driving G GPUs) or in one process each (n_local == 1: bench.py under torchrun).  Barnes-Hut:
 * tree build and walk sharded by Morton key range, level-K cell records and accelerations exchanged with
 * ncclAllGather (NCCL is dlopen'ed at the first multi-rank use), remote cells read through NVLink peer
 * memory; direct sum: targets sharded by body index.  The integrator state is replicated, so any rank
 * can hand the whole state back.  Replaces the single simulation thread of pipeline.rs:134-192 on the
 * transformers.rs:123-160 / verlet.rs:52-82 path; results are bit-identical to the single-GPU handle's. */
/* 128 bytes identifying one communicator; call on one process and hand the bytes to every other one
 * (processes: any channel, e.g. a torch.distributed broadcast).  Not needed when n_local == world. */
int pb200_comm_unique_id(uint8_t *out128);
/* local_ranks[i] runs on CUDA device devices[i] (i < n_local).  Collective over all `world` ranks when world > 1. */
void *pb200_msim_create(int kind, double theta, double e, double dt, int world, int n_local,
                        const int *local_ranks, const int *devices, const uint8_t *nccl_id128);
void pb200_msim_destroy(void *msim);
/* every process passes the whole state (all n bodies) */
int pb200_msim_upload(void *msim, const Entity *state, size_t n);
int pb200_msim_generate_cube(void *msim, size_t n, uint64_t seed, double spin, double mass, double size,
                             const double *centre3);
int pb200_msim_run(void *msim, size_t steps);
/* *ms = device time of the steps: CUDA events on every local rank's stream, the maximum over them */
int pb200_msim_run_timed(void *msim, size_t steps, float *ms);
int pb200_msim_download(void *msim, Entity *state, size_t n);
int pb200_msim_last_accelerations(void *msim, Acceleration *acc, size_t n);
int pb200_msim_stats(void *msim, Pb200Stats *out, uint64_t *sharded_steps, uint64_t *replicated_steps);
/* bodies / cells per rank at the last sharded step; returns world, or -1 when no sharded step has run */
int pb200_msim_rank_counts(void *msim, uint32_t *bodies, uint32_t *cells);
/* 0: the local ranks' copies of the replicated state are bit-identical, 1: they differ, -1: error */
int pb200_msim_replicas_identical(void *msim);
/* diagnostics: out[0..8] key cuts of local rank 0's next sharded build, [9] bodies its last one kept, [10] epoch,
   [11] per-rank capacity in bodies, [12] "a wait for a peer gave up" */
int pb200_msim_debug_shard(void *msim, uint64_t *out13);
int pb200_msim_profile(void *msim, int enable);
int pb200_msim_profile_report(void *msim, char *buf, size_t cap);

/* --- csvsink (utilities/src/csvsink.rs:44-80) ---------------------------
 * The headless renderer of a CPU run: one line per printed state, "x,y,z," per entity, numbers in
 * Rust's `{}` format for f64 (shortest round-trip digits, positional notation).  The first state
 * pushed (the initial one, pipeline.rs:129-131) is always printed, state k >= 1 iff k % print_n == 0.
 * file NULL or "" selects the reference's default "csvsink.csv"; the file is created / truncated. */
void *pb200_csvsink_create(const char *file, size_t print_n);
int pb200_csvsink_push(void *sink, const Entity *state, size_t n);
void pb200_csvsink_destroy(void *sink);
/* States received so far; whether the next one will be printed; count a state without printing it
 * (a producer that knows the state is not due can skip the copy back from the device). */
size_t pb200_csvsink_count(void *sink);
int pb200_csvsink_next_is_printed(void *sink);
int pb200_csvsink_skip(void *sink);
/* Rust `format!("{}", v)` for an f64 into buf (no NUL); returns the length (<= cap; 400 always fits). */
size_t pb200_csv_format_f64(double v, char *buf, size_t cap);
/* Device-resident loop feeding a csvsink the way the pipeline feeds its renderer: the current state
 * if the sink is fresh, then the state after each of `steps` steps (one D2H of the positions per
 * printed state). */
int pb200_sim_run_csvsink(void *sim, size_t steps, void *sink);

/* --- microbenchmarks used by bench.py for the roofline denominators ---------------------- */
/* FP32 FFMA issue-rate probe: returns achieved TFLOP/s (2 flop per FFMA) on the current device. */
double pb200_probe_fp32_tflops(void);

#ifdef __cplusplus
}
#endif
#endif /* PHYSIM_B200_H */

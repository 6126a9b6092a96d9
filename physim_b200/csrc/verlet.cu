// verlet.cu — physim's `verlet` integrator update (integrators/src/verlet.rs:24-82) on sm_100a.
//
// HBM-bound: per body 32 B {x,y,z,m} + 32 B previous (or velocity) + 16/24 B acceleration read,
// 32 + 32 + 32 B written, every access a 128-bit load/store.  The update is done in place: the
// reference's `previous_state = entities.to_vec()` (verlet.rs:26,81) is the store to `prev`.
// Arithmetic is fp64 with explicit round-to-nearest intrinsics in the reference's operation order
// (no FMA contraction), so given the same accelerations the new state is bit-identical.
#include "common.cuh"

namespace pb200 {
namespace {

struct D3 {
  double x, y, z;
};

template <bool ACC64>
__device__ __forceinline__ D3 load_acc(const float4* __restrict__ a32, const double* __restrict__ a64,
                                       size_t i) {
  if (ACC64) return D3{a64[3 * i], a64[3 * i + 1], a64[3 * i + 2]};
  const float4 a = a32[i];
  return D3{double(a.x), double(a.y), double(a.z)};
}

// x1 = x0 + v0·dt + ½·a·dt² ; v1 = v0 + a·dt                               (verlet.rs:31-38)
__device__ __forceinline__ double first_pos(double x, double v, double a, double dt, double dt2) {
  return __dadd_rn(__dadd_rn(x, __dmul_rn(v, dt)), __dmul_rn(__dmul_rn(0.5, a), dt2));
}
// x' = 2·x − x_prev + a·dt²                                                 (verlet.rs:64-66)
__device__ __forceinline__ double next_pos(double x, double p, double a, double dt2) {
  return __dadd_rn(__dsub_rn(__dmul_rn(2.0, x), p), __dmul_rn(a, dt2));
}

template <bool FIRST, bool ACC64>
__global__ void __launch_bounds__(256) verlet_kernel(double4* __restrict__ cur, double4* __restrict__ prev,
                                                     double4* __restrict__ vel,
                                                     const float4* __restrict__ acc32,
                                                     const double* __restrict__ acc64, size_t n,
                                                     double dt, double dt2, double2* __restrict__ out6) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double4 x = cur[i];
  const D3 a = load_acc<ACC64>(acc32, acc64, i);
  double4 nx, nv;
  if (FIRST) {
    const double4 v = vel[i];
    nx.x = first_pos(x.x, v.x, a.x, dt, dt2);
    nx.y = first_pos(x.y, v.y, a.y, dt, dt2);
    nx.z = first_pos(x.z, v.z, a.z, dt, dt2);
    nv.x = __dadd_rn(v.x, __dmul_rn(a.x, dt));
    nv.y = __dadd_rn(v.y, __dmul_rn(a.y, dt));
    nv.z = __dadd_rn(v.z, __dmul_rn(a.z, dt));
  } else {
    const double4 p = prev[i];
    nx.x = next_pos(x.x, p.x, a.x, dt2);
    nx.y = next_pos(x.y, p.y, a.y, dt2);
    nx.z = next_pos(x.z, p.z, a.z, dt2);
    nv.x = __ddiv_rn(__dsub_rn(nx.x, x.x), dt);  // (x' − x)/dt                (verlet.rs:68-70)
    nv.y = __ddiv_rn(__dsub_rn(nx.y, x.y), dt);
    nv.z = __ddiv_rn(__dsub_rn(nx.z, x.z), dt);
  }
  nx.w = x.w;  // mass passes through
  nv.w = 0.0;
  prev[i] = x;
  cur[i] = nx;
  vel[i] = nv;
  if (out6) {
    if (FIRST) {  // packed {x,y,z,vx,vy,vz} for the D2H copy at the host boundary (48 B/body)
      out6[3 * i] = make_double2(nx.x, nx.y);
      out6[3 * i + 1] = make_double2(nx.z, nv.x);
      out6[3 * i + 2] = make_double2(nv.y, nv.z);
    } else {
      // regular step: {x,y,z} only (24 B/body).  The velocity is (x' - x)/dt with x the caller's own input
      // (verlet.rs:68-70): the host derives it from what it already holds - the same IEEE subtraction and
      // division on the same operands, hence the same bits - and half the D2H bytes stay off the bus.
      double* o = reinterpret_cast<double*>(out6) + 3 * i;
      o[0] = nx.x; o[1] = nx.y; o[2] = nx.z;
    }
  }
}

// Device-resident form of the regular step (verlet.rs:52-82) for the loop that keeps its state in HBM:
// 32 + 32 + 16 B read, 32 B written per body (the general kernel above: 80 read, 96 written).  x_{n+1}
// replaces x_{n-1} in place and the caller swaps the buffers; the velocity (x_{n+1} - x_n) / dt is not
// stored - verlet_velocity_kernel derives it, bit for bit the same, when somebody asks for it.  The
// kernel also leaves max(|x|,|y|,|z|) of the new positions for the next tree build (transformers.rs:35-40).
__global__ void __launch_bounds__(256) verlet_lean_kernel(const double4* __restrict__ cur,
                                                          double4* __restrict__ prev_inout,
                                                          const float4* __restrict__ acc32, size_t acc_stride,
                                                          size_t n, double dt2,
                                                          unsigned long long* __restrict__ extent_out,
                                                          unsigned long long* __restrict__ extent_zero,
                                                          unsigned long long* __restrict__ extent_last) {
  pb_pdl_sync();
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i == 0) {  // the slot the build of this step consumed: kept for the statistics, then cleared for re-use
    *extent_last = *extent_zero;
    *extent_zero = 0ull;
  }
  double m = 0.0;
  if (i < n) {
    const double4 x = cur[i];
    const double4 p = prev_inout[i];
    const float4 a = acc32[i * acc_stride];
    double4 nx;
    nx.x = next_pos(x.x, p.x, double(a.x), dt2);
    nx.y = next_pos(x.y, p.y, double(a.y), dt2);
    nx.z = next_pos(x.z, p.z, double(a.z), dt2);
    nx.w = x.w;  // mass passes through
    prev_inout[i] = nx;
    m = fmax(fmax(fabs(nx.x), fabs(nx.y)), fabs(nx.z));  // fmax drops NaN operands, like Rust's f64::max
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = s[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) v = fmax(v, s[w]);
    // non-negative doubles order like their bit patterns
    if (v > 0.0) atomicMax(extent_out, static_cast<unsigned long long>(__double_as_longlong(v)));
  }
}

__global__ void __launch_bounds__(256) verlet_velocity_kernel(const double4* __restrict__ cur,
                                                              const double4* __restrict__ prev,
                                                              double4* __restrict__ vel, size_t n, double dt) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double4 x = cur[i], p = prev[i];
  vel[i] = make_double4(__ddiv_rn(__dsub_rn(x.x, p.x), dt), __ddiv_rn(__dsub_rn(x.y, p.y), dt),
                        __ddiv_rn(__dsub_rn(x.z, p.z), dt), 0.0);  // (x' - x)/dt     (verlet.rs:68-70)
}

// ---- resident host boundary: whole Entity records on the device --------------------------------------
// ent80 holds the caller's Entity array (80-byte records, physim-core/src/lib.rs:16-30).  split: the state the
// kernels work on ({x,y,z,m}, {vx,vy,vz,0}, fixed flags) out of it; merge: new positions / velocities back into
// it, the other fields (radius, mass, id, fixed) untouched - so that one DMA of ent80 IS new_state
// (verlet.rs:41-48 / :72-79 copy the entity and overwrite six fields).
__global__ void __launch_bounds__(256) entity_split_kernel(const double2* __restrict__ ent80, size_t n,
                                                           double4* __restrict__ pos, double4* __restrict__ vel,
                                                           uint8_t* __restrict__ fixed) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double2* e = ent80 + 5 * i;  // {x,y} {z,vx} {vy,vz} {radius,mass} {id,fixed}
  const double2 a = e[0], b = e[1], c = e[2], d = e[3], t = e[4];
  pos[i] = make_double4(a.x, a.y, b.x, d.y);
  vel[i] = make_double4(b.y, c.x, c.y, 0.0);
  fixed[i] = (__double_as_longlong(t.y) & 0xff) ? 1 : 0;
}

__global__ void __launch_bounds__(256) entity_merge_kernel(double2* __restrict__ ent80, size_t n,
                                                           const double4* __restrict__ pos,
                                                           const double4* __restrict__ vel) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double4 p = pos[i], v = vel[i];
  double2* e = ent80 + 5 * i;
  e[0] = make_double2(p.x, p.y);
  e[1] = make_double2(p.z, v.x);
  e[2] = make_double2(v.y, v.z);
}

// ---- rk4 (integrators/src/rk4.rs:23-183) -----------------------------------------------------
// One kernel per stage.  k = (dt * v_at, dt * a) is folded into the running sums
// S = ((k1 + 2 k2) + 2 k3) + k4 in the reference's left-to-right order, and the next evaluation
// point e + w k (w = 0.5, 0.5, then plain e + k) is written for the next force evaluation.  The
// last stage writes the new state e + S / 6 instead (fixed bodies: position kept, velocity zero).
template <int STAGE, bool ACC64>
__global__ void __launch_bounds__(256) rk4_stage_kernel(const double4* __restrict__ e_pos,
                                                        const double4* __restrict__ e_vel,
                                                        const uint8_t* __restrict__ fixed,
                                                        double4* __restrict__ t_pos, double4* __restrict__ t_vel,
                                                        double4* __restrict__ s_pos, double4* __restrict__ s_vel,
                                                        const float4* __restrict__ acc32,
                                                        const double* __restrict__ acc64, size_t n, double dt,
                                                        double4* __restrict__ out_pos, double4* __restrict__ out_vel,
                                                        double2* __restrict__ out6) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double4 ep = e_pos[i], ev = e_vel[i];
  const double4 at = STAGE == 1 ? ev : t_vel[i];  // velocity the accelerations were evaluated at
  const D3 a = load_acc<ACC64>(acc32, acc64, i);
  const double kx = __dmul_rn(dt, at.x), ky = __dmul_rn(dt, at.y), kz = __dmul_rn(dt, at.z);
  const double kvx = __dmul_rn(dt, a.x), kvy = __dmul_rn(dt, a.y), kvz = __dmul_rn(dt, a.z);
  double4 sp, sv;
  if (STAGE == 1) {
    sp = make_double4(kx, ky, kz, 0.0);
    sv = make_double4(kvx, kvy, kvz, 0.0);
  } else {
    sp = s_pos[i];
    sv = s_vel[i];
    const double w = STAGE == 4 ? 1.0 : 2.0;  // 2.0 * k2, 2.0 * k3, k4
    if (STAGE == 4) {
      sp.x = __dadd_rn(sp.x, kx); sp.y = __dadd_rn(sp.y, ky); sp.z = __dadd_rn(sp.z, kz);
      sv.x = __dadd_rn(sv.x, kvx); sv.y = __dadd_rn(sv.y, kvy); sv.z = __dadd_rn(sv.z, kvz);
    } else {
      sp.x = __dadd_rn(sp.x, __dmul_rn(w, kx)); sp.y = __dadd_rn(sp.y, __dmul_rn(w, ky));
      sp.z = __dadd_rn(sp.z, __dmul_rn(w, kz));
      sv.x = __dadd_rn(sv.x, __dmul_rn(w, kvx)); sv.y = __dadd_rn(sv.y, __dmul_rn(w, kvy));
      sv.z = __dadd_rn(sv.z, __dmul_rn(w, kvz));
    }
  }
  if (STAGE < 4) {
    s_pos[i] = sp;
    s_vel[i] = sv;
    double4 tp, tv;
    if (STAGE == 3) {  // e + k3
      tp = make_double4(__dadd_rn(ep.x, kx), __dadd_rn(ep.y, ky), __dadd_rn(ep.z, kz), ep.w);
      tv = make_double4(__dadd_rn(ev.x, kvx), __dadd_rn(ev.y, kvy), __dadd_rn(ev.z, kvz), 0.0);
    } else {  // e + 0.5 * k
      tp = make_double4(__dadd_rn(ep.x, __dmul_rn(0.5, kx)), __dadd_rn(ep.y, __dmul_rn(0.5, ky)),
                        __dadd_rn(ep.z, __dmul_rn(0.5, kz)), ep.w);
      tv = make_double4(__dadd_rn(ev.x, __dmul_rn(0.5, kvx)), __dadd_rn(ev.y, __dmul_rn(0.5, kvy)),
                        __dadd_rn(ev.z, __dmul_rn(0.5, kvz)), 0.0);
    }
    t_pos[i] = tp;
    t_vel[i] = tv;
    if (out6) {  // the generic (host-callback) form needs the evaluation point on the host
      out6[3 * i] = make_double2(tp.x, tp.y);
      out6[3 * i + 1] = make_double2(tp.z, tv.x);
      out6[3 * i + 2] = make_double2(tv.y, tv.z);
    }
  } else {
    double4 np, nv;
    if (!fixed[i]) {  // rk4.rs:155-162
      np = make_double4(__dadd_rn(ep.x, __ddiv_rn(sp.x, 6.0)), __dadd_rn(ep.y, __ddiv_rn(sp.y, 6.0)),
                        __dadd_rn(ep.z, __ddiv_rn(sp.z, 6.0)), ep.w);
      nv = make_double4(__dadd_rn(ev.x, __ddiv_rn(sv.x, 6.0)), __dadd_rn(ev.y, __ddiv_rn(sv.y, 6.0)),
                        __dadd_rn(ev.z, __ddiv_rn(sv.z, 6.0)), 0.0);
    } else {          // rk4.rs:163-170
      np = ep;
      nv = make_double4(0.0, 0.0, 0.0, 0.0);
    }
    out_pos[i] = np;
    out_vel[i] = nv;
    if (out6) {
      out6[3 * i] = make_double2(np.x, np.y);
      out6[3 * i + 1] = make_double2(np.z, nv.x);
      out6[3 * i + 2] = make_double2(nv.y, nv.z);
    }
  }
}

}  // namespace

cudaError_t rk4_stage(int stage, const double4* e_pos, const double4* e_vel, const uint8_t* fixed,
                      double4* t_pos, double4* t_vel, double4* s_pos, double4* s_vel, const float4* acc32,
                      const double* acc64, size_t n, double dt, double4* out_pos, double4* out_vel,
                      double* out6, cudaStream_t st, LaunchStats& ls) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  double2* o6 = reinterpret_cast<double2*>(out6);
#define PB_RK4(S, A)                                                                                      \
  PB_LAUNCH(ls, st, "rk4_stage_kernel",                                                                   \
            (rk4_stage_kernel<S, A><<<blocks, 256, 0, st>>>(e_pos, e_vel, fixed, t_pos, t_vel, s_pos, s_vel, \
                                                            acc32, acc64, n, dt, out_pos, out_vel, o6)))
  if (acc64) {
    if (stage == 1) PB_RK4(1, true); else if (stage == 2) PB_RK4(2, true);
    else if (stage == 3) PB_RK4(3, true); else PB_RK4(4, true);
  } else {
    if (stage == 1) PB_RK4(1, false); else if (stage == 2) PB_RK4(2, false);
    else if (stage == 3) PB_RK4(3, false); else PB_RK4(4, false);
  }
#undef PB_RK4
  return cudaGetLastError();
}

cudaError_t verlet_update_lean(const double4* cur, double4* prev_inout, const float4* acc32, size_t n, double dt,
                               unsigned long long* extent_out, unsigned long long* extent_zero,
                               unsigned long long* extent_last, cudaStream_t st, LaunchStats& ls, size_t acc_stride) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  const double dt2 = dt * dt;  // dt.powi(2)
  PB_LAUNCH(ls, st, "verlet_lean_kernel",
            pb_launch_pdl(verlet_lean_kernel, dim3(blocks), dim3(256), 0, st, cur, prev_inout, acc32, acc_stride, n, dt2, extent_out, extent_zero, extent_last));
  return cudaGetLastError();
}

cudaError_t entity_split(const void* ent80, size_t n, double4* pos, double4* vel, uint8_t* fixed, cudaStream_t st,
                         LaunchStats& ls) {
  if (n == 0) return cudaSuccess;
  PB_LAUNCH(ls, st, "entity_split_kernel",
            entity_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
                static_cast<const double2*>(ent80), n, pos, vel, fixed));
  return cudaGetLastError();
}

cudaError_t entity_merge(void* ent80, size_t n, const double4* pos, const double4* vel, cudaStream_t st, LaunchStats& ls) {
  if (n == 0) return cudaSuccess;
  PB_LAUNCH(ls, st, "entity_merge_kernel",
            entity_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(static_cast<double2*>(ent80), n,
                                                                                      pos, vel));
  return cudaGetLastError();
}

cudaError_t verlet_velocity(const double4* cur, const double4* prev, double4* vel, size_t n, double dt,
                            cudaStream_t st, LaunchStats& ls) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  PB_LAUNCH(ls, st, "verlet_velocity_kernel", verlet_velocity_kernel<<<blocks, 256, 0, st>>>(cur, prev, vel, n, dt));
  return cudaGetLastError();
}

cudaError_t verlet_update(double4* cur, double4* prev, double4* vel, const float4* acc32,
                          const double* acc64, size_t n, double dt, int first, cudaStream_t st,
                          LaunchStats& ls, double* out6) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  const double dt2 = dt * dt;  // dt.powi(2)
  if (acc64) {
    if (first)
      PB_LAUNCH(ls, st, "verlet_kernel", verlet_kernel<true, true><<<blocks, 256, 0, st>>>(cur, prev, vel, acc32, acc64, n, dt, dt2, reinterpret_cast<double2*>(out6)));
    else
      PB_LAUNCH(ls, st, "verlet_kernel", verlet_kernel<false, true><<<blocks, 256, 0, st>>>(cur, prev, vel, acc32, acc64, n, dt, dt2, reinterpret_cast<double2*>(out6)));
  } else {
    if (first)
      PB_LAUNCH(ls, st, "verlet_kernel", verlet_kernel<true, false><<<blocks, 256, 0, st>>>(cur, prev, vel, acc32, acc64, n, dt, dt2, reinterpret_cast<double2*>(out6)));
    else
      PB_LAUNCH(ls, st, "verlet_kernel", verlet_kernel<false, false><<<blocks, 256, 0, st>>>(cur, prev, vel, acc32, acc64, n, dt, dt2, reinterpret_cast<double2*>(out6)));
  }
  return cudaGetLastError();
}

}  // namespace pb200

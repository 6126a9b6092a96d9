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
  if (out6) {  // packed {x,y,z,vx,vy,vz} for the D2H copy at the host boundary (48 B/body)
    out6[3 * i] = make_double2(nx.x, nx.y);
    out6[3 * i + 1] = make_double2(nx.z, nv.x);
    out6[3 * i + 2] = make_double2(nv.y, nv.z);
  }
}

}  // namespace

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

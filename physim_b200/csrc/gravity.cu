// gravity.cu — force evaluation kernels of physim's astro / astro2 / simple_astro on sm_100a.
//
// Reference behaviour being reproduced (jhb123/physim v0.4.4):
//   astro/src/transformers.rs:32-69,123-160,220-244   orchestration, skip rules, acc += f/m
//   astro/src/octree.rs:59-215, quadtree.rs:59-194     tree topology, acceptance, octant rule
//   astro/src/lib.rs:84-113                            force law  m_b (p_b - p_a) / (|r| (r^2 + e))
// The pointer tree of the reference is replaced by: compare-and-halve keys (fp64, the reference's own
// octant arithmetic) -> stable sort (buckets cut at the previous step's key quantiles and sorted in shared
// memory, or global LSD radix passes) -> units and cell counts per sorted body (TMA-staged tiles) -> scan
// -> DFS pre-order cell table (cells_kernel) -> bottom-up centres of mass (kids_kernel, climb_kernel) ->
// stack-free walk, one lane per target.  The kernels of that chain are launched as programmatic
// dependents of one another (pb_launch_pdl / pb_pdl_sync).  oracle/physim_oracle.cpp ("oracle 2") builds
// the same table on the CPU; tests compare the two bit for bit.
#include <cuda/atomic>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "common.cuh"

namespace pb200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t NO_PARENT = 0xffffffffu;
constexpr int MERGE_RUN_CAP = 1024;  // longer chains of <1e-9 neighbours are left unmerged
constexpr unsigned char NSV_NONE = 255;  // entry of a body that heads no cell (merged into the unit before it)

// Number of bodies a tree kernel works on: a host value (single GPU: every body), or a value left on the
// device by the sort of the same build (sharded build: the bodies whose keys fall in this rank's key range
// are only counted on the device; grids are then sized for the capacity and surplus CTAs leave at once).
struct NRef {
  const uint32_t* dev;
  uint32_t host;
  __device__ __forceinline__ size_t get() const { return dev ? size_t(*dev) : size_t(host); }
};

// Where a rank's kernels store what the other ranks need (own rank included: local pointers).
struct PeerTargets {
  uint4* top_info[8];
  double4* top_com[8];
  uint32_t* meta[8];
  char* xacc[8];
  uint32_t* flags[8];
  int world, rank;
};

// Producer -> consumer hand-over between ranks.  A producer kernel (keys, level-K records, accelerations) only
// stores into the peers' buffers.  The kernel that FOLLOWS it on the same stream - launched in plain stream order,
// so the producer grid has completed and its stores are performed - begins with shard_signal_then_wait: one thread
// tells every rank "this rank's part of `epoch` is in place" (system fence, then the epoch into the peers' flag
// words), then the kernel waits until every rank has said so.  No fence or atomic in the producers.
__device__ __forceinline__ void shard_signal_then_wait(const PeerTargets& pt, int slot, uint32_t epoch, bool signaller) {
  if (signaller && threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < pt.world; ++r) *reinterpret_cast<volatile uint32_t*>(pt.flags[r] + slot + pt.rank) = epoch;
  }
  if (int(threadIdx.x) < pt.world) shard_wait_flag(pt.flags[pt.rank], slot + int(threadIdx.x), epoch);
  __syncthreads();
}

template <int DIM>
struct TreeDim {
  static constexpr int LM = (DIM == 3) ? 21 : 31;  // key levels: 63 / 62 bits
};

template <int DIM>
__device__ __forceinline__ int shared_levels(uint64_t a, uint64_t b) {
  const uint64_t x = a ^ b;
  if (x == 0) return TreeDim<DIM>::LM;
  const int top = 63 - __clzll(static_cast<long long>(x));
  return TreeDim<DIM>::LM - 1 - top / DIM;
}

// Lanes holding the same 8-bit digit (and ok): eight ballots instead of MATCH.ANY, which costs
// ~1 us per call with ~30 distinct values in the warp on sm_100.
__device__ __forceinline__ unsigned peers_of(unsigned d, bool ok) {
  unsigned peers = __ballot_sync(FULL, ok);
#pragma unroll
  for (int bit = 0; bit < 8; ++bit) {
    const bool one = (d >> bit) & 1u;
    const unsigned m = __ballot_sync(FULL, one);
    peers &= one ? m : ~m;
  }
  return peers;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Block-wide exclusive scan of one unsigned per thread (blockDim.x == 256). Returns the exclusive
// prefix; *total (optional, smem-broadcast) receives the block sum.
__device__ __forceinline__ unsigned block_exclusive_scan_256(unsigned v, unsigned* total) {
  __shared__ unsigned warp_sums[8];
  __shared__ unsigned block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < 8 ? warp_sums[lane] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned t = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < 8) warp_sums[lane] = wi - w;  // exclusive per warp
    if (lane == 7) block_total = wi;
  }
  __syncthreads();
  const unsigned r = warp_sums[warp] + incl - v;
  if (total) *total = block_total;
  __syncthreads();  // smem reusable by a following call
  return r;
}

// ---------------------------------------------------------------------------------------------
// K1  extent = max_i max(|x|,|y|,|z|)                        (transformers.rs:35-40 / :126-131)
// HBM: 32 B read per body.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) extent_kernel(const double4* __restrict__ pos, size_t n,
                                                     unsigned long long* __restrict__ out_bits) {
  pb_pdl_sync();
  double m = 0.0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x) {
    const double4 p = pos[i];
    m = fmax(m, fabs(p.x));  // fmax drops NaN operands, like Rust's f64::max
    m = fmax(m, fabs(p.y));
    m = fmax(m, fabs(p.z));
  }
  m = warp_max(m);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? s[threadIdx.x] : 0.0;
    v = warp_max(v);
    // non-negative doubles order like their bit patterns
    if (threadIdx.x == 0) atomicMax(out_bits, static_cast<unsigned long long>(__double_as_longlong(v)));
  }
}

// ---------------------------------------------------------------------------------------------
// K2  keys: one digit (x | y<<1 | z<<2) per level from the reference's strict compare and
//     centre ± extent/2 halving, in fp64, so every visited centre is the reference's double.
// HBM: 32 B read, 12 B written per body.
// ---------------------------------------------------------------------------------------------
// Sort passes over key bits [lo, lo + total): pass p covers `width(p)` bits starting at `shift(p)`,
// the first `rem` passes being one bit wider (plain arithmetic: no per-pass tables in the kernels).
struct SortPlan {
  int npass, lo, base, rem;
  __host__ __device__ int width(int p) const { return base + (p < rem ? 1 : 0); }
  __host__ __device__ int shift(int p) const { return lo + p * base + (p < rem ? p : rem); }
  __host__ __device__ unsigned mask(int p) const { return (1u << width(p)) - 1u; }
};

// pos > centre, then centre +/- half: one DSETP + one DADD per axis and level (the sign of `half`
// is set with an integer op on the high word instead of computing both candidates)
__device__ __forceinline__ double with_sign(double half, bool negative) {
  const int hi = __double2hiint(half) ^ (negative ? int(0x80000000u) : 0);
  return __hiloint2double(hi, __double2loint(half));
}

// ---- keys without the chain ------------------------------------------------------------------------
// The digit of level l says whether p > c_l, c_l being the centre the reference reaches by l rounded additions
// of +-extent/2^j.  c_l differs from the exact dyadic point c_l* = extent (k / 2^l) by at most l half-ulps of
// the extent (< 3.4e-15 extent over 31 levels), so for a body that is further than KEY_GUARD x extent from
// EVERY cell boundary of every level (they all lie on the finest grid) each decision equals the exact one, and
// the digits are the bits of floor((p + extent) 2^(LM-1) / extent) - three fp64 operations and a bit
// interleave per axis instead of LM dependent compare/add steps.  The quantised coordinate itself carries
// < 5e-16 x 2^LM grid units of rounding, far inside the guard.  Bodies inside the guard band of some boundary
// (2 x KEY_GUARD x 2^(LM-1) of them: 2e-4 of a quadtree's axes, 2e-7 of an octree's), NaNs, infinities and a zero or
// denormal extent take the chain (key_chain): the reference's operations, always right.
constexpr double KEY_GUARD = 1e-13;

template <int DIM>
__device__ __noinline__ uint64_t key_chain(double px, double py, double pz, double ext0) {
  double half = ext0, cx = 0.0, cy = 0.0, cz = 0.0;
  uint64_t k = 0;
#pragma unroll 1
  for (int l = 0; l < TreeDim<DIM>::LM; ++l) {
    half *= 0.5;  // == extent / 2.0 in IEEE arithmetic
    const bool bx = px > cx, by = py > cy;
    unsigned digit = unsigned(bx) | (unsigned(by) << 1);
    cx += with_sign(half, !bx);  // x + (-h) == x - h exactly
    cy += with_sign(half, !by);
    if (DIM == 3) {
      const bool bz = pz > cz;
      digit |= unsigned(bz) << 2;
      cz += with_sign(half, !bz);
    }
    k = (k << DIM) | digit;
  }
  return k;
}

template <int DIM>
__device__ __forceinline__ uint64_t spread_bits(uint64_t x) {
  if (DIM == 2) {  // 31 bits -> every second bit
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
  } else {  // 21 bits -> every third bit
    x = (x | (x << 32)) & 0x001f00000000ffffull;
    x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
  }
  return x;
}

// scale = 2^(LM-1) / extent (one division per thread, key_scale)
template <int DIM>
__device__ __forceinline__ double key_scale(double ext0) {
  return __ddiv_rn(double(1ull << (TreeDim<DIM>::LM - 1)), ext0);
}

template <int DIM>
__device__ __forceinline__ uint64_t key_of(double px, double py, double pz, double ext0, double scale) {
  constexpr int LM = TreeDim<DIM>::LM;
  constexpr double top = double(1ull << LM), guard = KEY_GUARD * double(1ull << (LM - 1));
  const double vx = (px + ext0) * scale, vy = (py + ext0) * scale, vz = DIM == 3 ? (pz + ext0) * scale : 0.5;
  const double fx = floor(vx), fy = floor(vy), fz = floor(vz);
  const double ex = vx - fx, ey = vy - fy, ez = vz - fz;  // exact (Sterbenz-like: same binade or smaller)
  // every comparison is written so that a NaN fails it
  const bool fast = vx > 0.0 && vx < top && vy > 0.0 && vy < top && vz > 0.0 && vz < top &&
                    ex >= guard && ex <= 1.0 - guard && ey >= guard && ey <= 1.0 - guard &&
                    (DIM != 3 || (ez >= guard && ez <= 1.0 - guard));
  if (!fast) return key_chain<DIM>(px, py, pz, ext0);
  uint64_t k = spread_bits<DIM>(static_cast<uint64_t>(static_cast<long long>(fx))) |
               (spread_bits<DIM>(static_cast<uint64_t>(static_cast<long long>(fy))) << 1);
  if (DIM == 3) k |= spread_bits<DIM>(static_cast<uint64_t>(static_cast<long long>(fz))) << 2;
  return k;
}

constexpr int SORT_HIST_SLOTS = 8;  // 8 sort passes at most (bucket sort: slot 0 holds the bucket cursors)

// Also accumulates the digit histograms of every sort pass (the keys are in registers anyway),
// warp-aggregated into shared memory, flushed once per block: gridDim.x is kept small.
template <int DIM>
__global__ void __launch_bounds__(256) encode_kernel(const double4* __restrict__ pos, size_t n,
                                                     const unsigned long long* __restrict__ extent_bits,
                                                     uint64_t* __restrict__ key,
                                                     uint32_t* __restrict__ idx, SortPlan plan,
                                                     unsigned* __restrict__ ghist) {
  __shared__ unsigned h[SORT_HIST_SLOTS * 256];
  for (int j = threadIdx.x; j < SORT_HIST_SLOTS * 256; j += 256) h[j] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double ext0 = __longlong_as_double(static_cast<long long>(*extent_bits));
  const double scale = key_scale<DIM>(ext0);
  const size_t stride = size_t(gridDim.x) * 256;
  // whole warps iterate together (the histogram step is warp-collective)
  for (size_t base = size_t(blockIdx.x) * 256 + (threadIdx.x & ~31); base < n; base += stride) {
    const size_t i = base + lane;
    const bool ok = i < n;
    uint64_t k = 0;
    if (ok) {
      const double4 p = pos[i];
      k = key_of<DIM>(p.x, p.y, p.z, ext0, scale);
      key[i] = k;
      idx[i] = static_cast<uint32_t>(i);
    }
    const unsigned okmask = __ballot_sync(FULL, ok);
    for (int p = 0; p < plan.npass; ++p) {
      const unsigned d = unsigned(k >> plan.shift(p)) & plan.mask(p);
      // the high digits are the same for the whole warp: one add; otherwise per-lane adds
      const unsigned d0 = __shfl_sync(FULL, d, 0);
      if (__all_sync(FULL, !ok || d == d0)) {
        if (lane == 0) atomicAdd(&h[p * 256 + d0], unsigned(__popc(okmask)));
      } else if (ok) {
        atomicAdd(&h[p * 256 + d], 1u);
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < SORT_HIST_SLOTS * 256; j += 256)
    if (h[j]) atomicAdd(&ghist[j], h[j]);
}

// ---------------------------------------------------------------------------------------------
// K3  LSD radix sort, <= 8-bit digits, stable, (u64 key, u32 value) pairs, "onesweep" form:
//     one histogram kernel for every pass up front, then ONE kernel per pass that ranks a tile,
//     obtains its global offsets by decoupled look-back over the preceding tiles, and scatters.
//     Only the key bits [lo, key_bits) are sorted (lo > 0 when the previous evaluation showed
//     that the tree is shallower than the key; validated after the build, see gravity_evaluate).
// HBM per body: 8 B (histograms) + passes x (12 B + 12 B).
// ---------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;

constexpr int SORT_MAX_PASSES = 8;
constexpr unsigned LB_INCL = 0x80000000u, LB_PART = 0x40000000u, LB_MASK = 0x3fffffffu;
constexpr unsigned LB_SPIN_LIMIT = 1u << 24;
constexpr int LB_BATCH = 16;

template <int SORT_ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, SORT_ITEMS >= 16 ? 2 : 4) sort_onesweep_pass(
    const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
    uint32_t* __restrict__ vout, size_t n, int shift, unsigned mask,
    const unsigned* __restrict__ ghist /*[256] of this pass*/, unsigned* status /*[tiles][256]*/,
    unsigned* tile_counter, unsigned* err_flag) {
  __shared__ unsigned whist[8][256];
  __shared__ unsigned tstart[256];  // first tile-local slot of each digit
  __shared__ unsigned gadj[256];    // global slot of a digit's first element minus its tile-local slot
  __shared__ unsigned tile_s;
  extern __shared__ __align__(16) unsigned char sort_smem[];
  constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
  uint64_t* ks = reinterpret_cast<uint64_t*>(sort_smem);            // tile keys, digit-sorted
  uint32_t* vs = reinterpret_cast<uint32_t*>(ks + SORT_TILE);       // tile values, digit-sorted
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) tile_s = atomicAdd(tile_counter, 1u);  // tiles are numbered in start order
  for (int j = tid; j < 8 * 256; j += SORT_THREADS) (&whist[0][0])[j] = 0;
  const unsigned dbase = block_exclusive_scan_256(ghist[tid], nullptr);  // start of each digit; syncs
  const unsigned tile = tile_s;
  const size_t tile_base = size_t(tile) * SORT_TILE;
  const size_t base = tile_base + size_t(warp) * 32 * SORT_ITEMS;
  const unsigned nvalid = static_cast<unsigned>(min(size_t(SORT_TILE), n - tile_base));

  uint64_t k[SORT_ITEMS];
  uint32_t v[SORT_ITEMS];
  unsigned rank[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const size_t g = base + size_t(i) * 32 + lane;
    const bool ok = g < n;
    k[i] = ok ? kin[g] : ~0ull;
    v[i] = ok ? vin[g] : 0u;
  }
  // stable rank inside the warp's contiguous segment: items in (i, lane) order == memory order
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const bool ok = (base + size_t(i) * 32 + lane) < n;
    const unsigned d = unsigned((k[i] >> shift) & mask);
    const unsigned peers = peers_of(d, ok);
    const unsigned before = ok ? whist[warp][d] : 0u;
    __syncwarp();
    const unsigned r = __popc(peers & ((1u << lane) - 1u));
    if (ok && r == 0u) whist[warp][d] = before + unsigned(__popc(peers));
    __syncwarp();
    rank[i] = before + r;
  }
  __syncthreads();
  // thread = digit: tile count, per-warp exclusive offsets, and the partial count published at once
  unsigned cnt = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const unsigned c = whist[w][tid];
    whist[w][tid] = cnt;  // exclusive offset of warp w inside the tile's run of this digit
    cnt += c;
  }
  // decoupled look-back: one 32-bit word per (tile, digit) carries flag + count
  volatile unsigned* st = status + size_t(tile) * 256 + tid;
  *st = cnt | (tile == 0 ? LB_INCL : LB_PART);
  const unsigned my_start = block_exclusive_scan_256(cnt, nullptr);  // syncs
  tstart[tid] = my_start;
  __syncthreads();
  // reorder the tile in shared memory (digit-sorted, stable), so that the global writes below are
  // coalesced: each digit's run is contiguous both here and at its destination
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const bool ok = (base + size_t(i) * 32 + lane) < n;
    if (ok) {
      const unsigned d = unsigned((k[i] >> shift) & mask);
      const unsigned slot = tstart[d] + whist[warp][d] + rank[i];
      ks[slot] = k[i];
      vs[slot] = v[i];
    }
  }
  {
    unsigned prefix = 0;
    if (tile != 0) {
      // walk back over the predecessors, LB_BATCH status words in flight at a time (the tiles of
      // one wave publish their partial counts at about the same moment, so the walk is long)
      unsigned t = tile;  // tiles t-1, t-2, ... remain to be examined
      unsigned spins = 0;
      bool done = false;
      while (!done) {
        unsigned vals[LB_BATCH];
#pragma unroll
        for (int j = 0; j < LB_BATCH; ++j)
          vals[j] = (t > unsigned(j))
                        ? *reinterpret_cast<volatile unsigned*>(status + size_t(t - 1 - j) * 256 + tid)
                        : unsigned(LB_INCL);  // before tile 0: inclusive prefix 0
        unsigned used = 0;
#pragma unroll
        for (int j = 0; j < LB_BATCH; ++j) {
          if (!done && used == unsigned(j)) {
            const unsigned val = vals[j];
            if ((val & (LB_INCL | LB_PART)) != 0u) {
              prefix += val & LB_MASK;
              ++used;
              if (val & LB_INCL) done = true;
            }
          }
        }
        t -= used;
        if (used == 0u && ++spins > LB_SPIN_LIMIT) {  // never expected; avoids an unbounded hang
          atomicExch(err_flag, 1u);
          break;
        }
      }
      *st = ((prefix + cnt) & LB_MASK) | LB_INCL;
    }
    gadj[tid] = dbase + prefix - my_start;
  }
  __syncthreads();
  for (unsigned slot = tid; slot < nvalid; slot += SORT_THREADS) {
    const uint64_t key = ks[slot];
    const unsigned d = unsigned((key >> shift) & mask);
    const unsigned dst = gadj[d] + slot;
    kout[dst] = key;
    vout[dst] = vs[slot];
  }
}

// ---------------------------------------------------------------------------------------------
// K3b  bucket sort (replaces K2 + K3 + K4 from the second evaluation on, for N up to ~3.7 M).
//      The stable LSD sort from index order is the same permutation as ANY sort by the pair
//      (key >> lo, original index), so neither step below needs to be stable:
//      encode_bucket_kernel  computes the keys (as K2) and appends (key, index) to one of 256
//                            fixed-capacity buckets: bucket = number of splitters <= key >> lo, the
//                            splitters being the 1/256 quantiles of the PREVIOUS evaluation's sorted
//                            keys (splitter_block) - bodies move little per step, so the buckets
//                            stay balanced whatever the distribution.  Any non-decreasing splitters
//                            keep the result exact; stale ones only unbalance the buckets.
//                            (per-CTA counts in shared memory, one global atomic per CTA and bucket)
//      sort_local_kernel     CTA b sorts bucket b in shared memory: counting sort on <= 12 bits of
//                            (key >> lo) - bucket base, then every element ranks itself inside its
//                            (tiny) bin by (key >> lo, index) - a bin of more than LOCAL_SCAN_LIMIT
//                            bodies (a clump inside the bucket) is sorted by a warp or the whole CTA instead;
//                            keys, permutation and the gathered {x,y,z,m} records go straight to their
//                            final places.
//      A bucket above capacity leaves the build flagged `bad`; the host re-runs it with the next capacity
//      class (or, beyond the largest, with the global passes).
// HBM per body: 32 R + 12 W, then 12 R + 12 W + 32 R + 32 W.
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ unsigned block_exclusive_scan_nt(unsigned v, unsigned* wsum /*smem[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned w = lane < NT / 32 ? wsum[lane] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - w;  // exclusive per warp
  }
  __syncthreads();
  const unsigned r = wsum[warp] + incl - v;
  __syncthreads();  // wsum reusable
  return r;
}

constexpr int ENC_ITEMS = 4;           // bodies per thread in encode_bucket_kernel
constexpr unsigned MAX_BUCKETS = 1024;  // 256 buckets up to ~3.7 M bodies, 512 / 1024 above (x 16384 keys of shared memory each)
constexpr int LOCAL_BIN_BITS = 12;     // counting-sort bins inside a bucket
// Members of a bin rank themselves against each other (quadratic in the bin): ~1 body per bin on smooth data.
// A bucket whose bodies sit in a few clumps of its key range (a rotating cube leaves the corners of its bounding
// cell empty: key ranges with nothing in them) puts 50-400 in a bin, which still costs microseconds; only a
// bin of more than this many (a thousand bodies on one spot) hands the build to the global passes.
constexpr unsigned LOCAL_SCAN_LIMIT = 48;     // bins up to this size: every member ranks itself by scanning the bin
constexpr unsigned LOCAL_BIN_LIMIT = 1024;    // bins up to this size are sorted by one warp, larger ones by the whole CTA
constexpr unsigned LOCAL_BIG_BINS = 30;       // CTA-sorted bins a bucket may have (16 k keys / 1024 = 16 at most)
constexpr unsigned LOCAL_MID_BINS = 352;      // warp-sorted bins a bucket may have (16 k keys / 48 = 341 at most)
constexpr unsigned LOCAL_SKEWED = 0x7fffffffu;  // published as "largest bucket" when a bucket has more big bins than that

template <int DIM, unsigned NB /* buckets: 256, 512 or 1024 */>
__global__ void __launch_bounds__(256) encode_bucket_kernel(const double4* __restrict__ pos, size_t n,
                                                            const unsigned long long* __restrict__ extent_bits,
                                                            const uint64_t* __restrict__ splitters /*[NB]*/,
                                                            int lo, uint64_t* __restrict__ bkey,
                                                            uint32_t* __restrict__ bidx, unsigned cap,
                                                            unsigned* __restrict__ cursor /*[nb]*/,
                                                            const uint64_t* __restrict__ cuts /* sharded build: keep
                                                            cuts[0] <= key < cuts[1]; nullptr: every body */) {
  pb_pdl_sync();
  constexpr unsigned nb = NB;
  __shared__ unsigned cnt[NB];
  __shared__ unsigned gbase[NB];
  __shared__ uint64_t spl[NB];
  const int tid = threadIdx.x;
#pragma unroll
  for (unsigned j = tid; j < nb; j += 256) {
    cnt[j] = 0u;
    spl[j] = j ? (splitters[j] >> lo) : 0ull;
  }
  __syncthreads();
  const double ext0 = __longlong_as_double(static_cast<long long>(*extent_bits));
  const size_t tile = size_t(blockIdx.x) * (256 * ENC_ITEMS);
  const double scale = key_scale<DIM>(ext0);
  double4 p[ENC_ITEMS];
  uint64_t k[ENC_ITEMS];
#pragma unroll
  for (int e = 0; e < ENC_ITEMS; ++e) {  // all loads first
    const size_t i = tile + size_t(e) * 256 + tid;
    p[e] = i < n ? pos[i] : make_double4(0.0, 0.0, 0.0, 0.0);
  }
#pragma unroll
  for (int e = 0; e < ENC_ITEMS; ++e) k[e] = key_of<DIM>(p[e].x, p[e].y, p[e].z, ext0, scale);
  unsigned r[ENC_ITEMS], d[ENC_ITEMS];
  const uint64_t cut_lo = cuts ? cuts[0] : 0ull, cut_hi = cuts ? cuts[1] : ~0ull;
  bool keep[ENC_ITEMS];
#pragma unroll
  for (int e = 0; e < ENC_ITEMS; ++e) {
    // bucket = number of splitters <= key >> lo, minus one (spl[0] = 0): log2(nb)-step search, no divergence
    const uint64_t kk = k[e] >> lo;
    unsigned b = 0;
#pragma unroll
    for (unsigned step = nb >> 1; step > 0; step >>= 1)
      if (spl[b + step] <= kk) b += step;
    d[e] = b;
    const size_t i = tile + size_t(e) * 256 + tid;
    keep[e] = i < n && k[e] >= cut_lo && k[e] < cut_hi;
    r[e] = keep[e] ? atomicAdd(&cnt[b], 1u) : 0u;
  }
  __syncthreads();
  for (unsigned j = tid; j < nb; j += 256) {
    const unsigned c = cnt[j];
    if (c) gbase[j] = atomicAdd(&cursor[j], c);
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < ENC_ITEMS; ++e) {
    const size_t i = tile + size_t(e) * 256 + tid;
    if (keep[e]) {
      const unsigned slot = gbase[d[e]] + r[e];
      if (slot < cap) {  // else: bucket over capacity, sort_local_kernel flags the build
        bkey[size_t(d[e]) * cap + slot] = k[e];
        bidx[size_t(d[e]) * cap + slot] = static_cast<uint32_t>(i);
      }
    }
  }
}

inline size_t sort_local_smem(unsigned cap) {
  // ... + seg[4] + key range (2 x u64) + the list of big bins
  return size_t(cap) * 12 + (size_t(1) << LOCAL_BIN_BITS) * 4 + 32 * 4 + 32 + (LOCAL_BIG_BINS + 2 + LOCAL_MID_BINS) * 4;
}

// The first RITEMS * NT elements of the bucket stay in registers between the counting and the
// scattering loop; anything beyond is read again (L2 hits).
template <int NT, int RITEMS>
__global__ void __launch_bounds__(NT, 1024 / NT) sort_local_kernel(
    const uint64_t* __restrict__ bkey, const uint32_t* __restrict__ bidx, const unsigned* __restrict__ cursor,
    const uint64_t* __restrict__ splitters, unsigned cap, int lo, int key_bits, uint64_t* __restrict__ keys,
    uint32_t* __restrict__ vals,
    const double4* __restrict__ pos, double4* __restrict__ spos, unsigned* __restrict__ bad,
    unsigned* __restrict__ stat_max, uint32_t* __restrict__ n_out /* sharded build: bodies sorted in all, capacity */,
    uint32_t n_cap, unsigned nb /* buckets = gridDim.x <= NT */) {
  pb_pdl_sync();
  static_assert(NT >= 256 && NT % 32 == 0 && (1 << LOCAL_BIN_BITS) % NT == 0, "scan layout");
  constexpr int NBINS = 1 << LOCAL_BIN_BITS, BPT = NBINS / NT;  // bins per thread in the scan
  extern __shared__ __align__(16) unsigned char sort_smem[];
  uint64_t* ks = reinterpret_cast<uint64_t*>(sort_smem);
  uint32_t* vs = reinterpret_cast<uint32_t*>(ks + cap);
  unsigned* bins = vs + cap;     // [NBINS]: counts -> running ends
  unsigned* wsum = bins + NBINS; // [32]
  unsigned* seg = wsum + 32;     // start, count of this bucket, largest bin, (sharded build) total of all buckets
  const int tid = threadIdx.x;
  {
    const unsigned c = unsigned(tid) < nb ? cursor[tid] : 0u;
    const unsigned ex = block_exclusive_scan_nt<NT>(c, wsum);
    if (tid == int(blockIdx.x)) { seg[0] = ex; seg[1] = c; }
    if (tid == 0) seg[2] = 0u;
    if (n_out && unsigned(tid) == nb - 1u) {
      seg[3] = ex + c;
      if (blockIdx.x == 0) {
        *n_out = min(ex + c, n_cap);  // (the later kernels never index past the capacity)
        if (ex + c > n_cap) {  // this rank's key range holds more bodies than its arrays: re-planned by the host
          *bad = 1u;
          stat_max[1] = 1u;    // sticky word 4
        }
      }
    }
    __syncthreads();
    if (n_out && seg[3] > n_cap) return;
  }
  const unsigned start = seg[0], cnt = seg[1];
  if (cnt == 0u) return;
  if (tid == 0) atomicMax(stat_max, cnt);  // worst case since the last host check (gravity_check)
  if (cnt > cap) {
    if (tid == 0) *bad = 1u;
    return;
  }
  for (int j = tid; j < NBINS; j += NT) bins[j] = 0u;
  const uint64_t* gk = bkey + size_t(blockIdx.x) * cap;
  const uint32_t* gv = bidx + size_t(blockIdx.x) * cap;
  uint64_t k[RITEMS];
  uint32_t v[RITEMS];
  uint64_t kmin = ~0ull, kmax = 0ull;
#pragma unroll
  for (int i = 0; i < RITEMS; ++i) {
    const unsigned p = unsigned(i) * NT + tid;
    const bool ok = p < cnt;
    k[i] = ok ? gk[p] : 0ull;
    v[i] = ok ? gv[p] : 0u;
    if (ok) { kmin = min(kmin, k[i]); kmax = max(kmax, k[i]); }
  }
  for (unsigned p = unsigned(RITEMS) * NT + tid; p < cnt; p += NT) {
    const uint64_t key = gk[p];
    kmin = min(kmin, key);
    kmax = max(kmax, key);
  }
  // bins: (key >> lo) - base, scaled so that the key range the bucket's bodies ACTUALLY span covers the bins
  // (not the range between its splitters: that may reach across key space nothing lives in, and the bodies
  // would crowd a few bins): any non-decreasing map keeps the result exact
  {
    unsigned long long* krange = reinterpret_cast<unsigned long long*>(seg + 4);  // [min, max], 8-byte aligned
    if (tid == 0) { krange[0] = ~0ull; krange[1] = 0ull; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(FULL, kmax, o));
    }
    if ((tid & 31) == 0) {
      atomicMin(&krange[0], static_cast<unsigned long long>(kmin));
      atomicMax(&krange[1], static_cast<unsigned long long>(kmax));
    }
    __syncthreads();
    kmin = krange[0];
    kmax = krange[1];
  }
  const uint64_t kbase = kmin >> lo;
  const uint64_t span = (kmax >> lo) - kbase + 1ull;
  const int span_bits = 64 - __clzll(static_cast<long long>(span - 1ull) | 1ll);
  const int bshift = span_bits > LOCAL_BIN_BITS ? span_bits - LOCAL_BIN_BITS : 0;
  constexpr uint64_t bmax = (1u << LOCAL_BIN_BITS) - 1u;
  auto bin_of = [&](uint64_t key) {
    const uint64_t rel = ((key >> lo) - kbase) >> bshift;
    return unsigned(rel < bmax ? rel : bmax);
  };
#pragma unroll
  for (int i = 0; i < RITEMS; ++i)
    if (unsigned(i) * NT + tid < cnt) atomicAdd(&bins[bin_of(k[i])], 1u);
  for (unsigned p = unsigned(RITEMS) * NT + tid; p < cnt; p += NT) atomicAdd(&bins[bin_of(gk[p])], 1u);
  __syncthreads();
  {  // exclusive scan of the bin counts (BPT consecutive bins per thread)
    unsigned c[BPT], sum = 0;
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      c[q] = bins[tid * BPT + q];
      sum += c[q];
    }
    unsigned run = block_exclusive_scan_nt<NT>(sum, wsum);  // syncs
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      bins[tid * BPT + q] = run;
      run += c[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RITEMS; ++i) {
    if (unsigned(i) * NT + tid < cnt) {
      const unsigned slot = atomicAdd(&bins[bin_of(k[i])], 1u);
      ks[slot] = k[i];
      vs[slot] = v[i];
    }
  }
  for (unsigned p = unsigned(RITEMS) * NT + tid; p < cnt; p += NT) {
    const uint64_t key = gk[p];
    const unsigned slot = atomicAdd(&bins[bin_of(key)], 1u);
    ks[slot] = key;
    vs[slot] = gv[p];
  }
  __syncthreads();
  // Crowded bins.  A bucket that holds a clump AND a few far bodies maps the clump onto a few of its bins (bodies
  // falling towards a heavy star do that to the cube within ~100 steps of the 1 M-body astro2 run).  Ranking by
  // scanning is quadratic in the bin, so only bins of <= LOCAL_SCAN_LIMIT bodies are ranked that way; a larger bin
  // is SORTED in place - by one warp up to LOCAL_BIN_LIMIT bodies (the warps share the bins out), by the whole CTA
  // above - with a bitonic network in its all-ascending form (first step of a merge against the mirrored partner),
  // which needs no padding: a partner index past the end stands for +infinity and never moves anything.  The order
  // is (key bits, index), the one the scan establishes; the members of a sorted bin then rank by position.
  unsigned* big = seg + 8;                      // [0] count, [1 ..] bin ids: CTA-sorted
  unsigned* mid = big + LOCAL_BIG_BINS + 2;     // [0] count, [1 ..] bin ids: warp-sorted
  if (tid == 0) big[0] = mid[0] = 0u;
  __syncthreads();
  for (unsigned b = tid; b < unsigned(NBINS); b += NT) {
    const unsigned sz = bins[b] - (b ? bins[b - 1u] : 0u);
    if (sz > LOCAL_BIN_LIMIT) {
      const unsigned at = atomicAdd(&big[0], 1u);
      if (at < LOCAL_BIG_BINS) big[1u + at] = b;
    } else if (sz > LOCAL_SCAN_LIMIT) {
      const unsigned at = atomicAdd(&mid[0], 1u);
      if (at < LOCAL_MID_BINS - 1u) mid[1u + at] = b;
    }
  }
  __syncthreads();
  const unsigned n_big = big[0], n_mid = mid[0];
  if (n_big > LOCAL_BIG_BINS || n_mid > LOCAL_MID_BINS - 1u) {  // (cannot happen within the capacity; kept as a guard)
    if (tid == 0) {
      *bad = 1u;
      atomicMax(stat_max, LOCAL_SKEWED);
    }
    return;
  }
  auto exchange = [&](unsigned s0, unsigned m, unsigned i, unsigned j) {  // ascending: the smaller (key bits, index) first
    if (j >= m) return;
    const uint64_t ki = ks[s0 + i], kj = ks[s0 + j];
    const uint32_t vi = vs[s0 + i], vj = vs[s0 + j];
    const uint64_t bi_ = ki >> lo, bj_ = kj >> lo;
    if (bi_ > bj_ || (bi_ == bj_ && vi > vj)) {
      ks[s0 + i] = kj; ks[s0 + j] = ki;
      vs[s0 + i] = vj; vs[s0 + j] = vi;
    }
  };
  for (unsigned bi = 0; bi < n_big; ++bi) {  // the whole CTA, one bin after the other
    const unsigned b = big[1u + bi];
    const unsigned s0 = b ? bins[b - 1u] : 0u, m = bins[b] - s0;
    unsigned P = 1u;
    while (P < m) P <<= 1;
    for (unsigned kk = 2u; kk <= P; kk <<= 1) {  // (powers of two: masks and shifts, no divisions)
      const unsigned hk = kk >> 1;
      for (unsigned t = tid; t < (P >> 1); t += NT) {
        const unsigned base = (t & ~(hk - 1u)) << 1, off = t & (hk - 1u);
        exchange(s0, m, base + off, base + kk - 1u - off);
      }
      __syncthreads();
      for (unsigned j2 = kk >> 2; j2 > 0u; j2 >>= 1) {
        for (unsigned t = tid; t < (P >> 1); t += NT) {
          const unsigned i = ((t & ~(j2 - 1u)) << 1) | (t & (j2 - 1u));
          exchange(s0, m, i, i + j2);
        }
        __syncthreads();
      }
    }
  }
  {  // one warp per bin
    const unsigned warp = unsigned(tid) >> 5, lane = unsigned(tid) & 31u;
    for (unsigned mi = warp; mi < n_mid; mi += NT / 32) {
      const unsigned b = mid[1u + mi];
      const unsigned s0 = b ? bins[b - 1u] : 0u, m = bins[b] - s0;
      unsigned P = 1u;
      while (P < m) P <<= 1;
      for (unsigned kk = 2u; kk <= P; kk <<= 1) {
        const unsigned hk = kk >> 1;
        for (unsigned t = lane; t < (P >> 1); t += 32u) {
          const unsigned base = (t & ~(hk - 1u)) << 1, off = t & (hk - 1u);
          exchange(s0, m, base + off, base + kk - 1u - off);
        }
        __syncwarp();
        for (unsigned j2 = kk >> 2; j2 > 0u; j2 >>= 1) {
          for (unsigned t = lane; t < (P >> 1); t += 32u) {
            const unsigned i = ((t & ~(j2 - 1u)) << 1) | (t & (j2 - 1u));
            exchange(s0, m, i, i + j2);
          }
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  // bins[b] is now the end of bin b (and the start of bin b+1)
  for (unsigned p = tid; p < cnt; p += NT) {
    const uint64_t key = ks[p];
    const uint32_t id = vs[p];
    const unsigned bin = bin_of(key);
    const unsigned s = bin ? bins[bin - 1u] : 0u, e = bins[bin];
    const uint64_t kme = key >> lo;
    unsigned rank = p - s;  // (a crowded bin: sorted above)
    if (e - s <= LOCAL_SCAN_LIMIT) {
      rank = 0;
      for (unsigned q = s; q < e; ++q) {
        const uint64_t kq = ks[q] >> lo;
        rank += (kq < kme || (kq == kme && vs[q] < id)) ? 1u : 0u;
      }
    }
    const size_t dst = size_t(start) + s + rank;
    keys[dst] = key;
    vals[dst] = id;
    spos[dst] = pos[id];
  }
}

// splitters for the NEXT evaluation's bucket sort: out[0] = smallest key, out[1..255] = the 1/256
// quantiles of the sorted keys, out[256] = largest key + 1; forced non-decreasing (whatever the
// state of `sorted`, e.g. after a build that was abandoned)
constexpr int SPLITTER_STRIDE = MAX_BUCKETS + 8;  // u64 words per splitter set (nb + 1 used)

__device__ __forceinline__ void splitter_block(const uint64_t* __restrict__ sorted, size_t n,
                                               uint64_t* __restrict__ out /*[nb + 1]*/, unsigned nb) {
  __shared__ uint64_t v[MAX_BUCKETS + 1];
  const unsigned t = threadIdx.x;
  for (unsigned j = t; j < nb; j += 256) v[j] = n ? sorted[(size_t(j) * n) / nb] : 0ull;
  if (t == 0) v[nb] = n ? sorted[n - 1] + 1ull : 0ull;
  __syncthreads();
  if (t == 0) {
    uint64_t run = 0;
    for (unsigned j = 0; j <= nb; ++j) {
      run = v[j] > run ? v[j] : run;
      v[j] = run;
    }
  }
  __syncthreads();
  for (unsigned j = t; j <= nb; j += 256) out[j] = v[j];
}

// second level of the range-minimum tables (see NsvTables below): block minima -> tables inside
// super-blocks of 256 blocks, super-block minima
__device__ __forceinline__ void nsv_level2_block(unsigned sb, uint8_t* __restrict__ t2, size_t b_pad, size_t nblocks,
                                                 uint8_t* __restrict__ t3) {
  __shared__ unsigned char tab[2][256 + 128];
  const int t = threadIdx.x;
  const size_t b = size_t(sb) * 256 + t;
  tab[0][t] = b < nblocks ? t2[b] : 255;  // NSV_NONE
  if (t < 128) tab[0][256 + t] = tab[1][256 + t] = 255;
  __syncthreads();
  int cur = 0;
#pragma unroll
  for (int k = 1; k <= 8; ++k) {
    const unsigned char m = min(tab[cur][t], tab[cur][t + (1 << (k - 1))]);
    tab[cur ^ 1][t] = m;
    if (b < b_pad) t2[size_t(k) * b_pad + b] = m;
    cur ^= 1;
    __syncthreads();
  }
  if (t == 0) t3[sb] = tab[cur][0];
}

// small jobs that ride along with the scan of the cell counts (each would otherwise be a launch of a
// few CTAs between two kernels of the build): the CTAs whose ticket is past the last scan tile do them
struct ScanSide {
  uint8_t* t2;                // CTAs [0, nsuper): nsv_level2_block (nsuper = super-blocks of the n bodies)
  size_t b_pad;               // stride of a level of t2 (allocation)
  uint8_t* t3;
  const uint64_t* sorted;     // CTA nsuper: splitter_block
  uint64_t* spl_out;
  unsigned nb;                // buckets the next evaluation's bucket sort will use
  const uint64_t* spl_keep;   // the set THIS build read (same nb), or nullptr; handed on unchanged when the build
  const unsigned* bad;        // ... was abandoned (a bucket over capacity: `sorted` is incomplete)
};

// ---------------------------------------------------------------------------------------------
// exclusive scan of u32[n] -> out[n+1] (out[n] = total): one kernel, decoupled look-back done by a
// warp (32 predecessor tiles per round; status word = flag << 32 | value)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;
constexpr unsigned long long SCAN_PART = 1ull << 32, SCAN_INCL = 2ull << 32;

template <bool SIDE>
__global__ void __launch_bounds__(256) scan_lookback_kernel(const uint32_t* __restrict__ in, NRef nref,
                                                            uint32_t* __restrict__ out,
                                                            unsigned long long* status,
                                                            unsigned* tile_counter, ScanSide side) {
  pb_pdl_sync();
  __shared__ unsigned tile_s, tile_prefix_s;
  if (threadIdx.x == 0) tile_s = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const unsigned tile = tile_s;
  const size_t n = nref.get();
  const unsigned tiles = max(1u, unsigned((n + SCAN_TILE - 1) / SCAN_TILE));  // (n = 0: one tile writes out[0] = 0)
  if (tile >= tiles) {  // (the last tickets: nothing of the scan waits for these CTAs)
    if (SIDE) {
      const unsigned job = tile - tiles;
      const size_t nblocks = (n + 255) / 256, nsuper = (nblocks + 255) / 256;
      if (job < nsuper) nsv_level2_block(job, side.t2, side.b_pad, nblocks, side.t3);
      else if (job == nsuper) {
        if (side.spl_keep && *side.bad) {
          for (unsigned j = threadIdx.x; j <= side.nb; j += 256) side.spl_out[j] = side.spl_keep[j];
        } else {
          splitter_block(side.sorted, n, side.spl_out, side.nb);
        }
      }
    }
    return;
  }
  const size_t base = size_t(tile) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS];
  unsigned sum = 0;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
      const uint4 w = reinterpret_cast<const uint4*>(in + base)[q];
      v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
  }
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) sum += v[i];
  unsigned total;
  const unsigned mine = block_exclusive_scan_256(sum, &total);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    volatile unsigned long long* st = status + tile;
    unsigned prefix = 0;
    if (tile == 0) {
      if (lane == 0) *st = SCAN_INCL | total;
    } else {
      if (lane == 0) *st = SCAN_PART | total;
      unsigned t = tile;  // tiles t-1, t-2, ... remain
      while (true) {
        const bool valid = t > unsigned(lane);
        unsigned long long w = SCAN_INCL;  // before tile 0: inclusive 0
        if (valid) {
          do {
            w = *reinterpret_cast<volatile unsigned long long*>(status + (t - 1 - lane));
          } while ((w >> 32) == 0ull);
        }
        const unsigned incl = __ballot_sync(FULL, (w >> 32) == 2ull);
        const int first = incl ? __ffs(incl) - 1 : 31;  // nearest predecessor with an inclusive value
        unsigned val = lane <= first ? unsigned(w) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULL, val, o);
        prefix += val;
        if (incl) break;
        t -= 32;
      }
      if (lane == 0) *st = SCAN_INCL | (unsigned long long)(prefix + total);
    }
    if (lane == 0) tile_prefix_s = prefix;
  }
  __syncthreads();
  unsigned run = tile_prefix_s + mine;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
      uint4 w;
      w.x = run; run += v[4 * q];
      w.y = run; run += v[4 * q + 1];
      w.z = run; run += v[4 * q + 2];
      w.w = run; run += v[4 * q + 3];
      reinterpret_cast<uint4*>(out + base)[q] = w;
    }
    if (base + SCAN_ITEMS == n) out[n] = run;
  } else {
    if (n == 0 && threadIdx.x == 0) out[0] = 0u;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      if (base + i < n) out[base + i] = run;
      run += v[i];
      if (base + i + 1 == n) out[n] = run;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4  gather sorted {x,y,z,m}.  HBM: 4 + 32 read, 32 written per body.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const double4* __restrict__ pos,
                                                     const uint32_t* __restrict__ perm, size_t n,
                                                     double4* __restrict__ spos) {
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (s < n) spos[s] = pos[perm[s]];
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) completing on an mbarrier ------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// K5  merged-leaf runs and the number of cells each sorted body heads.
//     a = levels shared with the previous unit, b = with the next; a unit heads the cells at
//     levels a+1 .. max(a,b)+1 (the last one is its leaf).  See oracle 2 for the rule's derivation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool close_1e9(const double4& u, const double4& v) {
  return fabs(u.x - v.x) < 1e-9 && fabs(u.y - v.y) < 1e-9 && fabs(u.z - v.z) < 1e-9;
}

// Persistent CTAs over tiles of 256 sorted bodies.  A tile's positions and keys (with one body of halo
// on either side) arrive in shared memory by two bulk copies (TMA) that complete on an mbarrier, and the
// copies of the CTA's next tile are in flight while this one is worked on: one element per thread with
// plain loads left every CTA waiting out a full DRAM latency before it could do anything (1.4 TB/s).
constexpr int UNIT_TILE = 256;
constexpr int UNIT_CTAS_PER_SM = 5;

template <int DIM>
__global__ void __launch_bounds__(256) unit_kernel(const uint64_t* __restrict__ key,
                                                   const double4* __restrict__ sp, NRef nref,
                                                   uchar2* __restrict__ ab,
                                                   uint32_t* __restrict__ cnt,
                                                   unsigned* __restrict__ max_shared_plus1,
                                                   int levels_sorted, uint8_t* __restrict__ nsv1,
                                                   size_t n_pad /* offset of the window minima: allocation */,
                                                   uint8_t* __restrict__ nsv2) {
  pb_pdl_sync();
  constexpr int LM = TreeDim<DIM>::LM;
  const size_t n = nref.get();
  const unsigned n_tiles = unsigned((n + UNIT_TILE - 1) / UNIT_TILE);
  // t_sp[b][i] = sp[s0 - 1 + i] (i = 0 .. 257), t_key[b][i] = key[s0 - 2 + i] (i = 0 .. 259): the windows
  // start at 16-byte aligned addresses and are a multiple of 16 bytes long, as bulk copies must be
  __shared__ __align__(128) double4 t_sp[2][UNIT_TILE + 2];
  __shared__ __align__(128) uint64_t t_key[2][UNIT_TILE + 4];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ unsigned char warp_min[8];
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](unsigned tile, unsigned buf) {  // (one thread)
    const size_t s0 = size_t(tile) * UNIT_TILE;
    const size_t p0 = s0 ? s0 - 1 : 0, p1 = min(n, s0 + UNIT_TILE + 1);                       // bodies [p0, p1)
    const size_t k0 = s0 ? s0 - 2 : 0, k1 = min((n + 1) & ~size_t(1), s0 + UNIT_TILE + 2);    // keys [k0, k1), even count
    // (key[n] may be read when n is odd: inside the allocation's slack, never used)
    const unsigned sp_bytes = unsigned(p1 - p0) * 32u, key_bytes = unsigned(k1 - k0) * 8u;
    mbar_expect_tx(&full[buf], sp_bytes + key_bytes);
    bulk_load(&t_sp[buf][p0 - (s0 - 1)], sp + p0, sp_bytes, &full[buf]);     // (s0 = 0: slot 1 on)
    bulk_load(&t_key[buf][k0 - (s0 - 2)], key + k0, key_bytes, &full[buf]);   // (s0 = 0: slot 2 on)
  };
  unsigned deepest = 0;  // 1 + deepest level shared by two neighbours with DIFFERENT keys
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0u);
  unsigned it = 0;
  for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
  const unsigned buf = it & 1u;
  // buffer buf ^ 1 was read in the previous iteration; the barrier at its end makes the refill safe
  if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1u);
  mbar_wait(&full[buf], (it >> 1) & 1u);
  const size_t s = size_t(tile) * UNIT_TILE + threadIdx.x;
  unsigned char a1 = NSV_NONE;  // 1 + levels shared with the previous unit (heads only)
  if (s < n) {
    const double4 me = t_sp[buf][threadIdx.x + 1];
    const uint64_t kme = t_key[buf][threadIdx.x + 2];
    const bool cp = s > 0 && close_1e9(t_sp[buf][threadIdx.x], me);
    const bool cn = s + 1 < n && close_1e9(me, t_sp[buf][threadIdx.x + 2]);
    const uint64_t kprev = t_key[buf][threadIdx.x + 1], knext = t_key[buf][threadIdx.x + 3];
    int a = s > 0 ? shared_levels<DIM>(kprev, kme) : -1;
    int b = s + 1 < n ? shared_levels<DIM>(kme, knext) : -1;
    if (s > 0 && kprev != kme) {
      deepest = max(deepest, unsigned(a + 1));
      // two different keys that agree on every sorted bit: their order is not the full sort's, the
      // arrays below would describe a broken tree -> every later kernel of this build bails out
      if (a >= levels_sorted) max_shared_plus1[2] = 1u;
    }
    bool head = true;
    if (cp || cn) {
      size_t r0 = s, r1 = s;
      int inner = LM + 1;
      bool capped = false;
      int steps = 0;
      while (r0 > 0 && close_1e9(sp[r0 - 1], sp[r0])) {
        inner = min(inner, shared_levels<DIM>(key[r0 - 1], key[r0]));
        --r0;
        if (++steps > MERGE_RUN_CAP) { capped = true; break; }
      }
      steps = 0;
      while (!capped && r1 + 1 < n && close_1e9(sp[r1], sp[r1 + 1])) {
        inner = min(inner, shared_levels<DIM>(key[r1], key[r1 + 1]));
        ++r1;
        if (++steps > MERGE_RUN_CAP) { capped = true; break; }
      }
      if (!capped && (r1 - r0 + 1) <= size_t(MERGE_RUN_CAP)) {
        const int ra = r0 > 0 ? shared_levels<DIM>(key[r0 - 1], key[r0]) : -1;
        const int rb = r1 + 1 < n ? shared_levels<DIM>(key[r1], key[r1 + 1]) : -1;
        if (inner >= min(max(ra, rb) + 1, LM)) {
          head = (s == r0);
          a = ra;
          b = rb;
        }
      }
    }
    ab[s] = head ? make_uchar2(static_cast<unsigned char>(a + 1), static_cast<unsigned char>(b + 1))
                 : make_uchar2(255, 255);  // NOT_HEAD: merged into the unit before it
    if (!head) max_shared_plus1[1] = 1u;  // "some leaf is a merged unit" (rare): slow summation paths
    cnt[s] = head ? unsigned(max(0, b - a) + 1) : 0u;
    if (head) a1 = static_cast<unsigned char>(a + 1);
  }
  // minima of a1 for the run-end queries (NsvTables): per body, per aligned window of 16 bodies (shuffles),
  // per block of 256 (one barrier)
  {
    const unsigned lane = threadIdx.x & 31u;
    nsv1[s] = a1;
    unsigned m = a1;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
    if ((lane & 15u) == 0u) nsv1[n_pad + (s >> 4)] = static_cast<unsigned char>(m);
    m = min(m, __shfl_xor_sync(FULL, m, 16));
    if (lane == 0u) warp_min[threadIdx.x >> 5] = static_cast<unsigned char>(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned bm = warp_min[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) bm = min(bm, unsigned(warp_min[w]));
      nsv2[tile] = static_cast<unsigned char>(bm);
    }
  }
  __syncthreads();  // everybody is done with this tile's buffer (and with warp_min)
  }
  // sizes the next sort and validates this one (a truncated sort is exact iff no two neighbours
  // with different keys agree on every sorted bit)
  deepest = __reduce_max_sync(FULL, deepest);
  if ((threadIdx.x & 31) == 0 && deepest > 0u) atomicMax(max_shared_plus1, deepest);
}

// ---------------------------------------------------------------------------------------------
// K6  cell table in DFS pre-order (children in ascending digit order).
//     cell(head s, level l) = cell_start[s] + (l - (a_s + 1)).
//     Run ends and parents come from short linear scans over the per-body shared-level bytes
//     (`ab`, L1-resident) with a galloping search on the keys as the fallback for large cells.
//     Cells with <= SMALL_CELL bodies also get their mass / centre of mass here, by summing their
//     bodies in order; only the larger cells are left to the bottom-up pass (K7).
// ---------------------------------------------------------------------------------------------
struct CellArrays {
  uint8_t* level;
  uint32_t* head;
  uint32_t* count;
  uint32_t* skip;
  uint32_t* parent;
  uint32_t* arrived;
  double4* centre_ext;  // {cx, cy, cz, half-width}
  double4* com;         // {X, Y, Z, M}
  uint32_t capacity;
  uint32_t small;        // cells with <= this many bodies are summed directly in cells_kernel ...
  const unsigned* bad;   // != 0: the keys are not fully ordered (truncated sort too short): skip the build
  uint32_t top_level;    // ... unless they lie above this level: those are always summed from their children
};

// The fourth word of a cell's centre record is not the half-width (extent / 2^level: derived, exactly, from
// the level) but the walk's link: skip link in the low 32 bits, level in the next 8.  One record load then tells
// the walk everything it needs to decide about a cell and to move on; the centre of mass is fetched only for
// the cells it takes (the walk is bound by L1 wavefronts: ncu l1tex__data_pipe_lsu_wavefronts 88 % of peak at
// theta = 0.7 with three unconditional loads per visit).
__device__ __forceinline__ double pack_link(uint32_t skip, int level) {
  return __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(level & 0xff) << 32) | skip));
}
__device__ __forceinline__ uint32_t link_skip(double w) { return static_cast<uint32_t>(__double_as_longlong(w)); }
__device__ __forceinline__ int link_level(double w) { return int((__double_as_longlong(w) >> 32) & 0xff); }
// extent / 2^level, the value `half *= 0.5` reaches after `level` steps (exact; the exponent trick needs a normal
// result, else ldexp)
__device__ __forceinline__ double half_width_at(double ext0, int level) {
  const int hi = __double2hiint(ext0);
  if (((hi >> 20) & 0x7ff) > level + 1 && ((hi >> 20) & 0x7ff) != 0x7ff)
    return __hiloint2double(hi - (level << 20), __double2loint(ext0));
  return ldexp(ext0, -level);
}

// Key levels [0, TOP_LEVEL) form the "top tree": 4^6 = 8^4 = 4096 level-K prefixes.  A sharded build cuts the
// key space between ranks at level-K prefixes, every rank builds the cells at levels >= K of its own range,
// and the cells above are recomputed on every rank from the level-K cells' sums (top_build_kernel).  For that
// to give the bits of the single-GPU build, a cell above level K must be summed the same way in both: always
// from its children in ascending digit order (never from its bodies, however few it has), with ComSum below.
template <int DIM>
struct TopTree {
  static constexpr int K = (DIM == 3) ? 4 : 6;
  static constexpr int R = 1 << DIM;
  static constexpr uint32_t SLOTS = 1u << (DIM * K);                       // level-K prefixes
  static constexpr uint32_t CELLS = ((1u << (DIM * (K + 1))) - 1u) / (R - 1);  // levels 0..K, dense
  __host__ __device__ static constexpr uint32_t offset(int level) { return ((1u << (DIM * level)) - 1u) / (R - 1); }
};

// Mass-weighted sum of child cells / units: the ONE definition every kernel uses (explicit fma: the bits
// must not depend on what the compiler contracts in which kernel).
struct ComSum {
  double m = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
  __device__ __forceinline__ void add(const double4& q) {
    m += q.w;
    sx = fma(q.w, q.x, sx);
    sy = fma(q.w, q.y, sy);
    sz = fma(q.w, q.z, sz);
  }
  // a massless cell (the reference panics there): geometric centre g instead of 0/0
  __device__ __forceinline__ double4 finish(const double4& g) const {
    if (m != 0.0) {
      const double inv = 1.0 / m;  // one division: the reference's own (...) * inv_total_mass form (lib.rs:43-49)
      return make_double4(sx * inv, sy * inv, sz * inv, m);
    }
    return make_double4(g.x, g.y, g.z, 0.0);
  }
};

// 24 instead of 16 (r02, 1 x B200): cells_kernel + 5 us (longer divergent sums), kids_kernel - 2 us, climb_kernel - 5 us
// (one level less of its store / atomic / load chain): c3 244.1 -> 241.5 us per step, octree 247.8 -> 244.9.
// The sums differ from the 16-body build's within fp64 rounding (another order of additions), not bit for bit.
constexpr uint32_t SMALL_CELL = 24;
constexpr unsigned char NOT_HEAD = 255;  // ab[j].x of a body merged into the unit before it

// {x, y, z, m} of the leaf made of sorted bodies [s, e): masses add, the member inserted last
// (largest original index) keeps its place (octree.rs:69-80)
__device__ __forceinline__ double4 unit_leaf(const double4* __restrict__ sp,
                                             const uint32_t* __restrict__ perm, size_t s, size_t e) {
  double4 q = sp[s];
  if (e > s + 1) {
    double m = 0.0;
    size_t newest = s;
    uint32_t newest_idx = perm[s];
    for (size_t j = s; j < e; ++j) {
      m += sp[j].w;
      const uint32_t pj = perm[j];
      if (pj > newest_idx) { newest_idx = pj; newest = j; }
    }
    q = sp[newest];
    q.w = m;
  }
  return q;
}

// Minima for "first body j >= j0 that starts a cell at level <= l", i.e. the end of a cell's run of
// bodies (in DFS pre-order the first later cell that is not deeper is the next non-descendant, and
// cell_start[] of the body that heads it is its index):
//   t1[j]      a1 of body j (255 for a body that heads no cell, and for the padding up to n_pad)
//   t16[w]     minimum over the aligned window of 16 bodies w             (unit_kernel, shuffles)
//   t2[0][b]   minimum over the block of 256 bodies b                      (unit_kernel)
//   t2[k][b]   range minima over 2^k blocks inside super-blocks of 256 blocks, k = 1..8 (nsv_level2_block)
//   t3[q]      super-block minima (scanned linearly: n / 65536 entries)
// A query runs in rounds of INDEPENDENT 16-byte loads - nearly every warp of cells_kernel holds a cell
// whose run is long, and a chain of dependent byte loads there stalls the whole warp:
//   (1) the bytes of the 17..32 bodies from j0 on + the 16 window minima of j0's block;
//   (2) the bytes of the window that holds the hit                      -> the run ends inside the block
//   (3) the minima of the next 17..32 blocks; (4) the window minima of the block that holds the hit;
//   (5) the bytes of its window                                         -> runs of up to ~4000 bodies
// Longer runs find their block by descending t2 / t3 (9 dependent byte loads per level), then (4), (5).
struct NsvTables {
  const uint8_t* t1;
  const uint8_t* t16;
  size_t n_pad;   // bodies rounded up to whole blocks
  const uint8_t* t2;
  size_t b_pad;   // stride of a level of t2
  const uint8_t* t3;
  size_t n, nblocks, nsuper;
  // the sizes that follow from the body count (n_pad here is the LOGICAL one: whole blocks of the n bodies;
  // t16 and b_pad carry the allocation's strides)
  __device__ __forceinline__ void set_n(size_t nn) {
    n = nn;
    nblocks = (nn + 255) / 256;
    n_pad = nblocks * 256;
    nsuper = (nblocks + 255) / 256;
  }
};

// Sharded build only (slot_cell == nullptr otherwise): where the cells that carry a level-K key prefix
// ended up in this rank's table.  slot_cell[q] = 1 + index of the deepest cell holding exactly the bodies
// whose keys start with the level-K prefix q: the level-K cell itself, or the leaf above level K that holds
// them (one unit).  Zeroed per build; 0 = no body of this rank has the prefix.
struct TopSlots {
  uint32_t* slot_cell;
  int level;  // K
};

// What a sharded build adds to the build of one rank (nullptr: single GPU, every body).
struct ShardBuild {
  const uint64_t* cuts;  // device: this rank keeps the bodies with cuts[0] <= key < cuts[1]
  uint32_t* n_local;     // device: how many those were (written by the sort, read by every later kernel)
  size_t n_cap;          // capacity of the per-rank arrays, in bodies
  TopSlots slots;
  uint32_t* perm_out;    // where the sort leaves the permutation (the rank's block of the exchange buffer)
};

struct BuildOut {
  CellArrays cells;
  NRef nref;
  size_t n_cap;
};


// first index in [j, end) (end - j <= 256, same 256-aligned block) whose entry is <= l, else end
__device__ __forceinline__ size_t nsv_descend(const uint8_t* __restrict__ tab, size_t stride, size_t j, size_t end,
                                              unsigned l) {
#pragma unroll
  for (int k = 8; k >= 0; --k) {
    const size_t step = size_t(1) << k;
    if (j + step <= end && tab[size_t(k) * stride + j] > l) j += step;
  }
  return j;
}

// first byte at or after `off` of the 16 in w that is <= l (splat = l in every byte), else 16
__device__ __forceinline__ int first_le_16(const uint4& w, unsigned splat, int off) {
  const unsigned m[4] = {__vcmpleu4(w.x, splat), __vcmpleu4(w.y, splat), __vcmpleu4(w.z, splat),
                         __vcmpleu4(w.w, splat)};
  int pos = 16;
#pragma unroll
  for (int q = 3; q >= 0; --q) {
    unsigned mm = m[q];
    const int skip = off - 4 * q;
    if (skip >= 4) mm = 0u;
    else if (skip > 0) mm &= 0xffffffffu << (8 * skip);
    if (mm) pos = 4 * q + ((__ffs(mm) - 1) >> 3);
  }
  return pos;
}

__device__ __forceinline__ uint4 ld16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

// the answer inside block b, whose minimum is known to be <= l: window minima, then the window's bytes
__device__ __forceinline__ size_t nsv_in_block(const NsvTables& tv, size_t b, unsigned splat) {
  const int w = first_le_16(ld16(tv.t16 + (b << 4)), splat, 0);
  if (w == 16) return tv.n;  // (not reached)
  const size_t jw = (b << 8) + (size_t(w) << 4);
  return jw + size_t(first_le_16(ld16(tv.t1 + jw), splat, 0));
}

__device__ __forceinline__ size_t nsv_next_le(const NsvTables& tv, size_t j0, unsigned l) {
  if (j0 >= tv.n) return tv.n;
  // (bytes of bodies past n are 255 and never match, so a hit is always < n)
  const unsigned splat = l * 0x01010101u;
  const size_t base = j0 & ~size_t(15), blk = j0 >> 8;
  const bool two = base + 32 <= tv.n_pad;
  const uint4 w0 = ld16(tv.t1 + base);
  const uint4 w1 = two ? ld16(tv.t1 + base + 16) : make_uint4(~0u, ~0u, ~0u, ~0u);
  const uint4 wm = ld16(tv.t16 + (blk << 4));
  int pos = first_le_16(w0, splat, int(j0 - base));
  if (pos < 16) return base + size_t(pos);
  pos = first_le_16(w1, splat, 0);
  if (pos < 16) return base + 16 + size_t(pos);
  // windows of j0's block past the bytes just examined (none if those reached the block's end)
  const size_t seen = base + (two ? 32 : 16);
  if ((seen >> 8) == blk) {
    pos = first_le_16(wm, splat, int((seen >> 4) - (blk << 4)));
    if (pos < 16) {
      const size_t jw = (blk << 8) + (size_t(pos) << 4);
      return jw + size_t(first_le_16(ld16(tv.t1 + jw), splat, 0));
    }
  }
  // following blocks: their minima, 17..32 at a time (entries of blocks past nblocks are not initialised:
  // a match there lies after every real block and is discarded by position)
  const size_t b0 = blk + 1;
  if (b0 >= tv.nblocks) return tv.n;
  const size_t bb = b0 & ~size_t(15);
  const bool two_b = bb + 32 <= tv.b_pad;
  const uint4 m0 = ld16(tv.t2 + bb);
  const uint4 m1 = two_b ? ld16(tv.t2 + bb + 16) : make_uint4(~0u, ~0u, ~0u, ~0u);
  pos = first_le_16(m0, splat, int(b0 - bb));
  if (pos == 16) {
    pos = first_le_16(m1, splat, 0);
    pos = pos < 16 ? pos + 16 : 32;
  }
  size_t b = bb + size_t(pos);
  if (pos == 32) {  // not within those: descend the block tables from the first block not yet examined
    b = bb + (two_b ? 32 : 16);
    if (b >= tv.nblocks) return tv.n;
    size_t sb_end = min(tv.nblocks, (b | 255) + 1);
    b = nsv_descend(tv.t2, tv.b_pad, b, sb_end, l);
    if (b == sb_end) {  // not in the rest of this super-block: whole super-blocks, then inside the one that has it
      if (sb_end == tv.nblocks) return tv.n;
      size_t q = sb_end >> 8;
      while (q < tv.nsuper && tv.t3[q] > l) ++q;
      if (q >= tv.nsuper) return tv.n;
      b = q << 8;
      sb_end = min(tv.nblocks, b + 256);
      b = nsv_descend(tv.t2, tv.b_pad, b, sb_end, l);
      if (b == sb_end) return tv.n;  // (not reached: t3[q] <= l)
    }
  }
  if (b >= tv.nblocks) return tv.n;
  return nsv_in_block(tv, b, splat);
}

// K6  one thread per unit head s, for the chain of cells it heads (levels a+1 .. leaf level).
//     Phase 1 (per lane, its own chain): one replay of the head's key digits from the root to its leaf,
//     writing level, head, geometric centre / half-width of every cell of the chain on the way, then the
//     leaf (the unit itself).
//     Phase 2: the INTERNAL cells of the warp's 32 chains (0 for most lanes, several for a few) are dealt
//     out evenly to the lanes - a prefix sum over the per-lane counts and a 5-step search by shuffles name
//     the owner of task t - so every lane runs one range-minimum query (the end of the cell's run of
//     bodies: first later body that starts a cell at a level <= its own), hence body count and skip link,
//     and - for cells of <= SMALL_CELL bodies - the mass / centre of mass as a sum over its units in
//     body order.  Larger cells are left to the bottom-up pass (K7).
template <int DIM, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) cells_kernel(const uint64_t* __restrict__ key,
                                                    const double4* __restrict__ sp,
                                                    const uint32_t* __restrict__ perm,
                                                    const uchar2* __restrict__ ab,
                                                    const uint32_t* __restrict__ cell_start, NRef nref,
                                                    const unsigned long long* __restrict__ extent_bits,
                                                    const unsigned* __restrict__ tree_meta,
                                                    unsigned* __restrict__ sticky, NsvTables tv, CellArrays cells,
                                                    TopSlots slots) {
  pb_pdl_sync();
  constexpr int LM = TreeDim<DIM>::LM;
  const size_t n = nref.get();
  tv.set_n(n);
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  if (size_t(blockIdx.x) * blockDim.x >= n && blockIdx.x != 0) return;  // (surplus CTAs of a grid sized for the capacity)
  const uint32_t total = cell_start[n];
  if (s == 0) {
    // worst case over every build since the last host check (builds run unverified in between):
    // [0] cells needed, [1] 1 + deepest level shared by distinct neighbouring keys
    atomicMax(&sticky[0], total);
    atomicMax(&sticky[1], tree_meta[0]);
  }
  if (total > cells.capacity || *cells.bad) return;  // host grows the table / sorts all bits and re-runs (grid-uniform)
  // everything phase 1 reads is fetched up front, unconditionally: one round of memory latency instead
  // of four dependent ones (the body after s is the end of the leaf unless it was merged into it)
  const bool in_range = s < n;
  const uchar2 abv = in_range ? ab[s] : make_uchar2(NOT_HEAD, NOT_HEAD);
  const uchar2 abn = s + 1 < n ? ab[s + 1] : make_uchar2(0, 0);
  const uint32_t cs0 = in_range ? cell_start[s] : 0u, cs1 = in_range ? cell_start[s + 1] : 0u;
  const uint64_t kme = in_range ? key[s] : 0ull;
  const double4 me = in_range ? sp[s] : make_double4(0.0, 0.0, 0.0, 0.0);
  const double ext0 = __longlong_as_double(static_cast<long long>(*extent_bits));
  const bool merged_units = tree_meta[1] != 0u;
  const bool is_head = abv.x != NOT_HEAD;  // (lanes past n and merged bodies stay for the shuffles)
  // Reference path: the centres along the path of the CTA's FIRST body, level by level, computed once by thread 0
  // while the loads above are in flight.  A body shares its leading `lcp` digits with that body, and the centre at
  // a level depends on the digits above it only, so every lane starts its own replay at level min(lcp, top)
  // with the reference's doubles - the same operations on the same operands as a replay from the root, hence
  // the same bits - and runs ~5 instead of ~11 iterations (256 neighbours of 10^6 sorted bodies share ~6 levels).
  __shared__ double s_ref[LM + 1][4];
  __shared__ uint64_t s_refkey;
  if (threadIdx.x == 0) {
    s_refkey = kme;  // (s < n for thread 0 of every CTA that gets here, or n == 0 and nobody is a head)
    double half = ext0, cx = 0.0, cy = 0.0, cz = 0.0;
    uint64_t digits = kme << (64 - DIM * LM);
    for (int l = 0; l <= LM; ++l) {
      s_ref[l][0] = cx; s_ref[l][1] = cy; s_ref[l][2] = cz; s_ref[l][3] = half;
      const unsigned digit = unsigned(digits >> (64 - DIM));
      digits <<= DIM;
      half *= 0.5;
      cx += with_sign(half, !(digit & 1u));
      cy += with_sign(half, !(digit & 2u));
      if (DIM == 3) cz += with_sign(half, !(digit & 4u));
    }
  }
  __syncthreads();
  int top = 0, leaf_level = 0;
  uint32_t c0 = 0;
  if (is_head) {
    c0 = cs0;
    if (s == 0) cells.parent[0] = NO_PARENT;
    const int a = int(abv.x) - 1, b = int(abv.y) - 1;
    top = a + 1;
    leaf_level = max(a, b) + 1;
    const uint64_t diff = (kme ^ s_refkey) << (64 - DIM * LM);
    const int lcp = diff ? __clzll(static_cast<long long>(diff)) / DIM : LM;
    const int l0 = min(min(lcp, top), leaf_level);
    const uint32_t leaf_skip = abn.x != NOT_HEAD ? cs1 : 0u;  // (merged unit: written once its end is known)
    double half = s_ref[l0][3];
    double cx = s_ref[l0][0], cy = s_ref[l0][1], cz = s_ref[l0][2];
    uint64_t digits = l0 < LM ? kme << (64 - DIM * LM + DIM * l0) : 0ull;  // next digit in the top DIM bits (pseudo levels compare the body)
    for (int l = l0; l <= leaf_level; ++l) {
      if (l >= top) {
        const uint32_t c = c0 + uint32_t(l - top);
        cells.level[c] = static_cast<uint8_t>(l);
        // {cx, cy, cz, link}: the leaf's link is complete here (its skip link is the next body's first cell);
        // an internal cell's carries the level, its skip half is filled in by phase 2 (a 4-byte store by the
        // lane that runs the cell's query; __syncwarp below orders the two stores)
        double2* rec = reinterpret_cast<double2*>(cells.centre_ext + c);
        rec[0] = make_double2(cx, cy);
        rec[1] = make_double2(cz, pack_link(l == leaf_level ? leaf_skip : 0u, l));
      }
      if (l < leaf_level) {  // from level l to level l+1 along the head body's path
        unsigned digit = unsigned(digits >> (64 - DIM));
        digits <<= DIM;
        if (l >= LM) {  // pseudo level below the key (rare: a real branch): compare the head body itself
          digit = unsigned(me.x > cx) | (unsigned(me.y > cy) << 1);
          if (DIM == 3) digit |= unsigned(me.z > cz) << 2;
        }
        half *= 0.5;
        cx += with_sign(half, !(digit & 1u));
        cy += with_sign(half, !(digit & 2u));
        if (DIM == 3) cz += with_sign(half, !(digit & 4u));
      }
    }
    // leaf: the unit itself
    const uint32_t c = c0 + uint32_t(leaf_level - top);
    if (abn.x != NOT_HEAD) {  // (always, unless bodies were merged)
      cells.count[c] = 1u;
      cells.skip[c] = cs1;
      cells.com[c] = me;
    } else {
      size_t e = s + 2;
      while (e < n && ab[e].x == NOT_HEAD) ++e;
      cells.count[c] = static_cast<uint32_t>(e - s);
      const uint32_t sk = cell_start[e];
      cells.skip[c] = sk;
      reinterpret_cast<uint32_t*>(cells.centre_ext + c)[6] = sk;  // (the low half of the link word)
      cells.com[c] = unit_leaf(sp, perm, s, e);
    }
    if (slots.slot_cell && top <= slots.level) {  // s is the first body of its level-K prefix
      const uint32_t q = uint32_t(kme >> (DIM * (LM - slots.level)));
      slots.slot_cell[q] = c0 + uint32_t(min(leaf_level, slots.level) - top) + 1u;
    }
  }

  // phase 2: task t of the warp = internal cell number (t - excl[o]) of the chain of lane o
  __syncwarp();  // phase 1's record stores before phase 2's stores into the same link words (other lanes)
  const int mine = is_head ? leaf_level - top : 0;
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, d);
    if (int(lane) >= d) incl += v;
  }
  const int n_tasks = __shfl_sync(FULL, incl, 31);
  const int excl = incl - mine;
  for (int t0 = 0; t0 < n_tasks; t0 += 32) {
    const int t = t0 + int(lane);
    int o = 0;  // number of lanes whose inclusive count is <= t  ==  the lane that owns task t
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int probe = __shfl_sync(FULL, incl, o + step - 1);
      if (probe <= t) o += step;
    }
    const int k = t - __shfl_sync(FULL, excl, o);
    const int lev = __shfl_sync(FULL, top, o) + k;  // internal levels top .. leaf_level-1
    const uint32_t c = __shfl_sync(FULL, c0, o) + uint32_t(k);
    if (t >= n_tasks) continue;
    const size_t so = s - lane + size_t(o);  // the chain's head
    const size_t e = nsv_next_le(tv, so + 1, unsigned(lev));
    const uint32_t cnt = static_cast<uint32_t>(e - so);
    cells.count[c] = cnt;
    const uint32_t sk_c = cell_start[e];
    cells.skip[c] = sk_c;
    reinterpret_cast<uint32_t*>(cells.centre_ext + c)[6] = sk_c;  // (low half of the link word; phase 1 wrote the level)
    if (cnt > cells.small || uint32_t(lev) < cells.top_level) continue;
    ComSum sum;
    if (!merged_units) {  // no merged unit anywhere: plain sums over the run
      double4 q = sp[so];
      for (size_t j = so; j < e; ++j) {  // (the next body is fetched before this one is added)
        const double4 nq = sp[min(j + 1, e - 1)];
        sum.add(q);
        q = nq;
      }
    } else {
      size_t j = so;
      while (j < e) {  // units in order
        size_t je = j + 1;
        while (je < e && ab[je].x == NOT_HEAD) ++je;
        sum.add(unit_leaf(sp, perm, j, je));
        j = je;
      }
    }
    if (sum.m != 0.0) {
      cells.com[c] = sum.finish(make_double4(0.0, 0.0, 0.0, 0.0));
    } else {
      // (written by another lane in phase 1: recompute instead of reading it back)
      double half = ext0;
      double gx = 0.0, gy = 0.0, gz = 0.0;
      const uint64_t ko = key[so];
      const double4 mo = sp[so];
      for (int l = 0; l < lev; ++l) {
        unsigned digit;
        if (l < LM) {
          digit = unsigned((ko >> (DIM * (LM - 1 - l))) & ((1u << DIM) - 1u));
        } else {
          digit = unsigned(mo.x > gx) | (unsigned(mo.y > gy) << 1);
          if (DIM == 3) digit |= unsigned(mo.z > gz) << 2;
        }
        half *= 0.5;
        gx += with_sign(half, !(digit & 1u));
        gy += with_sign(half, !(digit & 2u));
        if (DIM == 3) gz += with_sign(half, !(digit & 4u));
      }
      cells.com[c] = make_double4(gx, gy, gz, 0.0);
    }
  }
}

// K6c  one thread per internal cell: its children are p+1, skip[p+1], ... (at most 2^DIM of them,
//      whatever the size of the cell), each gets its parent link.
__global__ void __launch_bounds__(256) parent_kernel(const uint32_t* __restrict__ cell_start, size_t n,
                                                     CellArrays cells) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = cell_start[n];
  if (total > cells.capacity || *cells.bad || p >= total) return;
  const uint32_t end = cells.skip[p];
  for (uint32_t ch = p + 1u; ch < end;) {
    cells.parent[ch] = p;
    const uint32_t next = cells.skip[ch];
    if (next <= ch) break;  // (never in a well-formed table; keeps a broken one from hanging the GPU)
    ch = next;
  }
}

// cells.head (the first sorted body of every cell) is part of the inspected table only - no kernel of the step
// reads it - so it is filled when the table is read back: body s heads cells cell_start[s] .. cell_start[s+1]-1.
__global__ void __launch_bounds__(256) head_kernel(const uint32_t* __restrict__ cell_start, size_t n, CellArrays cells) {
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (s >= n || cell_start[n] > cells.capacity || *cells.bad) return;
  for (uint32_t c = cell_start[s]; c < cell_start[s + 1]; ++c) cells.head[c] = static_cast<uint32_t>(s);
}

// ---------------------------------------------------------------------------------------------
// K7  bottom-up masses and centres of mass of the cells with more than SMALL_CELL bodies.
//     One thread per unit: the shallowest cell of its chain that K6 already finished climbs; the
//     child whose arrival completes a cell (arrived bodies == cell bodies) sums the children in
//     pre-order and continues upward.  (reference: incremental pairwise update, octree.rs:83-88 /
//     lib.rs:40-50 — same value up to fp64 rounding; zero-mass cells fall back to the geometric
//     centre.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 ld_cg_double4(const double4* p) {
  const double2 lo = __ldcg(reinterpret_cast<const double2*>(p));
  const double2 hi = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// K7 (current form)  the same sums in two kernels without the per-unit scan and with one round of
//     loads per level instead of a chain through the skip links:
//     kids_kernel, one thread per cell, acts on the cells K6 left open (more than SMALL_CELL bodies, not
//     a leaf): walks the children once and leaves kid_tab[p] = {number of children, child 1, child 2, ..}
//     (child 0 is p + 1), the parent link of every child that is open itself, arrived[p] = the bodies of
//     the children K6 already finished, and ready[p] = "all of them": a climb starts here.
//     climb_kernel: the thread of a ready cell sums its children in pre-order, then arrives at the parent
//     (release atomic on the body counter); whoever completes a cell goes on with it.  Per level:
//     parent link -> atomic (the child table and the counts load meanwhile) -> the children's sums (one
//     round of ld.cg) -> store.  Same operations in the same order as com_kernel: identical bits.
constexpr uint32_t KIDS_OVERFLOW = 0xffffffffu;
constexpr uint32_t READY_LISTS = 16;
constexpr uint32_t READY_STRIDE = 32;  // words between the lists' counters: one 128-byte line (one L2 atomic unit) each  // more children than 2^DIM (sibling leaves of a pseudo level)

template <int DIM>
__global__ void __launch_bounds__(256) kids_kernel(const uint32_t* __restrict__ cell_start, NRef nref,
                                                   CellArrays cells, uint32_t* __restrict__ kid_tab,
                                                   uint32_t* __restrict__ ready_list, unsigned* __restrict__ n_ready,
                                                   uint32_t sub_cap) {
  pb_pdl_sync();
  constexpr uint32_t K = 1u << DIM;
  const uint32_t total = cell_start[nref.get()];
  if (total > cells.capacity || *cells.bad) return;
  // (grid-stride over the cells, whole warps together: the grid may be sized for fewer cells than exist)
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; (p & ~31u) < total; p += gridDim.x * blockDim.x) {
  const bool live = p < total;
  uint32_t cnt = 0, end = 0, lev = 0;
  if (live) {
    cnt = cells.count[p];
    end = cells.skip[p];
    lev = cells.level[p];
  }
  // open = left to the bottom-up pass by K6: more than `small` bodies, or above the top level (any size)
  const bool open = live && (cnt > cells.small || lev < cells.top_level) && end != p + 1u;
  if (!__any_sync(FULL, open)) continue;  // (most warps: nothing but finished cells)
  bool start = false;
  if (open) {
    uint32_t nk = 0, pre = 0;
    const bool kids_above_top = lev + 1u < cells.top_level;
    for (uint32_t ch = p + 1u; ch < end;) {
      const uint32_t c_cnt = cells.count[ch], next = cells.skip[ch];
      if ((c_cnt <= cells.small && !kids_above_top) || next == ch + 1u) pre += c_cnt;  // finished by K6
      else cells.parent[ch] = p;                                  // climbs later
      if (nk >= 1u && nk < K) kid_tab[size_t(p) * K + nk] = ch;
      ++nk;
      if (next <= ch) break;  // (never in a well-formed table; keeps a broken one from hanging the GPU)
      ch = next;
    }
    kid_tab[size_t(p) * K] = nk <= K ? nk : KIDS_OVERFLOW;
    cells.arrived[p] = pre;
    start = pre == cnt;  // a climb starts here
  }
  // the warp's climb starts go to one of READY_LISTS global lists with one atomic (one list: ~25000
  // same-address atomics with a return value cost ~20 us; a block-wide hand-over costs a barrier behind
  // the slowest walk; the order inside a list does not matter: every sum is taken by one thread, over
  // the children in pre-order)
  const unsigned who = __ballot_sync(FULL, start);
  if (!who) continue;
  const unsigned lane = threadIdx.x & 31u, leader = unsigned(__ffs(who) - 1);
  const uint32_t l = (p >> 5) % READY_LISTS;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(n_ready + l * READY_STRIDE, unsigned(__popc(who)));
  base = __shfl_sync(FULL, base, leader);
  if (start) ready_list[size_t(l) * sub_cap + base + __popc(who & ((1u << lane) - 1u))] = p;
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) climb_kernel(const uint32_t* __restrict__ cell_start, NRef nref,
                                                    CellArrays cells, const uint32_t* __restrict__ kid_tab,
                                                    const uint32_t* __restrict__ ready_list,
                                                    const unsigned* __restrict__ n_ready, uint32_t sub_cap) {
  pb_pdl_sync();
  constexpr int K = 1 << DIM;
  const uint32_t total = cell_start[nref.get()];
  if (total > cells.capacity || *cells.bad) return;
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t l = gt % READY_LISTS, starts = n_ready[l * READY_STRIDE];
  for (uint32_t it = gt / READY_LISTS; it < starts; it += gridDim.x * blockDim.x / READY_LISTS) {
  const uint32_t p0 = ready_list[size_t(l) * sub_cap + it];
  // what the climb needs to know about a cell (immutable during this kernel): fetched for the parent while
  // the children's sums of the current cell are still on their way, so that a level costs one round of
  // loads (the children's sums) + store + release atomic
  struct Meta {
    uint32_t parent, count;
    uint32_t kid[K];
  };
  auto load_meta = [&](uint32_t c) {
    Meta mt;
    mt.parent = cells.parent[c];
    mt.count = cells.count[c];
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
      const uint4 w = reinterpret_cast<const uint4*>(kid_tab + size_t(c) * K)[q];
      mt.kid[4 * q] = w.x; mt.kid[4 * q + 1] = w.y; mt.kid[4 * q + 2] = w.z; mt.kid[4 * q + 3] = w.w;
    }
    return mt;
  };
  uint32_t c = p0;
  Meta me = load_meta(c);
  while (true) {
    const uint32_t nk = me.kid[0];
    me.kid[0] = c + 1u;
    ComSum sum;
    double4 q[K];
    if (nk != KIDS_OVERFLOW) {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (uint32_t(i) < nk) q[i] = ld_cg_double4(&cells.com[me.kid[i]]);
    }
    const double4 g = cells.centre_ext[c];
    Meta up;  // the parent's, in case this thread completes it
    if (me.parent != NO_PARENT) up = load_meta(me.parent);
    if (nk != KIDS_OVERFLOW) {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (uint32_t(i) < nk) sum.add(q[i]);
    } else {
      const uint32_t end = cells.skip[c];
      for (uint32_t ch = c + 1u; ch < end;) {
        sum.add(ld_cg_double4(&cells.com[ch]));
        const uint32_t next = cells.skip[ch];
        if (next <= ch) break;
        ch = next;
      }
    }
    cells.com[c] = sum.finish(g);
    if (me.parent == NO_PARENT) break;
    // acq_rel: com[c] must be visible before the arrival is (release), and the thread that completes the
    // parent must see the sums its siblings published before their arrivals (acquire) - ordered by the PTX
    // memory model itself, not by the control dependency and ld.cg's bypass of the L1.
    cuda::atomic_ref<uint32_t, cuda::thread_scope_device> arrived(cells.arrived[me.parent]);
    const uint32_t old = arrived.fetch_add(me.count, cuda::std::memory_order_acq_rel);
    if (old + me.count != up.count) break;
    c = me.parent;
    me = up;
  }
  }
}

// ---------------------------------------------------------------------------------------------
// target lists for sharded evaluation: sorted positions whose original index is in [t0, t1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tgt_flag_kernel(const uint32_t* __restrict__ perm, size_t n,
                                                       uint32_t t0, uint32_t t1,
                                                       uint32_t* __restrict__ flags) {
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (s < n) {
    const uint32_t i = perm[s];
    flags[s] = (i >= t0 && i < t1) ? 1u : 0u;
  }
}
__global__ void __launch_bounds__(256) tgt_scatter_kernel(const uint32_t* __restrict__ flags,
                                                          const uint32_t* __restrict__ offs, size_t n,
                                                          uint32_t* __restrict__ list) {
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (s < n && flags[s]) list[offs[s]] = static_cast<uint32_t>(s);
}

// ---------------------------------------------------------------------------------------------
// K8  Barnes-Hut walk.  One thread per target, targets in Morton order, stack-free: each lane
//     follows its own pre-order path through the cell table (c -> c+1 when it opens a cell,
//     c -> skip[c] when it accepts one), so every target gets exactly the reference's interaction
//     list (octree.rs:130-158: accept iff half_width / |p − centre| < theta, leaves always).
//     The 32 lanes of a warp are spatial neighbours and touch mostly the same cell records, which
//     the L1 serves; a warp-cooperative variant (one cell per iteration for the whole warp, chosen
//     with __reduce_min_sync, lanes that do not need it idle) was measured 1.5x SLOWER at
//     theta = 0.7 (2.60 vs 1.71 ms at 4.2 M bodies): it executes the union of the lanes' paths
//     (210 cells per warp) instead of the longest one (~110).
//     Acceptance and p_b − p_a are evaluated in fp64 (a heavy body inside an accepted cell sits
//     ~1e-7 from that cell's centre of mass); the force law runs in fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Unconditional, un-sinkable loads: the compiler otherwise moves each of the three per-cell loads
// behind the branch that first uses it, and every visit pays three serial memory latencies.
__device__ __forceinline__ double4 ld_now_double4(const double4* p) {
  double4 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ bool accept_cell(double px, double py, double pz, const double4& ce,
                                            double theta, double theta2) {
  if (!(theta > 0.0)) return false;  // half_width / r >= 0 is never < theta <= 0
  const double ax = px - ce.x, ay = py - ce.y, az = pz - ce.z;
  const double d2 = ax * ax + ay * ay + az * az;
  const double lhs = ce.w * ce.w, rhs = theta2 * d2;
  if (fabs(lhs - rhs) > 1e-12 * fmax(lhs, rhs)) return lhs < rhs;
  // borderline: the reference's own expression, operation for operation (no contraction)
  const double e2 = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
  const double r = __dsqrt_rn(e2);
  return __ddiv_rn(ce.w, r) < theta;
}

// (Measured and rejected, r02: prefetch.global.L1 of the three records the skip link points to, issued as soon as
// the link is known - 46 -> 49 us at c3, 1.71 -> 3.03 ms at 4.2 M bodies / theta 0.7: the prefetches evict the
// lines the neighbouring lanes are about to reuse.)
__device__ __forceinline__ double4 ld_now_double4_256(const double4* p) {  // one 256-bit load (LDG.E.256, sm_100)
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

template <int DIM>
__global__ void __launch_bounds__(256, 5) walk_kernel(const double4* __restrict__ sp,
                                                   const uint32_t* __restrict__ perm,
                                                   const uint8_t* __restrict__ fixed,
                                                   const uint32_t* __restrict__ tgt_list,
                                                   size_t n_targets, const uint32_t* __restrict__ cell_start,
                                                   size_t n, CellArrays cells,
                                                   const unsigned long long* __restrict__ extent_bits, double theta,
                                                   float easing, float tiny, float4* __restrict__ acc) {
  pb_pdl_sync();
  const uint32_t total = cell_start[n];
  if (total > cells.capacity || *cells.bad) return;
  const double ext0 = __longlong_as_double(static_cast<long long>(*extent_bits));
  const size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  bool active = t < n_targets;
  uint32_t s = 0, orig = 0;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (active) {
    s = tgt_list ? tgt_list[t] : static_cast<uint32_t>(t);
    orig = perm[s];
    const double4 p = sp[s];
    px = p.x; py = p.y; pz = p.z;
    if (fixed[orig]) active = false;  // transformers.rs:139-141
  }
  const double theta2 = theta * theta;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  uint32_t inter = 0;
  // every lane follows its own pre-order path (c -> c+1 or skip[c]); lanes of a warp are spatial
  // neighbours, so they mostly touch the same cache lines, and no lane waits for cells that only
  // other lanes need
  // Two records per visit - {centre, link} and the centre of mass - requested together at the top of the
  // iteration (measured, r02: fetching the centre of mass only for the cells that are taken saves 30 % of the L1
  // wavefronts but puts a second memory latency into every accepted visit: 1.60 vs 1.53 ms at 4.2 M bodies,
  // theta 0.7); the skip link travels in the centre record, so there is no third load; each record is ONE
  // 256-bit load (1.39 vs 1.47 ms with two 128-bit loads each; 1.71 ms with the three-array form of round 1).
  uint32_t c = active ? 0u : total;
  while (c < total) {
    const double4 ce = ld_now_double4_256(cells.centre_ext + c);
    const double4 cm = ld_now_double4_256(cells.com + c);
    const uint32_t sk = link_skip(ce.w);
    const bool leaf = (sk == c + 1u);
    bool take = leaf;
    if (!leaf) {
      const double4 cw = make_double4(ce.x, ce.y, ce.z, half_width_at(ext0, link_level(ce.w)));
      take = accept_cell(px, py, pz, cw, theta, theta2);
    }
    if (take) {
      const float dx = static_cast<float>(cm.x - px);
      const float dy = static_cast<float>(cm.y - py);
      const float dz = static_cast<float>(cm.z - pz);
      float r2 = fmaf(dx, dx, tiny);  // tiny keeps r = 0 (self / coincident: skipped by the
      r2 = fmaf(dy, dy, r2);          // reference, transformers.rs:145-147) finite: w·0 = 0
      r2 = fmaf(dz, dz, r2);
      const float sft = r2 + easing;
      const float w = rsqrt_approx(r2 * sft * sft);  // 1 / (|r| (r² + e))
      const float mw = static_cast<float>(cm.w) * w;
      fx = fmaf(mw, dx, fx);
      fy = fmaf(mw, dy, fy);
      fz = fmaf(mw, dz, fz);
      ++inter;
      c = max(sk, c + 1u);  // (sk > c in a well-formed table; a broken one must not hang the walk)
    } else {
      c = c + 1u;
    }
  }
  if (t < n_targets) {
    // a_i = f/m_a: a massless target is 0/0 = NaN in the reference (transformers.rs:154-158)
    if (active && sp[s].w == 0.0) fx = fy = fz = __int_as_float(0x7fc00000);
    acc[orig] = make_float4(fx, fy, fz, __uint_as_float(inter));
  }
}

// ---------------------------------------------------------------------------------------------
// Sharded Barnes-Hut (one rank per GPU; SURVEY §8e, north_star "exchanges top-level tree nodes").
//
// The key space is cut between the ranks at level-K prefixes (TopTree<DIM>::K); every rank holds every
// body's position (the integrator state is replicated) and
//   1. keeps, sorts and builds the tree of the bodies whose keys fall in its own range (tree_build with a
//      ShardBuild): its cell table is exact for every cell at level >= K of its range;
//   2. stores one record per level-K prefix of its range into EVERY rank's dense top tree, through peer
//      mappings over NVLink (top_export_kernel; the prefixes partition between the ranks, so the stores of all
//      ranks together ARE the all-gather); the kernel that follows signals the step's epoch to every rank;
//   3. rebuilds the cells ABOVE level K, redundantly and identically on every rank, from those records
//      (top_build_kernel): counts, leaves, centres of mass in ascending digit order with ComSum - the
//      operations the single-GPU build applies to the same cells, hence the same bits;
//   4. walks its own bodies (a Morton-contiguous slice of the targets) over the dense top tree and,
//      below an opened level-K cell, over the OWNER's cell table - through a peer pointer (NVLink P2P
//      loads) when the owner is another rank.  The visiting order is the global pre-order, so the fp32
//      sums are those of the single-GPU walk;
//   5. stores its accelerations (sorted order) and the permutation into its block of EVERY rank's exchange
//      buffer from inside the walk kernel; shard_scatter_kernel signals, waits for all ranks' blocks and puts
//      the accelerations into original order, and every rank advances the replicated state with the
//      single-GPU lean verlet step.  No collective call per step.
// The cuts are planned from a replicated (whole-set) build and stay until the host re-plans (multi.cu).
// ---------------------------------------------------------------------------------------------
// meta layout (u32), double-buffered by the parity of the epoch (a rank one phase ahead already writes the next
// step's counts while a slower rank's verlet still reads this step's):
constexpr int META_STRIDE = 32, META_BODIES = 1, META_CELLS = 9, META_BAD = 17, META_ANY_BAD = 64;

// Step 2: one thread per level-K prefix.  The prefixes of this rank's key range - cuts[rank] .. cuts[rank+1],
// every prefix belongs to exactly one rank - get their record (or "no body") in EVERY rank's dense top tree:
//   info.x  units (0 none, 1 a single unit: a leaf, possibly above level K, 2 an internal level-K cell) | rank << 8
//   info.y  index of that cell in this rank's table, info.z its skip link (end of its subtree), info.w bodies
//   com     {X, Y, Z, M} of that cell
template <int DIM>
__global__ void __launch_bounds__(256) top_export_kernel(const uint32_t* __restrict__ slot_cell, NRef nref,
                                                         const uint32_t* __restrict__ cell_start, CellArrays cells,
                                                         const uint64_t* __restrict__ cuts, PeerTargets pt,
                                                         uint32_t epoch) {
  pb_pdl_sync();
  using TT = TopTree<DIM>;
  constexpr uint32_t S = TT::SLOTS;
  constexpr int shift = DIM * (TreeDim<DIM>::LM - TT::K);
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = nref.get();
  const uint32_t total = cell_start[n];
  const bool bad = total > cells.capacity || *cells.bad;
  const uint32_t q_lo = uint32_t(min(cuts[pt.rank] >> shift, uint64_t(S)));
  const uint32_t q_hi = pt.rank + 1 == pt.world ? S : uint32_t(min(cuts[pt.rank + 1] >> shift, uint64_t(S)));
  if (q == 0) {
    const int base = int(epoch & 1u) * META_STRIDE;
    for (int r = 0; r < pt.world; ++r) {
      pt.meta[r][base + META_BODIES + pt.rank] = uint32_t(n);
      pt.meta[r][base + META_CELLS + pt.rank] = total;
      pt.meta[r][base + META_BAD + pt.rank] = bad ? 1u : 0u;
    }
  }
  if (q >= q_lo && q < q_hi) {
    uint4 inf = make_uint4(0u, 0u, 0u, 0u);
    double4 com = make_double4(0.0, 0.0, 0.0, 0.0);
    const uint32_t c1 = bad ? 0u : slot_cell[q];
    if (c1 != 0u && c1 <= total) {
      const uint32_t c = c1 - 1u;
      com = cells.com[c];
      const uint32_t end = cells.skip[c];
      inf = make_uint4(((end == c + 1u) ? 1u : 2u) | (uint32_t(pt.rank) << 8), c, end, cells.count[c]);
    }
    for (int r = 0; r < pt.world; ++r) {
      pt.top_info[r][TT::offset(TT::K) + q] = inf;
      pt.top_com[r][TT::offset(TT::K) + q] = com;
    }
  }
}

// Dense top tree, levels 0..K (index TopTree::offset(level) + prefix), identical on every rank.
struct TopView {
  double4* centre_ext;  // {cx, cy, cz, half-width}
  double4* com;         // {X, Y, Z, M}
  uint4* info;          // see top_export_kernel
  uint32_t* meta;       // see META_*
  uint64_t* cuts;       // [world + 1] key cuts of the plan (read only)
};

// Step 3 (one CTA): signals this rank's records, waits for every rank's, then rebuilds the cells above level K from their children in
// ascending digit order with ComSum - the single-GPU build's operations on the same cells.  Level K-1 reads the
// level-K entries from global memory (one round of loads), the levels above live in shared memory.
template <int DIM>
__global__ void __launch_bounds__(1024) top_build_kernel(int world, size_t n_total,
                                                         const unsigned long long* __restrict__ extent_bits,
                                                         TopView top, PeerTargets pt, uint32_t epoch) {
  pb_pdl_sync();
  using TT = TopTree<DIM>;
  constexpr int K = TT::K, R = TT::R;
  constexpr uint32_t ABOVE = TT::offset(K);  // cells above level K
  extern __shared__ __align__(16) unsigned char top_smem[];
  double4* s_com = reinterpret_cast<double4*>(top_smem);     // [ABOVE]
  uint4* s_info = reinterpret_cast<uint4*>(s_com + ABOVE);   // [ABOVE]
  const int tid = threadIdx.x;
  shard_signal_then_wait(pt, SHARD_FLAG_EXPORT, epoch, true);  // this rank's records (top_export_kernel) are in place
  const double ext0 = __longlong_as_double(static_cast<long long>(*extent_bits));
  if (tid == 0) {
    unsigned bad = pt.flags[pt.rank][SHARD_TIMEOUT];
    for (int r = 0; r < world; ++r) bad |= top.meta[int(epoch & 1u) * META_STRIDE + META_BAD + r];
    top.meta[META_ANY_BAD] = bad;
  }
  auto centre_of = [&](int l, uint32_t p) {  // the arithmetic of cells_kernel: digits of the prefix, root first
    double half = ext0, cx = 0.0, cy = 0.0, cz = 0.0;
    for (int j = 0; j < l; ++j) {
      const unsigned digit = (p >> (DIM * (l - 1 - j))) & unsigned(R - 1);
      half *= 0.5;
      cx += with_sign(half, !(digit & 1u));
      cy += with_sign(half, !(digit & 2u));
      if (DIM == 3) cz += with_sign(half, !(digit & 4u));
    }
    return make_double4(cx, cy, cz, half);
  };
  for (int l = K - 1; l >= 0; --l) {
    const uint32_t cells_l = 1u << (DIM * l);
    for (uint32_t p = tid; p < cells_l; p += 1024) {
      uint4 ci[R];
      double4 cc[R];
#pragma unroll
      for (int d = 0; d < R; ++d) {  // all children at once (unconditional loads)
        const uint32_t child = p * R + d;
        if (l == K - 1) {
          ci[d] = top.info[ABOVE + child];
          cc[d] = top.com[ABOVE + child];
        } else {
          ci[d] = s_info[TT::offset(l + 1) + child];
          cc[d] = s_com[TT::offset(l + 1) + child];
        }
      }
      uint32_t units = 0, bodies = 0;
      uint4 only = make_uint4(0u, 0u, 0u, 0u);
      double4 only_com = make_double4(0.0, 0.0, 0.0, 0.0);
      ComSum sum;
#pragma unroll
      for (int d = 0; d < R; ++d) {
        const uint32_t u = ci[d].x & 0xffu;
        if (u == 0u) continue;
        sum.add(cc[d]);
        units += u;  // (1 + anything >= 1 is already "internal")
        bodies += ci[d].w;
        only = ci[d];
        only_com = cc[d];
      }
      uint4 inf = make_uint4(0u, 0u, 0u, 0u);
      double4 com = make_double4(0.0, 0.0, 0.0, 0.0);
      if (units == 1u) {  // one unit below: this cell is (or lies above) its leaf - the unit's own record
        inf = only;
        com = only_com;
      } else if (units > 1u) {
        com = sum.finish(centre_of(l, p));
        inf = make_uint4(2u, 0u, 0u, bodies);
      }
      s_info[TT::offset(l) + p] = inf;
      s_com[TT::offset(l) + p] = com;
    }
    __syncthreads();
  }
  for (uint32_t t = tid; t < ABOVE; t += 1024) {
    top.info[t] = s_info[t];
    top.com[t] = s_com[t];
  }
  // geometric centres / half-widths of every top cell
  for (uint32_t t = tid; t < TT::CELLS; t += 1024) {
    int l = 0;
    while (l < K && TT::offset(l + 1) <= t) ++l;
    top.centre_ext[t] = centre_of(l, t - TT::offset(l));
  }
  // The cuts stay as planned (gravity_shard_plan, from a fully sorted replicated build) until the host re-plans:
  // the splitters of a rank's buckets are quantiles of ITS range, so a cut that moved by one level-K cell would
  // pour that cell's bodies (n / 4096 and more) into one edge bucket.  The host watches the per-rank counts
  // (meta) and re-plans from a replicated step when the balance has drifted.
}

struct PeerTables {   // every rank's cell table, as seen from this device (own table: local pointers)
  const double4* centre_ext[8];
  const double4* com[8];
  const uint32_t* skip[8];
  uint32_t capacity;  // the same on every rank (symmetric allocation)
};

// Step 4.  The result of target s - acceleration and original index - is stored into block `rank` of EVERY
// rank's exchange buffer (coalesced 16 + 4 byte stores over NVLink, issued while the rest of the grid still
// walks), then the phase is signalled.
template <int DIM>
__global__ void __launch_bounds__(256) walk_sharded_kernel(const double4* __restrict__ sp,
                                                           const uint32_t* __restrict__ perm,
                                                           const uint8_t* __restrict__ fixed, NRef nref, TopView top,
                                                           PeerTables peers, PeerTargets pt, size_t n_cap, uint32_t epoch,
                                                           double theta, float easing, float tiny) {
  pb_pdl_sync();
  using TT = TopTree<DIM>;
  constexpr int K = TT::K, R = TT::R;
  const size_t n = nref.get();
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  // (some rank's build abandoned: nothing is walked, the host replays the chunk; the phase is still signalled)
  const bool run = top.meta[META_ANY_BAD] == 0u && size_t(blockIdx.x) * blockDim.x < n;
  bool active = run && s < n;
  double px = 0.0, py = 0.0, pz = 0.0, pm = 0.0;
  uint32_t orig = 0;
  if (active) {
    const double4 p = sp[s];
    px = p.x; py = p.y; pz = p.z; pm = p.w;
    orig = perm[s];
    if (fixed[orig]) active = false;  // transformers.rs:139-141
  }
  const double theta2 = theta * theta;
  const double ext0 = top.centre_ext[0].w;  // the root's half-width (top_build_kernel)
  float fx = 0.f, fy = 0.f, fz = 0.f;
  uint32_t inter = 0;
  auto interact = [&](const double4& cm) {
    const float dx = static_cast<float>(cm.x - px);
    const float dy = static_cast<float>(cm.y - py);
    const float dz = static_cast<float>(cm.z - pz);
    float r2 = fmaf(dx, dx, tiny);
    r2 = fmaf(dy, dy, r2);
    r2 = fmaf(dz, dz, r2);
    const float sft = r2 + easing;
    const float w = rsqrt_approx(r2 * sft * sft);
    const float mw = static_cast<float>(cm.w) * w;
    fx = fmaf(mw, dx, fx);
    fy = fmaf(mw, dy, fy);
    fz = fmaf(mw, dz, fz);
    ++inter;
  };
  int l = 0;
  uint32_t p = 0;
  bool done = !active;
  while (!done) {
    const uint32_t t = TT::offset(l) + p;
    // the three records of the cell at once (the top tree is a few hundred KB: L1 / L2 hits)
    const uint4 inf = __ldg(top.info + t);
    const double4 ce = ld_now_double4(top.centre_ext + t);
    const double4 cm = ld_now_double4(top.com + t);
    const uint32_t units = inf.x & 0xffu;
    bool descend = false;
    if (units == 1u) {
      interact(cm);  // a leaf: always taken (octree.rs:151-152)
    } else if (units != 0u) {
      if (accept_cell(px, py, pz, ce, theta, theta2)) {
        interact(cm);
      } else if (l < K) {
        descend = true;
      } else {
        // an opened level-K cell: its subtree in the owner's table, pre-order (c -> c+1 / skip[c])
        const uint32_t owner = inf.x >> 8;
        const double4* __restrict__ t_ce = peers.centre_ext[owner];
        const double4* __restrict__ t_cm = peers.com[owner];
        const uint32_t end = min(inf.z, peers.capacity);
        uint32_t c = inf.y + 1u;
        while (c < end) {  // (same loop as walk_kernel, on the owner's table)
          const double4 ce2 = ld_now_double4(t_ce + c);
          const double4 cm2 = ld_now_double4(t_cm + c);
          const uint32_t sk = link_skip(ce2.w);
          bool take = (sk == c + 1u);
          if (!take) {
            const double4 cw = make_double4(ce2.x, ce2.y, ce2.z, half_width_at(ext0, link_level(ce2.w)));
            take = accept_cell(px, py, pz, cw, theta, theta2);
          }
          if (take) {
            interact(cm2);
            c = max(sk, c + 1u);
          } else {
            c = c + 1u;
          }
        }
      }
    }
    if (descend) {
      ++l;
      p *= R;
    } else {
      while (l > 0 && (p & uint32_t(R - 1)) == uint32_t(R - 1)) {
        --l;
        p >>= DIM;
      }
      if (l == 0) done = true;
      else ++p;
    }
  }
  if (run && s < n) {
    // a_i = f/m_a: a massless target is 0/0 = NaN in the reference (transformers.rs:154-158)
    if (active && pm == 0.0) fx = fy = fz = __int_as_float(0x7fc00000);
    const float4 out = make_float4(fx, fy, fz, __uint_as_float(inter));
    const size_t block = size_t(pt.rank) * (n_cap * 20);
    for (int r = 0; r < pt.world; ++r) {
      reinterpret_cast<float4*>(pt.xacc[r] + block)[s] = out;
      if (r != pt.rank) reinterpret_cast<uint32_t*>(pt.xacc[r] + block + n_cap * sizeof(float4))[s] = orig;
    }
  }
}

// Step 5: signals this rank's accelerations (walk_sharded_kernel, just completed on this stream), then puts block
// r's record j - once rank r has signalled - where the integrator reads it: acc[perm_r[j]] (original order, all bodies).
// Coalesced 20-byte reads, 16-byte scattered stores; the lean verlet step that follows is the single-GPU one.
__global__ void __launch_bounds__(256) shard_scatter_kernel(const char* __restrict__ xacc, size_t n_cap,
                                                            const uint32_t* __restrict__ n_locals, PeerTargets pt,
                                                            uint32_t epoch, float4* __restrict__ acc, size_t n) {
  // Persistent CTAs (one wave): each takes its grid-stride share of EVERY rank's block, in the order in which the
  // ranks signal (own block first): what has arrived is put in place while slower ranks are still walking.
  __shared__ unsigned s_next;
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // this rank's accelerations are in place everywhere
    __threadfence_system();
    for (int q = 0; q < pt.world; ++q)
      *reinterpret_cast<volatile uint32_t*>(pt.flags[q] + SHARD_FLAG_WALK + pt.rank) = epoch;
  }
  const size_t j0 = blockIdx.x * size_t(blockDim.x) + threadIdx.x, stride = size_t(gridDim.x) * blockDim.x;
  const unsigned world = unsigned(pt.world), all = (1u << world) - 1u;
  unsigned done = 0u, spins = 0u;
  while (done != all) {
    if (threadIdx.x == 0) {
      unsigned found = world;
      volatile uint32_t* f = pt.flags[pt.rank] + SHARD_FLAG_WALK;
      for (unsigned k = 0; k < world; ++k) {
        const unsigned r = (unsigned(pt.rank) + k) % world;
        if (!((done >> r) & 1u) && int32_t(f[r] - epoch) >= 0) { found = r; break; }
      }
      if (found != world) __threadfence_system();
      s_next = found;
    }
    __syncthreads();
    const unsigned r = s_next;
    __syncthreads();
    if (r == world) {  // nobody new yet
      __nanosleep(128);
      if (++spins > (1u << 24)) {  // ~ seconds: a peer never signalled (the host's next check reports it)
        if (threadIdx.x == 0) pt.flags[pt.rank][SHARD_TIMEOUT] = 1u;
        break;
      }
      continue;
    }
    done |= 1u << r;
    const size_t n_r = min(size_t(n_locals[r]), n_cap);
    const char* block = xacc + size_t(r) * (n_cap * 20);
    for (size_t j = j0; j < n_r; j += stride) {
      const float4 a = reinterpret_cast<const float4*>(block)[j];
      const uint32_t i = reinterpret_cast<const uint32_t*>(block + n_cap * sizeof(float4))[j];
      // one 256-bit store to a 32-byte record: a whole DRAM sector, so the scattered write needs no read-modify-write
      // of the sector (16-byte records: 1.48 ms for 33.5 M bodies; the integrator reads the first half of each record)
      if (i < n)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(acc + 2 * size_t(i)), "f"(a.x), "f"(a.y),
                     "f"(a.z), "f"(a.w), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f)
                     : "memory");
    }
  }
}

// After a full (replicated) build: the first cuts, balanced on the sorted keys, and the splitters of this
// rank's 256 buckets (quantiles of the keys in its range) for its first sharded build.
template <int DIM>
__global__ void __launch_bounds__(256) shard_plan_kernel(const uint64_t* __restrict__ sorted, size_t n, int rank,
                                                         int world, uint64_t* __restrict__ cuts /* [world + 1] */,
                                                         uint64_t* __restrict__ spl_out /* [nb + 1] */, unsigned nb) {
  constexpr int shift = DIM * (TreeDim<DIM>::LM - TopTree<DIM>::K);
  __shared__ uint64_t s_cut[16];
  __shared__ size_t s_range[2];
  const int t = threadIdx.x;
  if (t <= world) {
    uint64_t cut = 0ull;
    if (t == world) cut = ~0ull;
    else if (t > 0) cut = (sorted[(n * size_t(t)) / size_t(world)] >> shift) << shift;
    s_cut[t] = cut;
    cuts[t] = cut;
  }
  __syncthreads();
  if (t < 2) {  // first sorted body with key >= cut
    const uint64_t cut = s_cut[rank + t];
    size_t lo = 0, hi = n;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (sorted[mid] >= cut) hi = mid; else lo = mid + 1;
    }
    s_range[t] = lo;
  }
  __syncthreads();
  splitter_block(sorted + s_range[0], s_range[1] - s_range[0], spl_out, nb);
}

// ---------------------------------------------------------------------------------------------
// K9  tiled direct sum (simple_astro, and astro/astro2 with theta <= 0 where no cell is ever
//     accepted).  FP32 FMA-pipe bound: 13 FP32 lane-operations + 1 MUFU.RSQ per interaction
//     (19 flop).  Sources are staged through shared memory and broadcast to the warp; each thread
//     keeps T targets in registers.
// ---------------------------------------------------------------------------------------------
// fp32 SoA copy of the sources, [x | y | z | m | 1/m | e/m] with n_pad floats each (n_pad = n rounded
// up to a whole tile), so that a tile is a few contiguous 4 KB runs: the unit of the bulk (TMA)
// copies below.  Massless bodies and the padding get 1/m = 0, e/m = +inf: (r² + e)/m = inf, so
// their weight rsqrt(inf) is exactly 0.  Also publishes the ranges that decide whether the
// folded-mass form of the kernel is safe (see direct_kernel_x2):
//   meta[0] max |coordinate - box centre| (float bits), meta[1] 0x7fffffff - bits(min positive mass),
//   meta[2] bits(max mass), meta[3] != 0: a negative or non-finite mass exists
// Bounding box of the finite coordinates, for the fp32 copy below: the direct kernel differences fp32
// positions, so they are taken relative to the centre of the box (in fp64, BEFORE rounding to fp32) - a
// system far from the origin (`cube centre=[1000,0,0]`) would otherwise lose ~|x| * 2^-24 of every
// difference.  box[0..2] = min x,y,z, box[3..5] = max, as order-preserving u64 images of the doubles.
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long o) {
  const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double(static_cast<long long>(b));
}
constexpr unsigned long long BOX_EMPTY_MIN = ~0ull, BOX_EMPTY_MAX = 0ull;

__global__ void __launch_bounds__(256) bbox_kernel(const double4* __restrict__ pos, size_t n,
                                                   unsigned long long* __restrict__ box) {
  unsigned long long lo[3] = {BOX_EMPTY_MIN, BOX_EMPTY_MIN, BOX_EMPTY_MIN}, hi[3] = {BOX_EMPTY_MAX, BOX_EMPTY_MAX, BOX_EMPTY_MAX};
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const double4 p = pos[i];
    const double c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (fabs(c[a]) <= 1.7976931348623157e308) {  // finite
        const unsigned long long o = ordered_bits(c[a]);
        lo[a] = min(lo[a], o);
        hi[a] = max(hi[a], o);
      }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(FULL, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(FULL, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (lo[a] != BOX_EMPTY_MIN) atomicMin(&box[a], lo[a]);
      if (hi[a] != BOX_EMPTY_MAX) atomicMax(&box[3 + a], hi[a]);
    }
  }
}

// centre of the box (0 for an axis without a finite coordinate)
__device__ __forceinline__ double box_centre(const unsigned long long* __restrict__ box, int a) {
  const unsigned long long lo = box[a], hi = box[3 + a];
  if (lo == BOX_EMPTY_MIN || hi == BOX_EMPTY_MAX) return 0.0;
  return 0.5 * from_ordered_bits(lo) + 0.5 * from_ordered_bits(hi);
}

__global__ void __launch_bounds__(256) to_soa_kernel(const double4* __restrict__ pos, size_t n, size_t n_pad,
                                                     float easing, const unsigned long long* __restrict__ box,
                                                     float* __restrict__ soa, unsigned* __restrict__ meta) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  float xm = 0.f, mlo = __int_as_float(0x7f800000), mhi = 0.f;
  bool odd = false;
  const double ox = box_centre(box, 0), oy = box_centre(box, 1), oz = box_centre(box, 2);
  if (i < n_pad) {
    double4 p = make_double4(ox, oy, oz, 0.0);
    if (i < n) p = pos[i];
    const float x = float(p.x - ox), y = float(p.y - oy), z = float(p.z - oz), m = float(p.w);
    soa[i] = x;
    soa[n_pad + i] = y;
    soa[2 * n_pad + i] = z;
    soa[3 * n_pad + i] = m;
    const bool massive = m > 0.f;
    soa[4 * n_pad + i] = massive ? 1.0f / m : 0.f;
    soa[5 * n_pad + i] = massive ? easing / m : __int_as_float(0x7f800000);
    xm = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
    if (massive) { mlo = m; mhi = m; }
    odd = !(m >= 0.f) || !(m < __int_as_float(0x7f800000)) || !(xm < __int_as_float(0x7f800000));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xm = fmaxf(xm, __shfl_xor_sync(FULL, xm, o));
    mlo = fminf(mlo, __shfl_xor_sync(FULL, mlo, o));
    mhi = fmaxf(mhi, __shfl_xor_sync(FULL, mhi, o));
  }
  const bool any_odd = __any_sync(FULL, odd);
  if ((threadIdx.x & 31) == 0) {
    if (xm > 0.f) atomicMax(&meta[0], __float_as_uint(xm));
    if (mhi > 0.f) {
      atomicMax(&meta[1], 0x7fffffffu - __float_as_uint(mlo));
      atomicMax(&meta[2], __float_as_uint(mhi));
    }
    if (any_odd) meta[3] = 1u;
  }
}

constexpr int DIRECT_THREADS = 256;
constexpr int DIRECT_TILE = 1024;  // sources per shared-memory tile (16 KB)

// Blackwell's FFMA2 / FADD2 / FMUL2 (fma/add/sub/mul .f32x2) process two fp32 values per issue
// slot.  Two SOURCES are paired per instruction (the target's coordinates are a loop-invariant
// broadcast pair), so one interaction costs 6.5 FP32 issue slots + 1 MUFU instead of 13 + 1 and the
// kernel is bound by the FP32 pipe instead of by instruction issue (measured: 2.36e12 vs 2.06e12
// interactions/s for the scalar-FP32 form of the same loop).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// One tile loop of the direct sum.  FOLDED: the source mass is folded into the softened distance,
//   m / (|r| (r² + e)) = rsqrt(r² · ((r² + e)/m)²),   (r² + e)/m = fma(r², 1/m, e/m),
// which saves the multiply by m: 12 instead of 13 packed FP32 operations per source pair.
template <int T, bool FOLDED>
__device__ __forceinline__ void direct_tiles(const float* __restrict__ soa, size_t n_pad, size_t sb,
                                             unsigned n_tiles, float (*tile)[5][DIRECT_TILE], uint64_t* full,
                                             const f32x2 (&px)[T], const f32x2 (&py)[T], const f32x2 (&pz)[T],
                                             f32x2 (&ax)[T], f32x2 (&ay)[T], f32x2 (&az)[T], f32x2 e2,
                                             f32x2 tiny2, double (*wide)[DIRECT_THREADS]) {
  constexpr int NARR = FOLDED ? 5 : 4;  // x y z 1/m e/m  |  x y z m
  constexpr unsigned kTileBytes = unsigned(NARR) * DIRECT_TILE * sizeof(float);
  auto issue = [&](unsigned t) {  // tile t of this split -> buffer t & 1
    const size_t j0 = sb + size_t(t) * DIRECT_TILE;
    uint64_t* bar = &full[t & 1u];
    mbar_expect_tx(bar, kTileBytes);
#pragma unroll
    for (int q = 0; q < NARR; ++q) {
      const int arr = (FOLDED && q >= 3) ? q + 1 : q;  // folded form: arrays 4 and 5 instead of 3
      bulk_load(tile[t & 1u][q], soa + size_t(arr) * n_pad + j0, DIRECT_TILE * sizeof(float), bar);
    }
  };
  if (threadIdx.x == 0 && n_tiles > 0) issue(0);
  for (unsigned t = 0; t < n_tiles; ++t) {
    // buffer (t+1)&1 was consumed in iteration t-1; the barrier at the end of that iteration makes
    // it safe to refill now, while tile t is computed on
    if (threadIdx.x == 0 && t + 1 < n_tiles) issue(t + 1);
    mbar_wait(&full[t & 1u], (t >> 1) & 1u);
    const float* tx = tile[t & 1u][0];
    const float* ty = tile[t & 1u][1];
    const float* tz = tile[t & 1u][2];
    const float* tm = tile[t & 1u][3];  // m, or 1/m in the folded form
    const float* te = tile[t & 1u][4];  // e/m (folded form only)
#pragma unroll 2
    for (int j = 0; j < DIRECT_TILE; j += 4) {
      const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&tx[j]);  // (x0,x1) (x2,x3)
      const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&ty[j]);
      const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&tz[j]);
      const ulonglong2 M = *reinterpret_cast<const ulonglong2*>(&tm[j]);
      ulonglong2 E = M;
      if (FOLDED) E = *reinterpret_cast<const ulonglong2*>(&te[j]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const f32x2 sx = h ? X.y : X.x, sy = h ? Y.y : Y.x, sz = h ? Z.y : Z.x, sm = h ? M.y : M.x;
        const f32x2 se = h ? E.y : E.x;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const f32x2 dx = sub2(sx, px[k]);
          const f32x2 dy = sub2(sy, py[k]);
          const f32x2 dz = sub2(sz, pz[k]);
          f32x2 r2 = fma2(dx, dx, tiny2);
          r2 = fma2(dy, dy, r2);
          r2 = fma2(dz, dz, r2);
          f32x2 mw;
          if (FOLDED) {
            const f32x2 sft = fma2(r2, sm, se);  // (r² + e) / m
            const f32x2 u = mul2(mul2(r2, sft), sft);
            float u0, u1;
            unpack2(u, u0, u1);
            mw = pack2(rsqrt_approx(u0), rsqrt_approx(u1));
          } else {
            const f32x2 sft = add2(r2, e2);
            const f32x2 u = mul2(mul2(r2, sft), sft);
            float u0, u1;
            unpack2(u, u0, u1);
            mw = mul2(sm, pack2(rsqrt_approx(u0), rsqrt_approx(u1)));
          }
          ax[k] = fma2(mw, dx, ax[k]);
          ay[k] = fma2(mw, dy, ay[k]);
          az[k] = fma2(mw, dz, az[k]);
        }
      }
    }
    // The fp32 accumulators only ever hold ONE tile's terms (512 per packed lane); the tiles are added in fp64.
    // At 2^24 sources a single fp32 running sum per lane drifts to ~1e-3 of the (heavily cancelling) net force;
    // per-tile flushing keeps the sum at fp32-term accuracy for 24 DADD per 13 k packed FP32 instructions.
#pragma unroll
    for (int k = 0; k < T; ++k) {
      float a0, a1, b0, b1, c0, c1;
      unpack2(ax[k], a0, a1);
      unpack2(ay[k], b0, b1);
      unpack2(az[k], c0, c1);
      wide[3 * k + 0][threadIdx.x] += double(a0) + double(a1);
      wide[3 * k + 1][threadIdx.x] += double(b0) + double(b1);
      wide[3 * k + 2][threadIdx.x] += double(c0) + double(c1);
      ax[k] = ay[k] = az[k] = pack2(0.f, 0.f);
    }
    __syncthreads();  // everyone is done with buffer t&1 before it is refilled
  }
}

template <int T>
__global__ void __launch_bounds__(DIRECT_THREADS) direct_kernel_x2(
    const float* __restrict__ soa, size_t n_pad, size_t src_per_split, size_t t0, size_t n_targets,
    float easing, float tiny, const unsigned* __restrict__ meta, float4* __restrict__ part /* [splits][n_targets] */) {
  // Double-buffered SoA source tiles, filled by one elected thread with 4 KB bulk copies (TMA) that
  // complete on an mbarrier while the previous tile is being consumed.  Consecutive sources are
  // adjacent, so an aligned 16-byte shared read yields two source pairs.
  __shared__ __align__(128) float tile[2][5][DIRECT_TILE];
  __shared__ __align__(8) uint64_t full[2];
  // fp64 sums over the tiles, private to each thread (no register cost); dynamic: static shared memory ends at 48 KB
  extern __shared__ __align__(16) unsigned char direct_dyn[];
  double (*wide)[DIRECT_THREADS] = reinterpret_cast<double (*)[DIRECT_THREADS]>(direct_dyn);
  const float* gx = soa;
  const float* gy = soa + n_pad;
  const float* gz = soa + 2 * n_pad;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t sb = size_t(blockIdx.y) * src_per_split;
  const size_t se = min(sb + src_per_split, n_pad);
  const unsigned n_tiles = static_cast<unsigned>((se - sb) / DIRECT_TILE);  // whole tiles by construction

  // The folded-mass form is used when every u = r²((r²+e)/m)² it can meet stays a normal fp32 number
  // (ranges published by to_soa_kernel; the same verdict in every CTA): largest r² with the smallest
  // mass must not overflow, r² = tiny with the largest mass must not underflow.
  bool folded = meta[3] == 0u && meta[2] != 0u;
  if (folded) {
    const float xmax = __uint_as_float(meta[0]);
    const float mmin = __uint_as_float(0x7fffffffu - meta[1]), mmax = __uint_as_float(meta[2]);
    const float r2max = 12.f * xmax * xmax + tiny;
    const float hi = (r2max + easing) / mmin, lo = (tiny + easing) / mmax;
    folded = hi < 1e18f && r2max < 1e-2f * (1e36f / (hi * hi)) && lo > 1e-18f && tiny * (lo * lo) > 1e-36f;
  }

  const size_t tbase = size_t(blockIdx.x) * (DIRECT_THREADS * T);
  f32x2 px[T], py[T], pz[T], ax[T], ay[T], az[T];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const size_t lt = tbase + size_t(k) * DIRECT_THREADS + threadIdx.x;
    const size_t g = lt < n_targets ? t0 + lt : 0;
    const float x = gx[g], y = gy[g], z = gz[g];
    px[k] = pack2(x, x);
    py[k] = pack2(y, y);
    pz[k] = pack2(z, z);
    ax[k] = ay[k] = az[k] = pack2(0.f, 0.f);
    wide[3 * k + 0][threadIdx.x] = wide[3 * k + 1][threadIdx.x] = wide[3 * k + 2][threadIdx.x] = 0.0;
  }
  const f32x2 e2 = pack2(easing, easing), tiny2 = pack2(tiny, tiny);
  if (folded) direct_tiles<T, true>(soa, n_pad, sb, n_tiles, tile, full, px, py, pz, ax, ay, az, e2, tiny2, wide);
  else direct_tiles<T, false>(soa, n_pad, sb, n_tiles, tile, full, px, py, pz, ax, ay, az, e2, tiny2, wide);
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const size_t lt = tbase + size_t(k) * DIRECT_THREADS + threadIdx.x;
    if (lt < n_targets)
      part[size_t(blockIdx.y) * n_targets + lt] =
          make_float4(float(wide[3 * k + 0][threadIdx.x]), float(wide[3 * k + 1][threadIdx.x]),
                      float(wide[3 * k + 2][threadIdx.x]), 0.f);
  }
}

// Direct sum for small systems (n <= DIRECT_SMALL_N: the few-body configurations - solar.toml has 67
// bodies, moons 0.01 from their planets): p_b - p_a is taken in fp64 like the reference and like the tree
// walk, so close pairs keep their direction however small the separation is relative to the coordinates;
// the force law stays fp32 (one MUFU.RSQ per pair), the terms are summed in fp64.  One thread per target, sources staged through shared
// memory as the fp64 {x,y,z,m} records themselves (no fp32 copy).  At these sizes the FP64 subtractions
// cost microseconds; the packed-FP32 kernel above is the one that matters from ~10^5 bodies up.
constexpr size_t DIRECT_SMALL_N = 32768;
constexpr int DIRECT_SMALL_THREADS = 128;

__global__ void __launch_bounds__(DIRECT_SMALL_THREADS) direct_small_kernel(
    const double4* __restrict__ pos, const uint8_t* __restrict__ fixed, size_t n, size_t t0, size_t n_targets,
    float easing, float tiny, float4* __restrict__ acc) {
  __shared__ double4 tile[DIRECT_SMALL_THREADS];
  const size_t lt = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const bool live = lt < n_targets;
  const size_t i = live ? t0 + lt : 0;
  const double4 me = pos[i < n ? i : 0];
  // the per-pair terms are fp32; they are ADDED in fp64 (2000 fp32 additions next to one dominant partner
  // term lose ~1e-5 of it; negligible cost at these sizes)
  double ax = 0.0, ay = 0.0, az = 0.0;
  for (size_t j0 = 0; j0 < n; j0 += DIRECT_SMALL_THREADS) {
    const size_t j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < n ? pos[j] : make_double4(me.x, me.y, me.z, 0.0);
    __syncthreads();
    const int cnt = int(min(size_t(DIRECT_SMALL_THREADS), n - j0));
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) {
      const double4 sb = tile[q];
      const float dx = static_cast<float>(sb.x - me.x);
      const float dy = static_cast<float>(sb.y - me.y);
      const float dz = static_cast<float>(sb.z - me.z);
      float r2 = fmaf(dx, dx, tiny);  // tiny keeps r = 0 pairs (self / coincident: skipped by the reference) at w * 0 = 0
      r2 = fmaf(dy, dy, r2);
      r2 = fmaf(dz, dz, r2);
      const float sft = r2 + easing;
      const float mw = static_cast<float>(sb.w) * rsqrt_approx(r2 * sft * sft);
      ax += double(mw * dx);
      ay += double(mw * dy);
      az += double(mw * dz);
    }
    __syncthreads();
  }
  if (!live) return;
  float fx = float(ax), fy = float(ay), fz = float(az);
  uint32_t inter = static_cast<uint32_t>(n);
  if (fixed[i]) {
    fx = fy = fz = 0.f;
    inter = 0;
  } else if (me.w == 0.0) {
    fx = fy = fz = __int_as_float(0x7fc00000);
  }
  acc[i] = make_float4(fx, fy, fz, __uint_as_float(inter));
}

// sums the source splits in a fixed order and applies the per-target rules
__global__ void __launch_bounds__(256) direct_finish_kernel(const float4* __restrict__ part, int splits,
                                                            size_t t0, size_t n_targets,
                                                            const double4* __restrict__ pos,
                                                            const uint8_t* __restrict__ fixed,
                                                            uint32_t n_src, float4* __restrict__ acc) {
  const size_t lt = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (lt >= n_targets) return;
  float x = 0.f, y = 0.f, z = 0.f;
  for (int sidx = 0; sidx < splits; ++sidx) {
    const float4 p = part[size_t(sidx) * n_targets + lt];
    x += p.x; y += p.y; z += p.z;
  }
  const size_t i = t0 + lt;
  uint32_t inter = n_src;
  if (fixed[i]) {
    x = y = z = 0.f;
    inter = 0;
  } else if (pos[i].w == 0.0) {
    x = y = z = __int_as_float(0x7fc00000);
  }
  acc[i] = make_float4(x, y, z, __uint_as_float(inter));
}

__global__ void __launch_bounds__(256) count_kernel(const float4* __restrict__ acc, size_t n,
                                                    unsigned long long* __restrict__ out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x)
    s += __float_as_uint(acc[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ---------------------------------------------------------------------------------------------
// FP32 FFMA issue-rate probe (roofline denominator check)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

inline unsigned blocks_for(size_t n, int per_block) {
  return static_cast<unsigned>((n + per_block - 1) / per_block);
}

// u32 exclusive scan: in[n] -> out[n+1]
cudaError_t exclusive_scan(const uint32_t* in, uint32_t* out, size_t n, DevBuf& tmp, cudaStream_t st,
                           LaunchStats& ls) {
  const unsigned tiles = blocks_for(n, SCAN_TILE);
  PB_PASS(tmp.ensure(size_t(tiles) * 8 + 16));  // [tile status words][tile counter]
  PB_CUDA(cudaMemsetAsync(tmp.p, 0, size_t(tiles) * 8 + 16, st));
  PB_LAUNCH(ls, st, "scan_lookback_kernel",
            scan_lookback_kernel<false><<<tiles, 256, 0, st>>>(in, NRef{nullptr, uint32_t(n)}, out, tmp.as<unsigned long long>(),
                                                               reinterpret_cast<unsigned*>(tmp.as<unsigned long long>() + tiles),
                                                               ScanSide{}));
  return cudaGetLastError();
}

// the scan of the build's cell counts, with the side jobs; `scratch` ([tiles status words][counter]) is
// part of the build's zeroed scratch block
inline size_t scan_scratch_bytes(size_t n) { return size_t(blocks_for(n, SCAN_TILE)) * 8 + 16; }
// n_cap: the host's bound on the body count (== the count itself unless nref.dev is set); sizes the grid and
// places the ticket counter behind n_cap's status words
cudaError_t exclusive_scan_with_side(const uint32_t* in, uint32_t* out, NRef nref, size_t n_cap, unsigned long long* scratch,
                                     const ScanSide& side, cudaStream_t st, LaunchStats& ls) {
  const unsigned tiles = blocks_for(n_cap, SCAN_TILE);
  const unsigned nsuper = blocks_for(blocks_for(n_cap, 256), 256);
  PB_LAUNCH(ls, st, "scan_lookback_kernel",
            pb_launch_pdl(scan_lookback_kernel<true>, dim3(tiles + nsuper + 1), dim3(256), 0, st,
                in, nref, out, scratch, reinterpret_cast<unsigned*>(scratch + tiles), side));
  return cudaGetLastError();
}

struct SortBuffers {
  SortPlan plan;   // global LSD passes (none in the bucket modes)
  int mode;        // 0: global LSD passes, 1 / 2 / 3: bucket sort with 4608 / 8192 / 16384-key buckets
  int lo, key_bits;
  unsigned nb;     // buckets in the bucket modes (256 / 512 / 1024)
  unsigned cap;    // bucket capacity in the bucket modes
  unsigned tiles;
  int items;
  unsigned *ghist, *counters, *err_flag, *status;
};

// bucket capacities of modes 1..3: two CTAs of 512 threads per SM, then one CTA of 1024 threads per SM
constexpr unsigned LOCAL_CAP[4] = {0u, 4608u, 8192u, 16384u};

inline SortPlan even_plan(int lo, int total) {
  SortPlan plan;
  plan.npass = (total + 7) / 8;
  plan.lo = lo;
  plan.base = plan.npass ? total / plan.npass : 0;
  plan.rem = plan.npass ? total % plan.npass : 0;
  return plan;
}

// plans the sort of key bits [lo, key_bits) and clears histograms / look-back state
constexpr size_t SORT_HEAD_WORDS = SORT_HIST_SLOTS * 256 + SORT_MAX_PASSES + 8;
// Buckets and capacity for n bodies: the fewest buckets whose share (n / nb + 12 % headroom: bodies drift between
// evaluations) fits a shared-memory tile, the smallest tile that takes it.  mode 0: too many bodies for 1024 tiles.
struct BucketPlan {
  unsigned nb;
  int mode;
};
inline BucketPlan bucket_plan(size_t n) {
  for (unsigned nb = 256; nb <= MAX_BUCKETS; nb *= 2) {
    const size_t want = n / nb + n / (8 * size_t(nb)) + 64;
    for (int m = 1; m <= 3; ++m)
      if (want <= LOCAL_CAP[m] && (m >= 2 || nb <= 512)) return BucketPlan{nb, m};  // (mode 1: 512-thread CTAs scan <= 512 cursors)
  }
  return BucketPlan{256, 0};
}

cudaError_t sort_prepare(GravityWorkspace& ws, size_t n, int key_bits, int lo, int mode, unsigned nb, cudaStream_t st,
                         unsigned* head /* SORT_HEAD_WORDS zeroed words */, SortBuffers* sb) {
  // keys per thread, measured on B200: 8 wins at 1e5 bodies and from 4e6 up (more CTAs in flight),
  // 16 at 1e6 (one wave of 245 CTAs)
  static const int items_env = std::getenv("PB200_SORT_ITEMS") ? std::atoi(std::getenv("PB200_SORT_ITEMS")) : 0;
  sb->items = items_env ? items_env : ((n <= (size_t(1) << 19) || n >= (size_t(1) << 21)) ? 8 : 16);
  sb->tiles = blocks_for(n, SORT_THREADS * sb->items);
  if (key_bits - lo < 8) mode = 0;
  sb->mode = mode;
  sb->nb = nb;
  sb->lo = lo;
  sb->key_bits = key_bits;
  sb->cap = LOCAL_CAP[mode];
  sb->plan = even_plan(lo, mode == 0 ? key_bits - lo : 0);
  const SortPlan& plan = sb->plan;
  // head (zeroed by the caller): [ghist: 8 x 256 (bucket modes: slot 0 = bucket cursors)][tile counters: 8]
  // [error flag + pad: 8]; look-back status of the global passes: npass x tiles x 256 words
  sb->ghist = head;
  sb->counters = sb->ghist + SORT_HIST_SLOTS * 256;
  sb->err_flag = sb->counters + SORT_MAX_PASSES;
  sb->status = nullptr;
  if (plan.npass) {
    const size_t words = size_t(plan.npass) * sb->tiles * 256;
    PB_PASS(ws.tile_counts.ensure(words * 4));
    sb->status = ws.tile_counts.as<unsigned>();
    PB_CUDA(cudaMemsetAsync(sb->status, 0, words * 4, st));
  }
  if (ws.sticky.p) sb->err_flag = ws.sticky.as<unsigned>() + 2;  // survives until the next host check
  ws.sort_err_flag = sb->err_flag;
  if (mode != 0) {
    PB_PASS(ws.bucket_key.ensure(size_t(nb) * sb->cap * 8));
    PB_PASS(ws.bucket_idx.ensure(size_t(nb) * sb->cap * 4));
  }
  if (!ws.splitters.p) {
    PB_PASS(ws.splitters.ensure(2 * SPLITTER_STRIDE * 8));
    PB_CUDA(cudaMemsetAsync(ws.splitters.p, 0, 2 * SPLITTER_STRIDE * 8, st));
  }
  return cudaSuccess;
}

// keys -> sorted keys + permutation (ws.sorted_key / ws.perm) + sorted {x,y,z,m} (ws.spos64).
// `bad` is the build's "keys not ordered" flag.
inline int shard_sort_mode(size_t n_cap) { return bucket_plan(n_cap).mode; }

// n: every body of ws.pos64.  sh != nullptr: only the bodies in this rank's key range are kept (bucket forms only)
template <int DIM>
cudaError_t encode_and_sort(GravityWorkspace& ws, size_t n, const SortBuffers& sb, unsigned* bad, cudaStream_t st,
                            LaunchStats& ls, uint64_t** splitters_out, const ShardBuild* sh) {
  uint64_t* k[2] = {ws.key0.as<uint64_t>(), ws.key1.as<uint64_t>()};
  uint32_t* v[2] = {(sh && sh->perm_out) ? sh->perm_out : ws.idx0.as<uint32_t>(), ws.idx1.as<uint32_t>()};
  unsigned* stat_max = ws.sticky.as<unsigned>() + 3;
  const unsigned nb = blocks_for(n, 256);
  // splitters: read the set the previous evaluation left, write the other one
  const uint64_t* spl_in = ws.splitters.as<uint64_t>() + SPLITTER_STRIDE * ws.splitter_cur;
  uint64_t* spl_out = ws.splitters.as<uint64_t>() + SPLITTER_STRIDE * (ws.splitter_cur ^ 1);
  ws.splitter_cur ^= 1;
  *splitters_out = spl_out;  // written by the side job of the scan that follows the sort
  // (Sharded build: every rank computes the keys of ALL bodies itself and keeps those in its range.  Measured,
  // r02: computing the keys of index slice r on rank r and storing them into every rank's key array - worth it
  // while a key cost a 21 / 31-step fp64 chain - loses to recomputing the quantised keys: 0.430 vs 0.397 ms per
  // step at 2 x 1 M bodies, and 8 B x n x (G-1)/G of NVLink stores per rank less.)
  if (sb.mode != 0) {
#define PB_ENCODE(NBV)                                                                                            \
  PB_LAUNCH(ls, st, "encode_bucket_kernel",                                                                      \
            pb_launch_pdl(encode_bucket_kernel<DIM, NBV>, dim3(blocks_for(n, 256 * ENC_ITEMS)), dim3(256), 0, st, \
                          ws.pos64, n, ws.extent_cur, spl_in, sb.lo, ws.bucket_key.as<uint64_t>(),              \
                          ws.bucket_idx.as<uint32_t>(), sb.cap, sb.ghist, sh ? sh->cuts : nullptr))
    if (sb.nb == 256) PB_ENCODE(256u);
    else if (sb.nb == 512) PB_ENCODE(512u);
    else PB_ENCODE(1024u);
#undef PB_ENCODE
    uint32_t* n_out = sh ? sh->n_local : nullptr;
    const uint32_t n_cap = sh ? uint32_t(sh->n_cap) : 0u;
    const size_t smem = sort_local_smem(sb.cap);
    if (sb.mode == 1) {
      PB_CUDA(cudaFuncSetAttribute(sort_local_kernel<512, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      PB_LAUNCH(ls, st, "sort_local_kernel",
                pb_launch_pdl(sort_local_kernel<512, 9>, dim3(sb.nb), dim3(512), smem, st, 
                    ws.bucket_key.as<uint64_t>(), ws.bucket_idx.as<uint32_t>(), sb.ghist, spl_in, sb.cap, sb.lo,
                    sb.key_bits, k[0], v[0], ws.pos64, ws.spos64.as<double4>(), bad, stat_max, n_out, n_cap, sb.nb));
    } else {
      PB_CUDA(cudaFuncSetAttribute(sort_local_kernel<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      PB_LAUNCH(ls, st, "sort_local_kernel",
                pb_launch_pdl(sort_local_kernel<1024, 8>, dim3(sb.nb), dim3(1024), smem, st, 
                    ws.bucket_key.as<uint64_t>(), ws.bucket_idx.as<uint32_t>(), sb.ghist, spl_in, sb.cap, sb.lo,
                    sb.key_bits, k[0], v[0], ws.pos64, ws.spos64.as<double4>(), bad, stat_max, n_out, n_cap, sb.nb));
    }
    ws.sorted_key = k[0];
    ws.perm = v[0];
    return cudaGetLastError();
  }
  PB_LAUNCH(ls, st, "encode_kernel",
            encode_kernel<DIM><<<min(nb, 148u * 4u), 256, 0, st>>>(
                ws.pos64, n, ws.extent_cur, k[0], v[0], sb.plan, sb.ghist));
  // 48 KB of dynamic + 10 KB of static shared memory for the 16-keys-per-thread tile: opt in
  // (per device, so per call: the handle may live on any device)
  if (sb.items == 16)
    PB_CUDA(cudaFuncSetAttribute(sort_onesweep_pass<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SORT_THREADS * 16 * 12));
  int cur = 0;
  for (int p = 0; p < sb.plan.npass; ++p) {
    const size_t smem = size_t(SORT_THREADS) * sb.items * 12;  // digit-sorted tile: u64 keys + u32 values
    if (sb.items == 8)
      PB_LAUNCH(ls, st, "sort_onesweep_pass",
                sort_onesweep_pass<8><<<sb.tiles, SORT_THREADS, smem, st>>>(
                    k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, sb.plan.shift(p), sb.plan.mask(p),
                    sb.ghist + p * 256, sb.status + size_t(p) * sb.tiles * 256, sb.counters + p, sb.err_flag));
    else
      PB_LAUNCH(ls, st, "sort_onesweep_pass",
                sort_onesweep_pass<16><<<sb.tiles, SORT_THREADS, smem, st>>>(
                    k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, sb.plan.shift(p), sb.plan.mask(p),
                    sb.ghist + p * 256, sb.status + size_t(p) * sb.tiles * 256, sb.counters + p, sb.err_flag));
    cur ^= 1;
  }
  ws.sorted_key = k[cur];
  ws.perm = v[cur];
  PB_LAUNCH(ls, st, "gather_kernel", gather_kernel<<<nb, 256, 0, st>>>(ws.pos64, ws.perm, n, ws.spos64.as<double4>()));
  return cudaGetLastError();
}

// keys -> sort -> units -> scan -> cell table -> centres of mass, for all ws.n bodies or (sh != nullptr)
// for those in this rank's key range
template <int DIM>
cudaError_t tree_build(GravityWorkspace& ws, const ShardBuild* sh, cudaStream_t st, LaunchStats& ls, BuildOut* out) {
  const size_t n_all = ws.n;
  const size_t n = sh ? sh->n_cap : n_all;  // capacity of everything indexed by sorted body
  const NRef nref{sh ? sh->n_local : nullptr, uint32_t(n)};
  // one zeroed scratch block per build: [extent / flags: 32 B][scan status + counter][sort head][climb-start counters]
  const size_t scan_bytes = (scan_scratch_bytes(n) + 15) / 16 * 16;
  const size_t ready_off = (32 + scan_bytes + SORT_HEAD_WORDS * 4 + 127) / 128 * 128;
  const size_t scratch_bytes = ready_off + size_t(READY_LISTS) * READY_STRIDE * 4;
  PB_PASS(ws.extent_bits.ensure(scratch_bytes));
  if (!ws.sticky.p) {
    PB_PASS(ws.sticky.ensure(32));
    PB_CUDA(cudaMemsetAsync(ws.sticky.p, 0, 32, st));
  }
  PB_PASS(ws.key0.ensure(n * 8 + 16));
  PB_PASS(ws.key1.ensure(n * 8));
  PB_PASS(ws.idx0.ensure(n * 4));
  PB_PASS(ws.idx1.ensure(n * 4));
  PB_PASS(ws.spos64.ensure(n * sizeof(double4)));
  PB_PASS(ws.ab.ensure(n * 2));
  PB_PASS(ws.cell_start.ensure((n + 1) * 4));
  PB_PASS(ws.tgt_flags.ensure(n * 4));  // also the per-body cell counts before the scan

  // extent_bits words: [0,1] extent (u64 bits)  [2] 1 + deepest level shared by distinct keys
  // [3] some leaf is a merged unit  [4] keys not fully ordered (truncated sort too short)
  PB_CUDA(cudaMemsetAsync(ws.extent_bits.p, 0, scratch_bytes, st));
  unsigned long long* scan_scratch = ws.extent_bits.as<unsigned long long>() + 4;
  unsigned* sort_head = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws.extent_bits.p) + 32 + scan_bytes);
  unsigned* n_ready = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws.extent_bits.p) + ready_off);
  unsigned* max_shared_plus1 = reinterpret_cast<unsigned*>(ws.extent_bits.as<unsigned long long>() + 1);
  const int key_bits = DIM * TreeDim<DIM>::LM;
  if (ws.tree_dim != DIM) ws.sort_lo = 0, ws.sort_mode = 0, ws.bucket_min_mode = 1;  // depth / bucket estimates belong to the other tree kind
  ws.tree_dim = DIM;
  // Only the global LSD passes get cheaper with fewer key bits (fewer passes).  The bucket sort ranks the
  // members of a bin by comparing whole keys, so it sorts ALL bits at no extra cost - and a depth guess
  // that can no longer be wrong cannot make a chunk of the resident loop replay (one replay per ~150
  // steps of c3 while only the guessed bits were sorted: tools/steps_diag.py).
  const int lo = (!sh && ws.sort_mode == 0 && ws.sort_lo > 0 && ws.sort_lo < key_bits) ? ws.sort_lo : 0;
  ws.last_lo = lo;
  ws.unchecked_builds += 1;
  const unsigned nb = blocks_for(n, 256);
  // the extent of pos64: left on the device by the resident verlet step, else reduced here
  const unsigned long long* extent = ws.extent_pre;
  ws.extent_pre = nullptr;  // (good for one build)
  if (!extent) {
    const unsigned nb_all = blocks_for(n_all, 256);
    PB_LAUNCH(ls, st, "extent_kernel", pb_launch_pdl(extent_kernel, dim3(min(nb_all, 148u * 8u)), dim3(256), 0, st, ws.pos64, n_all, ws.extent_bits.as<unsigned long long>()));
    extent = ws.extent_bits.as<unsigned long long>();
  }
  ws.extent_cur = extent;
  SortBuffers sb;
  const BucketPlan bp = bucket_plan(n);
  int mode = sh ? (bp.mode ? std::max(bp.mode, ws.bucket_min_mode) : 0) : (bp.mode == 0 ? 0 : ws.sort_mode);
  // the splitter set this build reads must have been written for the same number of buckets
  if (mode != 0 && ws.spl_nb[ws.splitter_cur] != bp.nb) {
    if (sh) {
      set_error("sharded build: no splitters for %u buckets (gravity_shard_plan first)", bp.nb);
      return cudaErrorInvalidValue;
    }
    mode = 0;
  }
  PB_PASS(sort_prepare(ws, n, key_bits, lo, mode, bp.nb, st, sort_head, &sb));
  if (sh && sb.mode == 0) {
    set_error("sharded build: %zu bodies per rank exceed the bucket sort's capacity", n);
    return cudaErrorInvalidValue;
  }
  ws.last_mode = sb.mode;
  uint64_t* spl_out = nullptr;
  PB_PASS(encode_and_sort<DIM>(ws, n_all, sb, max_shared_plus1 + 2, st, ls, &spl_out, sh));
  // range-minimum tables over the sorted bodies' shared-level bytes (see NsvTables)
  const size_t n_pad = size_t(nb) * 256, nblocks = nb, b_pad = (nblocks + 255) / 256 * 256, nsuper = b_pad / 256;
  PB_PASS(ws.nsv1.ensure(n_pad + n_pad / 16));  // [a1 bytes][window minima]
  PB_PASS(ws.nsv2.ensure(9 * b_pad + nsuper));
  uint8_t* nsv3 = ws.nsv2.as<uint8_t>() + 9 * b_pad;
  PB_LAUNCH(ls, st, "unit_kernel", pb_launch_pdl(unit_kernel<DIM>, dim3(min(nb, 148u * unsigned(UNIT_CTAS_PER_SM))), dim3(256), 0, st, ws.sorted_key, ws.spos64.as<double4>(), nref, ws.ab.as<uchar2>(),
                                       ws.tgt_flags.as<uint32_t>(), max_shared_plus1,
                                       lo > 0 ? (key_bits - lo) / DIM : TreeDim<DIM>::LM + 2,
                                       ws.nsv1.as<uint8_t>(), n_pad, ws.nsv2.as<uint8_t>()));
  // (encode_and_sort toggled splitter_cur: [cur] is the set the scan's side job writes, [cur ^ 1] the one this build read)
  const uint64_t* spl_read = ws.splitters.as<uint64_t>() + SPLITTER_STRIDE * (ws.splitter_cur ^ 1);
  const ScanSide side{ws.nsv2.as<uint8_t>(), b_pad, nsv3, ws.sorted_key, spl_out, bp.nb,
                      (sb.mode != 0 && ws.spl_nb[ws.splitter_cur ^ 1] == bp.nb) ? spl_read : nullptr, max_shared_plus1 + 2};
  ws.spl_nb[ws.splitter_cur] = bp.nb;  // (encode_and_sort toggled splitter_cur: this is the set the scan's side job writes)
  PB_PASS(exclusive_scan_with_side(ws.tgt_flags.as<uint32_t>(), ws.cell_start.as<uint32_t>(), nref, n, scan_scratch, side, st, ls));

  // cell table capacity: grows when a previous evaluation reported more cells
  size_t cap = ws.n_cells ? ws.n_cells + ws.n_cells / 4 + 1024 : n * 5 / 2 + 1024;
  if (cap < ws.cell_cap) cap = ws.cell_cap;  // never shrink
  if (cap > 0xfffffff0ull) cap = 0xfffffff0ull;
  ws.cell_cap = cap;
  PB_PASS(ws.c_level.ensure(cap + 16));
  PB_PASS(ws.c_head.ensure(cap * 4));
  PB_PASS(ws.c_count.ensure(cap * 4));
  PB_PASS(ws.c_skip.ensure(cap * 4));
  PB_PASS(ws.c_parent.ensure(cap * 4));
  PB_PASS(ws.c_arrived.ensure(cap * 4));
  PB_PASS(ws.c_centre_ext.ensure(cap * sizeof(double4)));
  PB_PASS(ws.c_com.ensure(cap * sizeof(double4)));
  static const uint32_t small_cell =
      std::getenv("PB200_SMALL_CELL") ? uint32_t(std::atoi(std::getenv("PB200_SMALL_CELL"))) : SMALL_CELL;  // (tuning runs)
  CellArrays cells{ws.c_level.as<uint8_t>(),   ws.c_head.as<uint32_t>(),  ws.c_count.as<uint32_t>(),
                   ws.c_skip.as<uint32_t>(),   ws.c_parent.as<uint32_t>(), ws.c_arrived.as<uint32_t>(),
                   ws.c_centre_ext.as<double4>(), ws.c_com.as<double4>(), static_cast<uint32_t>(cap), small_cell,
                   max_shared_plus1 + 2, uint32_t(TopTree<DIM>::K)};
  const NsvTables tv{ws.nsv1.as<uint8_t>(), ws.nsv1.as<uint8_t>() + n_pad, n_pad, ws.nsv2.as<uint8_t>(), b_pad, nsv3, n, nblocks, nsuper};
  const TopSlots slots = sh ? sh->slots : TopSlots{nullptr, 0};
  // Resident CTAs per SM the compiler makes room for: 5 (48 registers, no spills; 62 % of the warp slots) beats 4
  // (64 registers) by 2.5 us at c3 and 15 us at 4.2 M bodies - the kernel waits on dependent loads and on divergent
  // paths, more warps hide them; 6 (40 registers) spills and loses 6 us.  PB200_CELLS_BLOCKS=4: the earlier build.
  static const int cells_blocks = std::getenv("PB200_CELLS_BLOCKS") ? std::atoi(std::getenv("PB200_CELLS_BLOCKS")) : 5;  // (tuning runs)
#define PB_CELLS_LAUNCH(MB)                                                                                          \
  PB_LAUNCH(ls, st, "cells_kernel",                                                                                  \
            (pb_launch_pdl(cells_kernel<DIM, MB>, dim3(nb), dim3(256), 0, st, ws.sorted_key, ws.spos64.as<double4>(), \
                           ws.perm, ws.ab.as<uchar2>(), ws.cell_start.as<uint32_t>(), nref, ws.extent_cur,           \
                           max_shared_plus1, ws.sticky.as<unsigned>(), tv, cells, slots)))
  if (cells_blocks == 4) PB_CELLS_LAUNCH(4);
  else PB_CELLS_LAUNCH(5);
#undef PB_CELLS_LAUNCH
  ws.parents_filled = false;
  {
    PB_PASS(ws.c_kids.ensure(cap * (size_t(4) << DIM)));
    // (a warp of 32 cells holds at most 32 starts; list l takes the warps = l mod READY_LISTS)
    const unsigned kid_blocks = blocks_for(cap, 256);
    const uint32_t sub_cap = ((kid_blocks + READY_LISTS - 1) / READY_LISTS) * 256u;
    PB_PASS(ws.c_ready.ensure(size_t(READY_LISTS) * sub_cap * 4));
    // the grid follows the cells expected for THIS build's bodies (the table may be sized for a full build)
    const unsigned kid_grid = std::min(kid_blocks, blocks_for(n * 5 / 2 + 1024, 256));
    PB_LAUNCH(ls, st, "kids_kernel",
              pb_launch_pdl(kids_kernel<DIM>, dim3(kid_grid), dim3(256), 0, st, ws.cell_start.as<uint32_t>(), nref, cells,
                                                           ws.c_kids.as<uint32_t>(), ws.c_ready.as<uint32_t>(), n_ready, sub_cap));
    PB_LAUNCH(ls, st, "climb_kernel",
              pb_launch_pdl(climb_kernel<DIM>, dim3(148 * 4), dim3(128), 0, st, ws.cell_start.as<uint32_t>(), nref, cells,
                                                         ws.c_kids.as<uint32_t>(), ws.c_ready.as<uint32_t>(), n_ready, sub_cap));
  }
  out->cells = cells;
  out->nref = nref;
  out->n_cap = n;
  return cudaGetLastError();
}

template <int DIM>
cudaError_t tree_evaluate(GravityWorkspace& ws, const GravityParams& prm, size_t t0, size_t t1,
                          float easing, float tiny, cudaStream_t st, LaunchStats& ls) {
  const size_t n = ws.n;
  BuildOut bo;
  PB_PASS(tree_build<DIM>(ws, nullptr, st, ls, &bo));
  const CellArrays& cells = bo.cells;
  const unsigned nb = blocks_for(n, 256);
  const uint32_t* list = nullptr;
  const size_t n_targets = t1 - t0;
  if (n_targets != n) {
    PB_PASS(ws.tgt_list.ensure((n + 1) * 4 + n_targets * 4));
    uint32_t* offs = ws.tgt_list.as<uint32_t>();
    uint32_t* lst = offs + (n + 1);
    PB_LAUNCH(ls, st, "tgt_flag_kernel", tgt_flag_kernel<<<nb, 256, 0, st>>>(ws.perm, n, uint32_t(t0), uint32_t(t1), ws.tgt_flags.as<uint32_t>()));
    PB_PASS(exclusive_scan(ws.tgt_flags.as<uint32_t>(), offs, n, ws.scan_tmp, st, ls));
    PB_LAUNCH(ls, st, "tgt_scatter_kernel", tgt_scatter_kernel<<<nb, 256, 0, st>>>(ws.tgt_flags.as<uint32_t>(), offs, n, lst));
    list = lst;
  }
  if (n_targets) {
    PB_LAUNCH(ls, st, "walk_kernel", pb_launch_pdl(walk_kernel<DIM>, dim3(blocks_for(n_targets, 256)), dim3(256), 0, st,
        ws.spos64.as<double4>(), ws.perm, ws.fixed, list, n_targets, ws.cell_start.as<uint32_t>(), n,
        cells, ws.extent_cur, prm.theta, easing, tiny, ws.acc.as<float4>()));
  }
  return cudaGetLastError();
}

cudaError_t direct_evaluate(GravityWorkspace& ws, size_t t0, size_t t1, float easing, float tiny,
                            cudaStream_t st, LaunchStats& ls) {
  const size_t n = ws.n, n_targets = t1 - t0;
  const char* small_env = std::getenv("PB200_DIRECT_SMALL");  // =0: the packed kernel at every size (read per call: tests flip it)
  const bool small_off = small_env && std::atoi(small_env) == 0;
  if (n <= DIRECT_SMALL_N && !small_off) {
    if (n_targets)
      PB_LAUNCH(ls, st, "direct_small_kernel",
                direct_small_kernel<<<blocks_for(n_targets, DIRECT_SMALL_THREADS), DIRECT_SMALL_THREADS, 0, st>>>(
                    ws.pos64, ws.fixed, n, t0, n_targets, easing, tiny, ws.acc.as<float4>()));
    return cudaGetLastError();
  }
  const size_t n_pad = (n + DIRECT_TILE - 1) / DIRECT_TILE * DIRECT_TILE;
  PB_PASS(ws.src4.ensure(6 * n_pad * sizeof(float)));
  PB_PASS(ws.counters.ensure(16 + 48));
  unsigned* meta = ws.counters.as<unsigned>();
  unsigned long long* box = reinterpret_cast<unsigned long long*>(meta + 4);  // [min x,y,z][max x,y,z]
  PB_CUDA(cudaMemsetAsync(meta, 0, 16 + 48, st));
  PB_CUDA(cudaMemsetAsync(box, 0xff, 24, st));
  PB_LAUNCH(ls, st, "bbox_kernel", bbox_kernel<<<min(blocks_for(n, 256), 148u * 8u), 256, 0, st>>>(ws.pos64, n, box));
  PB_LAUNCH(ls, st, "to_soa_kernel",
            to_soa_kernel<<<blocks_for(n_pad, 256), 256, 0, st>>>(ws.pos64, n, n_pad, easing, box, ws.src4.as<float>(), meta));
  if (!n_targets) return cudaGetLastError();
  // 4 targets per thread when that still fills the chip, else 1; split the sources across
  // blockIdx.y until there are >= 2 CTAs per SM (partials summed in a fixed order afterwards)
  // T targets per thread.  Measured on B200 at 2^20 bodies, all targets: T=1 2.30e12, T=2 2.32e12,
  // T=4 2.36e12 interactions/s.  T is chosen for the best (speed x wave occupancy): 2 resident CTAs
  // per SM, so a wave is 296 CTAs.
  int T = 1;
  {
    const double speed[3] = {2.30, 2.32, 2.364};
    const int cand[3] = {1, 2, 4};
    double best = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double blocks = double(blocks_for(n_targets, DIRECT_THREADS * cand[i]));
      const double waves = std::ceil(blocks / 296.0);
      const double eff = blocks < 296.0 ? 1.0 : blocks / (waves * 296.0);  // small grids get source splits
      if (speed[i] * eff > best) { best = speed[i] * eff; T = cand[i]; }
    }
  }
  const unsigned tb = blocks_for(n_targets, DIRECT_THREADS * T);
  // The source splits decide the order in which a target's partial sums are added, so they follow from the
  // number of SOURCES alone (as if every body were a target, 4 per thread): a run whose targets are sharded
  // over several GPUs then adds in the same order - and gives the same bits - as one GPU evaluating them all.
  unsigned splits = 1;
  const unsigned max_splits = blocks_for(n, DIRECT_TILE);
  const unsigned tb_all = blocks_for(n, DIRECT_THREADS * 4);
  while (tb_all * splits < 148u * 2u && splits * 2 <= max_splits && splits < 64) splits *= 2;
  size_t per = (n + splits - 1) / splits;
  per = (per + DIRECT_TILE - 1) / DIRECT_TILE * DIRECT_TILE;
  splits = blocks_for(n, int(per));
  PB_PASS(ws.acc_part.ensure(size_t(splits) * n_targets * sizeof(float4)));
  const dim3 grid(tb, splits);
#define PB_DIRECT(TT)                                                                                         \
  PB_CUDA(cudaFuncSetAttribute(direct_kernel_x2<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                               3 * TT * DIRECT_THREADS * int(sizeof(double))));                               \
  PB_LAUNCH(ls, st, "direct_kernel_x2",                                                                       \
            direct_kernel_x2<TT><<<grid, DIRECT_THREADS, 3 * TT * DIRECT_THREADS * sizeof(double), st>>>(     \
                ws.src4.as<float>(), n_pad, per, t0, n_targets, easing, tiny, meta, ws.acc_part.as<float4>()))
  if (T == 4) { PB_DIRECT(4); }
  else if (T == 2) { PB_DIRECT(2); }
  else { PB_DIRECT(1); }
#undef PB_DIRECT
  PB_LAUNCH(ls, st, "direct_finish_kernel", direct_finish_kernel<<<blocks_for(n_targets, 256), 256, 0, st>>>(
      ws.acc_part.as<float4>(), int(splits), t0, n_targets, ws.pos64, ws.fixed, uint32_t(n),
      ws.acc.as<float4>()));
  return cudaGetLastError();
}

}  // namespace

// ---- sharded Barnes-Hut: host side ---------------------------------------------------------------
namespace {
inline uint32_t top_slots(int dim) { return dim == 3 ? TopTree<3>::SLOTS : TopTree<2>::SLOTS; }
inline uint32_t top_cells(int dim) { return dim == 3 ? TopTree<3>::CELLS : TopTree<2>::CELLS; }
}  // namespace

void ShardState::release() {
  DevBuf* all[] = {&cuts, &n_local, &slot_cell, &top_ce, &top_com, &top_info, &top_meta, &xacc, &flags};
  for (DevBuf* b : all) b->release();
  planned = false;
}

cudaError_t gravity_shard_setup(GravityWorkspace& ws, int kind, int rank, int world, size_t n) {
  ShardState& sh = ws.shard;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) {
    set_error("sharded run: world must be 1..8 (got rank %d of %d)", rank, world);
    return cudaErrorInvalidValue;
  }
  const int dim = kind == PB200_ASTRO ? 2 : 3;
  sh.rank = rank;
  sh.world = world;
  sh.planned = false;
  // 1/world of the bodies + 12.5 % (the cuts follow the level-K histogram, one step late) + one level-K cell's worth
  size_t cap = n / size_t(world) + n / size_t(8 * world) + n / 2048 + 4096;
  cap = (cap + 1023) / 1024 * 1024;
  if (cap > n + 1024) cap = (n + 1023) / 1024 * 1024;
  const bool fresh = !sh.flags.p;
  sh.n_cap = cap;
  PB_PASS(sh.cuts.ensure(16 * 8));
  PB_PASS(sh.n_local.ensure(16));
  PB_PASS(sh.slot_cell.ensure(size_t(top_slots(dim)) * 4));
  PB_PASS(sh.top_ce.ensure(size_t(top_cells(dim)) * sizeof(double4)));
  PB_PASS(sh.top_com.ensure(size_t(top_cells(dim)) * sizeof(double4)));
  PB_PASS(sh.top_info.ensure(size_t(top_cells(dim)) * sizeof(uint4)));
  PB_PASS(sh.top_meta.ensure(128 * 4));
  PB_PASS(sh.xacc.ensure(size_t(world) * sh.xacc_block_bytes()));
  PB_PASS(sh.flags.ensure(32 * 4));
  if (fresh) {  // epochs only ever grow: the flags are cleared once, when the buffer is made
    PB_CUDA(cudaMemset(sh.flags.p, 0, 32 * 4));
    PB_CUDA(cudaMemset(sh.top_meta.p, 0, 128 * 4));
    sh.epoch = 0;
  }
  return cudaSuccess;
}

bool gravity_shard_fits(const GravityWorkspace& ws) { return shard_sort_mode(ws.shard.n_cap) != 0; }

cudaError_t gravity_shard_plan(GravityWorkspace& ws, cudaStream_t st, LaunchStats& ls) {
  ShardState& sh = ws.shard;
  if (ws.tree_dim == 0 || !ws.sorted_key || ws.n == 0) {
    set_error("gravity_shard_plan: no full tree build to plan from");
    return cudaErrorInvalidValue;
  }
  // the splitter set the NEXT build reads (encode_and_sort toggles splitter_cur before it writes)
  uint64_t* spl = ws.splitters.as<uint64_t>() + SPLITTER_STRIDE * ws.splitter_cur;
  const unsigned nb = bucket_plan(sh.n_cap).nb;
  ws.spl_nb[ws.splitter_cur] = nb;
  if (ws.tree_dim == 2)
    PB_LAUNCH(ls, st, "shard_plan_kernel",
              shard_plan_kernel<2><<<1, 256, 0, st>>>(ws.sorted_key, ws.n, sh.rank, sh.world, sh.cuts.as<uint64_t>(), spl, nb));
  else
    PB_LAUNCH(ls, st, "shard_plan_kernel",
              shard_plan_kernel<3><<<1, 256, 0, st>>>(ws.sorted_key, ws.n, sh.rank, sh.world, sh.cuts.as<uint64_t>(), spl, nb));
  sh.planned = true;
  return cudaGetLastError();
}

namespace {
PeerTargets peer_targets(const ShardState& sh) {
  PeerTargets pt;
  for (int r = 0; r < 8; ++r) {
    const int q = r < sh.world ? r : sh.rank;
    pt.top_info[r] = static_cast<uint4*>(sh.peers.top_info[q]);
    pt.top_com[r] = static_cast<double4*>(sh.peers.top_com[q]);
    pt.meta[r] = static_cast<uint32_t*>(sh.peers.top_meta[q]);
    pt.xacc[r] = static_cast<char*>(sh.peers.xacc[q]);
    pt.flags[r] = static_cast<uint32_t*>(sh.peers.flags[q]);
  }
  pt.world = sh.world;
  pt.rank = sh.rank;
  return pt;
}

template <int DIM>
cudaError_t shard_build(GravityWorkspace& ws, cudaStream_t st, LaunchStats& ls) {
  ShardState& sh = ws.shard;
  using TT = TopTree<DIM>;
  sh.epoch += 1;
  PB_CUDA(cudaMemsetAsync(sh.slot_cell.p, 0, size_t(TT::SLOTS) * 4, st));
  ShardBuild sb;
  sb.cuts = sh.cuts.as<uint64_t>() + sh.rank;
  sb.n_local = sh.n_local.as<uint32_t>();
  sb.n_cap = sh.n_cap;
  sb.slots = TopSlots{sh.slot_cell.as<uint32_t>(), TT::K};
  // the permutation goes straight into this rank's block of the exchange buffer (behind the accelerations)
  sb.perm_out = reinterpret_cast<uint32_t*>(static_cast<char*>(sh.xacc.p) + size_t(sh.rank) * sh.xacc_block_bytes() +
                                            sh.n_cap * sizeof(float4));
  BuildOut bo;
  PB_PASS(tree_build<DIM>(ws, &sb, st, ls, &bo));
  PB_LAUNCH(ls, st, "top_export_kernel",
            pb_launch_pdl(top_export_kernel<DIM>, dim3(TT::SLOTS / 256), dim3(256), 0, st, sh.slot_cell.as<uint32_t>(),
                          bo.nref, ws.cell_start.as<uint32_t>(), bo.cells, sh.cuts.as<uint64_t>(), peer_targets(sh),
                          sh.epoch));
  return cudaGetLastError();
}

template <int DIM>
cudaError_t shard_walk(GravityWorkspace& ws, const GravityParams& prm, float easing, float tiny, cudaStream_t st,
                       LaunchStats& ls) {
  ShardState& sh = ws.shard;
  using TT = TopTree<DIM>;
  TopView top{sh.top_ce.as<double4>(), sh.top_com.as<double4>(), sh.top_info.as<uint4>(), sh.top_meta.as<uint32_t>(),
              sh.cuts.as<uint64_t>()};
  const size_t smem = size_t(TT::offset(TT::K)) * (sizeof(double4) + sizeof(uint4));
  PB_CUDA(cudaFuncSetAttribute(top_build_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  // (plain stream order, not a programmatic dependent: see shard_signal_then_wait)
  PB_LAUNCH(ls, st, "top_build_kernel",
            top_build_kernel<DIM><<<1, 1024, smem, st>>>(sh.world, ws.n, ws.extent_cur, top, peer_targets(sh), sh.epoch));
  PeerTables pt;
  for (int r = 0; r < 8; ++r) {
    pt.centre_ext[r] = static_cast<const double4*>(sh.peers.centre_ext[r < sh.world ? r : sh.rank]);
    pt.com[r] = static_cast<const double4*>(sh.peers.com[r < sh.world ? r : sh.rank]);
    pt.skip[r] = static_cast<const uint32_t*>(sh.peers.skip[r < sh.world ? r : sh.rank]);
  }
  pt.capacity = sh.peers.capacity;
  const NRef nref{sh.n_local.as<uint32_t>(), uint32_t(sh.n_cap)};
  PB_LAUNCH(ls, st, "walk_sharded_kernel",
            pb_launch_pdl(walk_sharded_kernel<DIM>, dim3(blocks_for(sh.n_cap, 256)), dim3(256), 0, st,
                          ws.spos64.as<double4>(), ws.perm, ws.fixed, nref, top, pt, peer_targets(sh), sh.n_cap, sh.epoch,
                          prm.theta, easing, tiny));
  return cudaGetLastError();
}
}  // namespace

cudaError_t gravity_shard_build(GravityWorkspace& ws, const GravityParams& prm, cudaStream_t st, LaunchStats& ls) {
  if (!ws.shard.planned) {
    set_error("gravity_shard_build: not planned");
    return cudaErrorInvalidValue;
  }
  return prm.kind == PB200_ASTRO ? shard_build<2>(ws, st, ls) : shard_build<3>(ws, st, ls);
}

cudaError_t gravity_shard_walk(GravityWorkspace& ws, const GravityParams& prm, cudaStream_t st, LaunchStats& ls) {
  const float easing = static_cast<float>(prm.easing);
  const float tiny = (prm.easing >= 3e-5) ? 1e-24f : 1e-11f;
  return prm.kind == PB200_ASTRO ? shard_walk<2>(ws, prm, easing, tiny, st, ls) : shard_walk<3>(ws, prm, easing, tiny, st, ls);
}

cudaError_t gravity_shard_scatter(GravityWorkspace& ws, cudaStream_t st, LaunchStats& ls) {
  ShardState& sh = ws.shard;
  PB_PASS(ws.acc.ensure(2 * ws.n * sizeof(float4)));  // 32-byte records (SHARD_ACC_STRIDE float4 each)
  const uint32_t* n_locals = sh.top_meta.as<uint32_t>() + (sh.epoch & 1u) * uint32_t(META_STRIDE) + uint32_t(META_BODIES);
  // (plain stream order, not a programmatic dependent: see shard_signal_then_wait)
  PB_LAUNCH(ls, st, "shard_scatter_kernel",
            shard_scatter_kernel<<<std::min(blocks_for(sh.n_cap, 256), 148u * 8u), 256, 0, st>>>(
                static_cast<const char*>(sh.xacc.p), sh.n_cap, n_locals, peer_targets(sh), sh.epoch, ws.acc.as<float4>(), ws.n));
  return cudaGetLastError();
}

void GravityWorkspace::release_all() {
  shard.release();
  DevBuf* all[] = {&src4, &key0, &key1, &idx0, &idx1, &bucket_key, &bucket_idx, &splitters, &nsv1, &nsv2, &spos64, &ab, &cell_start, &scan_tmp,
                   &tile_counts, &digit_base, &extent_bits, &tgt_list, &tgt_flags, &c_level, &c_head,
                   &c_count, &c_skip, &c_parent, &c_arrived, &c_kids, &c_ready, &c_centre_ext, &c_com, &acc, &acc_part, &sticky,
                   &counters};
  for (DevBuf* b : all) b->release();
}

cudaError_t gravity_evaluate(GravityWorkspace& ws, const GravityParams& prm, size_t t0, size_t t1,
                             cudaStream_t st, LaunchStats& ls, bool host_check) {
  const size_t n = ws.n;
  if (n == 0) return cudaSuccess;
  if (n >= 0xfffffff0ull) {
    set_error("n = %zu exceeds the 32-bit body index range", n);
    return cudaErrorInvalidValue;
  }
  if (t1 > n) t1 = n;
  if (t0 > t1) t0 = t1;
  PB_PASS(ws.acc.ensure(n * sizeof(float4)));
  if (t1 - t0 != n) PB_CUDA(cudaMemsetAsync(ws.acc.p, 0, n * sizeof(float4), st));
  const float easing = static_cast<float>(prm.easing);
  // r² gets `tiny` added so that r = 0 pairs give w·0 = 0 instead of inf·0; chosen so that
  // tiny·(tiny+e)² stays a normal fp32 number
  const float tiny = (prm.easing >= 3e-5) ? 1e-24f : 1e-11f;
  const bool direct = prm.kind == PB200_SIMPLE_ASTRO || !(prm.theta > 0.0);
  if (direct) {
    ws.n_cells = 0;
    ws.tree_dim = 0;
    return direct_evaluate(ws, t0, t1, easing, tiny, st, ls);
  }
  for (int attempt = 0; attempt < 3; ++attempt) {
    cudaError_t e = prm.kind == PB200_ASTRO ? tree_evaluate<2>(ws, prm, t0, t1, easing, tiny, st, ls)
                                            : tree_evaluate<3>(ws, prm, t0, t1, easing, tiny, st, ls);
    if (e != cudaSuccess) return e;
    if (!host_check) return cudaSuccess;  // caller runs gravity_check() later
    // the only host read of the build: cell total, tree depth, sort status (one small sync)
    TreeCheck chk;
    PB_PASS(gravity_check(ws, st, &chk));
    if (chk.ok()) return cudaSuccess;
    if (chk.sort_error) break;
    // overflow / truncated sort too short: gravity_check() already adjusted capacity / sort_lo
  }
  set_error("tree build did not converge (cell table overflow or sort failure)");
  return cudaErrorUnknown;
}

cudaError_t gravity_check(GravityWorkspace& ws, cudaStream_t st, TreeCheck* out) {
  *out = TreeCheck();
  if (ws.n == 0 || !ws.cell_start.p || !ws.sticky.p || ws.tree_dim == 0) return cudaSuccess;
  if (ws.unchecked_builds == 0) {  // nothing built since the last check: same verdict, no side effects
    out->total = ws.last_total;
    out->deepest_shared = ws.last_deepest;
    return cudaSuccess;
  }
  ws.unchecked_builds = 0;
  // {max cells, 1 + max deepest shared level, sort error, largest bucket, sharded build over capacity} since the last check
  uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  PB_CUDA(cudaMemcpyAsync(h, ws.sticky.p, 32, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaMemsetAsync(ws.sticky.p, 0, 32, st));
  PB_CUDA(cudaStreamSynchronize(st));
  out->shard_overflow = h[4] != 0;
  const int dim = ws.tree_dim;
  const int key_bits = dim * (dim == 3 ? 21 : 31);
  out->total = h[0];
  out->deepest_shared = int(h[1]) - 1;
  out->overflow = h[0] > ws.cell_cap;
  out->sort_error = h[2] != 0;
  // a sort of key bits [lo, key_bits) is exact iff no two neighbours with different keys agree on
  // all of those bits, i.e. share fewer than floor((key_bits - lo) / dim) levels
  out->sort_short = ws.last_lo > 0 && out->deepest_shared >= (key_bits - ws.last_lo) / dim;
  // the bucket-local sort leaves a bucket larger than its shared-memory tile unsorted (and the build
  // skipped): fall back to the mode the observed largest bucket allows and let the caller re-run
  out->max_bucket = h[3];
  ws.last_max_bucket = h[3];
  out->bucket_overflow = ws.last_mode != 0 && h[3] > LOCAL_CAP[ws.last_mode];
  static const char* mode_env = std::getenv("PB200_SORT_MODE");  // "lsd": global passes only (A/B runs)
  if (h[3] >= LOCAL_SKEWED) {
    ws.bucket_ban = 16;  // a thousand bodies on one spot: global passes for a while
  } else if (out->bucket_overflow) {
    // fast movers (bodies close to a heavy star cross many keys per step) make the quantile buckets of the
    // previous step fluctuate: go on with the next capacity class that takes the bucket just seen with 12 % to
    // spare (4608 -> 8192 -> 16384 keys of shared memory); only when none does, global passes for a while
    int m = ws.last_mode + 1;
    while (m <= 3 && LOCAL_CAP[m] < h[3] + h[3] / 8) ++m;
    if (m <= 3) ws.bucket_min_mode = m;
    else ws.bucket_ban = 16;
  } else if (ws.bucket_ban > 0) {
    --ws.bucket_ban;
  }
  // the splitters left by a verified build balance the buckets at ~n/256 bodies: pick the smallest
  // capacity with ~12 % headroom (bodies drift between evaluations)
  if ((mode_env && !std::strcmp(mode_env, "lsd")) || ws.bucket_ban > 0 || out->overflow || out->sort_short || out->sort_error)
    ws.sort_mode = 0;
  else {
    const int m = bucket_plan(ws.n).mode;
    ws.sort_mode = m ? std::max(m, ws.bucket_min_mode) : 0;
  }
  static const bool debug_check = std::getenv("PB200_DEBUG_CHECK") != nullptr;
  if (debug_check)
    std::fprintf(stderr, "[physim_b200] check: cells %u (cap %zu) deepest %d lo %d mode %d max_bucket %u overflow %d short %d "
                 "sort_error %d bucket_overflow %d ban %d -> next mode %d\n", h[0], ws.cell_cap, out->deepest_shared,
                 ws.last_lo, ws.last_mode, h[3], int(out->overflow), int(out->sort_short), int(out->sort_error),
                 int(out->bucket_overflow), ws.bucket_ban, ws.sort_mode);
  ws.n_cells = h[0];
  ws.last_total = h[0];
  ws.last_deepest = out->deepest_shared;
  if (out->sort_short) {
    ws.sort_lo = 0;
    if (ws.sort_extra_levels < 8) ++ws.sort_extra_levels;
  } else if (!out->sort_error) {
    // next time sort two levels deeper than anything seen now (plus one more for every time that
    // turned out too shallow); a deeper pair re-runs the build
    const int want_levels = out->deepest_shared + 3 + ws.sort_extra_levels;
    ws.sort_lo = std::max(0, key_bits - dim * want_levels);
  }
  return cudaSuccess;
}

// every cell's parent link, for inspection (the bottom-up sums only set the links they climb through)
cudaError_t gravity_fill_parents(GravityWorkspace& ws, cudaStream_t st, LaunchStats& ls) {
  if (ws.parents_filled || ws.n == 0 || ws.tree_dim == 0 || !ws.c_parent.p || !ws.cell_cap) return cudaSuccess;
  unsigned* flags = reinterpret_cast<unsigned*>(ws.extent_bits.as<unsigned long long>() + 1);
  CellArrays cells{ws.c_level.as<uint8_t>(),   ws.c_head.as<uint32_t>(),  ws.c_count.as<uint32_t>(),
                   ws.c_skip.as<uint32_t>(),   ws.c_parent.as<uint32_t>(), ws.c_arrived.as<uint32_t>(),
                   ws.c_centre_ext.as<double4>(), ws.c_com.as<double4>(), static_cast<uint32_t>(ws.cell_cap), SMALL_CELL,
                   flags + 2, 0u};
  PB_LAUNCH(ls, st, "parent_kernel",
            parent_kernel<<<blocks_for(ws.cell_cap, 256), 256, 0, st>>>(ws.cell_start.as<uint32_t>(), ws.n, cells));
  PB_LAUNCH(ls, st, "head_kernel", head_kernel<<<blocks_for(ws.n, 256), 256, 0, st>>>(ws.cell_start.as<uint32_t>(), ws.n, cells));
  ws.parents_filled = true;
  return cudaGetLastError();
}

cudaError_t gravity_cell_total(GravityWorkspace& ws, cudaStream_t st, uint32_t* total) {
  TreeCheck chk;
  PB_PASS(gravity_check(ws, st, &chk));
  *total = chk.total;
  return cudaSuccess;
}

cudaError_t gravity_count_interactions(GravityWorkspace& ws, cudaStream_t st, LaunchStats& ls,
                                       uint64_t* out) {
  *out = 0;
  if (ws.n == 0) return cudaSuccess;
  PB_PASS(ws.counters.ensure(8));
  PB_CUDA(cudaMemsetAsync(ws.counters.p, 0, 8, st));
  PB_LAUNCH(ls, st, "count_kernel", count_kernel<<<148 * 4, 256, 0, st>>>(ws.acc.as<float4>(), ws.n, ws.counters.as<unsigned long long>()));
  unsigned long long h = 0;
  PB_CUDA(cudaMemcpyAsync(&h, ws.counters.p, 8, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  *out = h;
  return cudaSuccess;
}

cudaError_t probe_fp32(double* tflops) {
  float* d = nullptr;
  const int blocks = 148 * 8, iters = 4096;
  PB_CUDA(cudaMalloc(&d, size_t(blocks) * 256 * 4));
  cudaEvent_t e0, e1;
  PB_CUDA(cudaEventCreate(&e0));
  PB_CUDA(cudaEventCreate(&e1));
  ffma_probe_kernel<<<blocks, 256>>>(d, 64);  // warm-up
  PB_CUDA(cudaEventRecord(e0));
  ffma_probe_kernel<<<blocks, 256>>>(d, iters);
  PB_CUDA(cudaEventRecord(e1));
  PB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = double(blocks) * 256.0 * iters * 16.0 * 8.0 * 2.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return cudaSuccess;
}

}  // namespace pb200

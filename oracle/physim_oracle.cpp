// physim_oracle.cpp — CPU restatement (fp64, one thread) of physim's gravity hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing the product ships may import, link or call this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
// only as the checker or as the timed CPU baseline.
//
// PARITY UNPINNED (values): the reference (jhb123/physim v0.4.4) holds no golden vector, known-answer
// test or fixture that asserts an acceleration, a centre of mass or a verlet output, and the Rust
// reference cannot be compiled in the build container (no rustc/cargo, nightly-only, no network).
// What IS pinned: every tree test the reference ships (leaf counts / inequalities,
// astro/src/octree.rs:218-519, astro/src/quadtree.rs:197-492) is restated in
// tests/test_oracle_tree.py against this file, and the verlet restatement is pinned against the
// analytic SHM solution of example_pipelines/shm.toml.
//
// Two oracles live here:
//   oracle 1  ("pointer tree")  — statement-for-statement restatement of the reference algorithm:
//              astro/src/lib.rs:27,39-114        (G, centre_of_mass, fake, force law)
//              astro/src/octree.rs:16-25,49-215  (node, push, walk, octant id / centre)
//              astro/src/quadtree.rs:16-25,49-194
//              astro/src/transformers.rs:31-69,122-160,219-244 (astro, astro2, simple_astro)
//              integrators/src/verlet.rs:23-107  (first step, regular step, dispatch)
//              physim-core/src/pipeline.rs:137-182 (step loop with its clones)
//   oracle 2  ("level array")   — the same tree rebuilt from sorted compare-and-halve keys as a
//              DFS pre-order cell table.  This is the data layout the CUDA path produces; it is
//              checked cell-for-cell against oracle 1 and the GPU is checked bit-for-bit against it.
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -ffp-contract=off; no fast-math, no FMA
// contraction: rustc never contracts a*b+c).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>

extern "C" {

// physim-core/src/lib.rs:16-37 ; c_plugin/physim.h:37-54
typedef struct Entity {
  double x, y, z, vx, vy, vz, radius, mass;
  uintptr_t id;
  bool fixed;
} Entity;
typedef struct Acceleration {
  double x, y, z;
} Acceleration;

}  // extern "C"

static_assert(sizeof(Entity) == 80, "Entity must be 80 bytes");
static_assert(sizeof(Acceleration) == 24, "Acceleration must be 24 bytes");

namespace {

constexpr double kG = 1.0;  // astro/src/lib.rs:27

struct OraclePanic : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- Star trait for Entity (astro/src/lib.rs:39-114) ---------------------------------------

inline void centre_of_mass(const Entity& a, const Entity& b, double out[3]) {
  const double total = a.mass + b.mass;
  const double inv = 1.0 / total;
  out[0] = (a.mass * a.x + b.mass * b.x) * inv;
  out[1] = (a.mass * a.y + b.mass * b.y) * inv;
  out[2] = (a.mass * a.z + b.mass * b.z) * inv;
}

inline Entity fake(const double c[3], double mass) {
  if (std::isnan(c[0])) throw OraclePanic("fake(): NaN centre (lib.rs:65-67)");
  Entity e;
  std::memset(&e, 0, sizeof e);
  e.x = c[0];
  e.y = c[1];
  e.z = c[2];
  e.mass = mass;  // radius 0, velocities 0, id 0, fixed false
  return e;
}

// F on a due to b:  r̂ · G·ma·mb / (r² + e), r̂ = (pb − pa)/|r|      (lib.rs:84-113)
inline void newton(const Entity& a, const Entity& b, double easing, double out[3]) {
  const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  const double r_norm = std::pow(dx * dx + dy * dy + dz * dz, 0.5);
  const double r_easing = dx * dx + dy * dy + dz * dz + easing;
  const double ux = (b.x - a.x) / r_norm;
  const double uy = (b.y - a.y) / r_norm;
  const double uz = (b.z - a.z) / r_norm;
  out[0] = ux * kG * a.mass * b.mass / r_easing;
  out[1] = uy * kG * a.mass * b.mass / r_easing;
  out[2] = uz * kG * a.mass * b.mass / r_easing;
}

// ---- bump arena (the reference uses bumpalo; only allocation behaviour, no arithmetic) -------

class Arena {
 public:
  void* alloc(size_t bytes) {
    bytes = (bytes + 15) & ~size_t(15);
    if (chunks_.empty() || used_ + bytes > cap_) {
      cap_ = std::max<size_t>(bytes, size_t(1) << 24);
      chunks_.emplace_back(static_cast<char*>(std::malloc(cap_)));
      if (!chunks_.back()) throw std::bad_alloc();
      used_ = 0;
    }
    void* p = chunks_.back().get() + used_;
    used_ += bytes;
    return p;
  }

 private:
  struct Free {
    void operator()(char* p) const { std::free(p); }
  };
  std::vector<std::unique_ptr<char, Free>> chunks_;
  size_t used_ = 0, cap_ = 0;
};

// ---- oracle 1: pointer tree (octree.rs / quadtree.rs) --------------------------------------
// DIM = 3: octree (8 children); DIM = 2: quadtree (4 children, z never split, z-centre inherited).

template <int DIM>
struct Node {
  static constexpr int NCH = 1 << DIM;
  double centre[3];
  double extent;
  bool has_entity;
  Entity entity;
  Node* child[NCH];
  // bookkeeping that the reference does not keep (does not feed back into the algorithm)
  uint32_t n_bodies;

  void init(const double c[3], double ext) {
    centre[0] = c[0];
    centre[1] = c[1];
    centre[2] = c[2];
    extent = ext;
    has_entity = false;
    n_bodies = 0;
    for (auto& ch : child) ch = nullptr;
  }

  bool childless() const {
    for (auto* ch : child)
      if (ch) return false;
    return true;
  }

  // octree.rs:160-165 / quadtree.rs:161-165 — strict '>'
  int octant(const Entity& p) const {
    int id = int(p.x > centre[0]) | (int(p.y > centre[1]) << 1);
    if (DIM == 3) id |= int(p.z > centre[2]) << 2;
    return id;
  }

  // octree.rs:167-215 / quadtree.rs:167-194
  void octant_centre(int id, double out[3]) const {
    const double half = extent / 2.0;
    out[0] = (id & 1) ? centre[0] + half : centre[0] - half;
    out[1] = (id & 2) ? centre[1] + half : centre[1] - half;
    if (DIM == 3)
      out[2] = (id & 4) ? centre[2] + half : centre[2] - half;
    else
      out[2] = centre[2];
  }

  Node* make_child(int id, Arena& arena) const {
    double c[3];
    octant_centre(id, c);
    Node* n = static_cast<Node*>(arena.alloc(sizeof(Node)));
    n->init(c, extent / 2.0);
    return n;
  }

  // octree.rs:59-128
  void push(const Entity& item, size_t depth, Arena& arena) {
    if (depth > 64) throw OraclePanic("Recursion too deep (octree.rs:60-62)");
    if (!has_entity) {
      entity = item;
      has_entity = true;
      n_bodies = 1;
      return;
    }
    const bool leaf = childless();
    if (leaf && std::fabs(entity.x - item.x) < 1e-9 && std::fabs(entity.y - item.y) < 1e-9 &&
        std::fabs(entity.z - item.z) < 1e-9) {
      // merge: keeps the NEW position, sums the masses (octree.rs:69-80)
      const double c[3] = {item.x, item.y, item.z};
      entity = fake(c, entity.mass + item.mass);
      n_bodies += 1;
      return;
    }
    double com[3];
    centre_of_mass(entity, item, com);
    const Entity resident = entity;
    entity = fake(com, resident.mass + item.mass);
    const uint32_t resident_bodies = n_bodies;
    n_bodies += 1;
    if (leaf) {
      const int id = octant(resident);
      Node* n = make_child(id, arena);
      n->entity = resident;
      n->has_entity = true;
      n->n_bodies = resident_bodies;
      child[id] = n;
    }
    const int id = octant(item);
    if (child[id]) {
      child[id]->push(item, depth + 1, arena);
    } else {
      Node* n = make_child(id, arena);
      n->entity = item;
      n->has_entity = true;
      n->n_bodies = 1;
      child[id] = n;
    }
  }

  // octree.rs:130-158 / quadtree.rs:130-159 — explicit stack, children pushed 0..NCH-1
  template <class Sink>
  void walk(const double loc[3], double theta, Sink&& sink) const {
    std::vector<const Node*> stack;
    stack.reserve(100);
    stack.push_back(this);
    while (!stack.empty()) {
      const Node* n = stack.back();
      stack.pop_back();
      if (!n->has_entity) continue;  // only the empty root
      const double ax = loc[0] - n->centre[0], ay = loc[1] - n->centre[1], az = loc[2] - n->centre[2];
      const double r = std::sqrt(ax * ax + ay * ay + az * az);
      if (n->extent / r < theta) {
        sink(n->entity);
        continue;
      }
      if (n->childless()) {
        sink(n->entity);
      } else {
        for (int k = 0; k < NCH; ++k)
          if (n->child[k]) stack.push_back(n->child[k]);
      }
    }
  }
};

template <int DIM>
struct Tree {
  Arena arena;
  Node<DIM> root;
  Tree(const double c[3], double extent) { root.init(c, extent); }
  void push(const Entity& e) { root.push(e, 0, arena); }
};

// transformers.rs:35-40 / :126-131 — max |coord| over x,y,z of every body; 1.0 for an empty state.
double state_extent(const Entity* s, size_t n) {
  if (n == 0) return 1.0;
  double m = std::fabs(s[0].x);
  // Rust's f64::max ignores NaN operands; std::fmax does the same.
  for (size_t i = 0; i < n; ++i) {
    m = std::fmax(m, std::fabs(s[i].x));
    m = std::fmax(m, std::fabs(s[i].y));
    m = std::fmax(m, std::fabs(s[i].z));
  }
  return m;
}

// transformers.rs:32-69 (DIM=2, "astro") and :123-160 (DIM=3, "astro2")
template <int DIM>
void bh_transform(double theta, double easing, const Entity* state, size_t n, Acceleration* acc,
                  uint32_t* n_inter, double* phase_seconds) {
  auto now = [] {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  };
  const double t0 = now();
  const double zero[3] = {0.0, 0.0, 0.0};
  Tree<DIM> tree(zero, 1.0 * state_extent(state, n));
  for (size_t i = 0; i < n; ++i) tree.push(state[i]);
  const double t1 = now();
  std::vector<Entity> found;
  found.reserve(100);
  for (size_t i = 0; i < n; ++i) {
    const Entity& a = state[i];
    if (n_inter) n_inter[i] = 0;
    if (a.fixed) continue;
    double f[3] = {0.0, 0.0, 0.0};
    const double loc[3] = {a.x, a.y, a.z};
    found.clear();
    tree.root.walk(loc, theta, [&](const Entity& e) { found.push_back(e); });
    if (n_inter) n_inter[i] = static_cast<uint32_t>(found.size());
    for (const Entity& b : found) {
      if (a.x == b.x && a.y == b.y && a.z == b.z) continue;
      double fij[3];
      newton(a, b, easing, fij);
      f[0] += fij[0];
      f[1] += fij[1];
      f[2] += fij[2];
    }
    acc[i].x += f[0] / a.mass;
    acc[i].y += f[1] / a.mass;
    acc[i].z += f[2] / a.mass;
  }
  const double t2 = now();
  if (phase_seconds) {
    phase_seconds[0] += t1 - t0;
    phase_seconds[1] += t2 - t1;
  }
}

// transformers.rs:220-244 ("simple_astro"); targets restricted to [t0, t1) for sampled timing
void direct_transform(double easing, const Entity* state, size_t n, Acceleration* acc, size_t t0,
                      size_t t1) {
  for (size_t i = t0; i < t1; ++i) {
    const Entity& a = state[i];
    if (a.fixed) continue;
    double f[3] = {0.0, 0.0, 0.0};
    for (size_t j = 0; j < n; ++j) {
      const Entity& b = state[j];
      if (a.x == b.x && a.y == b.y && a.z == b.z) continue;
      double fij[3];
      newton(a, b, easing, fij);
      f[0] += fij[0];
      f[1] += fij[1];
      f[2] += fij[2];
    }
    acc[i].x += f[0] / a.mass;
    acc[i].y += f[1] / a.mass;
    acc[i].z += f[2] / a.mass;
  }
}

// ---- verlet (integrators/src/verlet.rs:23-107) --------------------------------------------

typedef void (*AccFn)(void* ctx, const Entity* state, size_t n, Acceleration* acc);

struct Verlet {
  std::vector<Entity> previous;

  void step(const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx, double dt) {
    std::vector<Acceleration> acc(n, Acceleration{0.0, 0.0, 0.0});  // verlet.rs:93
    fn(ctx, ent, n, acc.data());                                     // verlet.rs:94
    const double dt2 = dt * dt;                                      // dt.powi(2)
    if (previous.size() != n) {                                      // verlet.rs:102-106
      previous.assign(ent, ent + n);
      for (size_t i = 0; i < n; ++i) {
        const Entity& e = ent[i];
        const Acceleration& a = acc[i];
        Entity o = e;
        o.x = e.x + e.vx * dt + 0.5 * a.x * dt2;
        o.y = e.y + e.vy * dt + 0.5 * a.y * dt2;
        o.z = e.z + e.vz * dt + 0.5 * a.z * dt2;
        o.vx = e.vx + a.x * dt;
        o.vy = e.vy + a.y * dt;
        o.vz = e.vz + a.z * dt;
        out[i] = o;
      }
    } else {
      for (size_t i = 0; i < n; ++i) {
        const Entity& e = ent[i];
        const Entity& p = previous[i];
        const Acceleration& a = acc[i];
        Entity o = e;
        o.x = 2.0 * e.x - p.x + a.x * dt2;
        o.y = 2.0 * e.y - p.y + a.y * dt2;
        o.z = 2.0 * e.z - p.z + a.z * dt2;
        o.vx = (o.x - e.x) / dt;
        o.vy = (o.y - e.y) / dt;
        o.vz = (o.z - e.z) / dt;
        out[i] = o;
      }
      previous.assign(ent, ent + n);  // verlet.rs:81
    }
  }
};

// ---- euler (integrators/src/euler.rs:17-50) and rk4 (integrators/src/rk4.rs:23-183) --------------
// the integrators either side of verlet.  euler is verlet's first-step formula
// every step; rk4 evaluates the accelerations four times on temporaries and honours `fixed`.

struct Euler {
  void step(const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx, double dt) {
    std::vector<Acceleration> acc(n, Acceleration{0.0, 0.0, 0.0});  // euler.rs:25
    fn(ctx, ent, n, acc.data());
    const double dt2 = dt * dt;
    for (size_t i = 0; i < n; ++i) {
      const Entity& e = ent[i];
      const Acceleration& a = acc[i];
      Entity o = e;
      o.x = e.x + e.vx * dt + 0.5 * a.x * dt2;  // euler.rs:30-32
      o.y = e.y + e.vy * dt + 0.5 * a.y * dt2;
      o.z = e.z + e.vz * dt + 0.5 * a.z * dt2;
      o.vx = e.vx + a.x * dt;                   // euler.rs:35-37
      o.vy = e.vy + a.y * dt;
      o.vz = e.vz + a.z * dt;
      out[i] = o;
    }
  }
};

struct Rk4 {
  // k = (dt * e.v, dt * a) with the other fields of e               (rk4.rs:36-50 and repeats)
  static void slope(const std::vector<Entity>& at, const std::vector<Acceleration>& f, double dt,
                    std::vector<Entity>& k) {
    k = at;
    for (size_t i = 0; i < at.size(); ++i) {
      k[i].x = dt * at[i].vx; k[i].y = dt * at[i].vy; k[i].z = dt * at[i].vz;
      k[i].vx = dt * f[i].x;  k[i].vy = dt * f[i].y;  k[i].vz = dt * f[i].z;
    }
  }
  // temp = e + w * k (w = 0.5, 0.5, 1.0; the last one is written `e.x + k.x`)   (rk4.rs:53-66 ...)
  static void advance(const Entity* e, const std::vector<Entity>& k, double w, bool plain,
                      std::vector<Entity>& t) {
    t.assign(e, e + k.size());
    for (size_t i = 0; i < k.size(); ++i) {
      if (plain) {
        t[i].x = e[i].x + k[i].x;   t[i].y = e[i].y + k[i].y;   t[i].z = e[i].z + k[i].z;
        t[i].vx = e[i].vx + k[i].vx; t[i].vy = e[i].vy + k[i].vy; t[i].vz = e[i].vz + k[i].vz;
      } else {
        t[i].x = e[i].x + w * k[i].x;   t[i].y = e[i].y + w * k[i].y;   t[i].z = e[i].z + w * k[i].z;
        t[i].vx = e[i].vx + w * k[i].vx; t[i].vy = e[i].vy + w * k[i].vy; t[i].vz = e[i].vz + w * k[i].vz;
      }
    }
  }
  void step(const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx, double dt) {
    std::vector<Entity> cur(ent, ent + n), k1, k2, k3, k4, tmp;
    std::vector<Acceleration> f(n);
    auto eval = [&](const std::vector<Entity>& at, std::vector<Entity>& k) {
      std::fill(f.begin(), f.end(), Acceleration{0.0, 0.0, 0.0});
      fn(ctx, at.data(), n, f.data());
      slope(at, f, dt, k);
    };
    eval(cur, k1);
    advance(ent, k1, 0.5, false, tmp);
    eval(tmp, k2);
    advance(ent, k2, 0.5, false, tmp);
    eval(tmp, k3);
    advance(ent, k3, 1.0, true, tmp);
    eval(tmp, k4);
    for (size_t i = 0; i < n; ++i) {  // rk4.rs:150-176
      const Entity& e = ent[i];
      Entity& ns = out[i];
      if (!e.fixed) {
        ns.x = e.x + (k1[i].x + 2.0 * k2[i].x + 2.0 * k3[i].x + k4[i].x) / 6.0;
        ns.y = e.y + (k1[i].y + 2.0 * k2[i].y + 2.0 * k3[i].y + k4[i].y) / 6.0;
        ns.z = e.z + (k1[i].z + 2.0 * k2[i].z + 2.0 * k3[i].z + k4[i].z) / 6.0;
        ns.vx = e.vx + (k1[i].vx + 2.0 * k2[i].vx + 2.0 * k3[i].vx + k4[i].vx) / 6.0;
        ns.vy = e.vy + (k1[i].vy + 2.0 * k2[i].vy + 2.0 * k3[i].vy + k4[i].vy) / 6.0;
        ns.vz = e.vz + (k1[i].vz + 2.0 * k2[i].vz + 2.0 * k3[i].vz + k4[i].vz) / 6.0;
      } else {
        ns.x = e.x; ns.y = e.y; ns.z = e.z;
        ns.vx = 0.0; ns.vy = 0.0; ns.vz = 0.0;
      }
      ns.mass = e.mass; ns.radius = e.radius; ns.id = e.id; ns.fixed = e.fixed;
    }
  }
};

// kind: 0 verlet, 1 euler, 2 rk4
struct Integrator {
  int kind;
  Verlet verlet;
  Euler euler;
  Rk4 rk4;
  void step(const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx, double dt) {
    if (kind == 0) verlet.step(ent, out, n, fn, ctx, dt);
    else if (kind == 1) euler.step(ent, out, n, fn, ctx, dt);
    else rk4.step(ent, out, n, fn, ctx, dt);
  }
};

// ---- oracle 2: level-array tree ------------------------------------------------------------
// Keys: one digit per level, digit = octant id of the reference (x | y<<1 | z<<2), obtained by
// repeating the reference's compare (strict >) and halve (centre ± extent/2) in fp64 so the
// visited centres are the very doubles the pointer tree holds.  LMAX levels: 21 (octree, 63 bits)
// or 31 (quadtree, 62 bits).  Bodies that share all LMAX digits and chain within 1e-9 of each
// other in index order form one merged leaf (octree.rs:69-80); bodies that share all digits but
// are further apart become sibling leaves at pseudo-level LMAX+1 (stated deviation: the reference
// would keep splitting, to at most depth 64).

template <int DIM>
constexpr int lmax() {
  return DIM == 3 ? 21 : 31;
}

template <int DIM>
uint64_t encode_key(double px, double py, double pz, double extent) {
  double cx = 0.0, cy = 0.0, cz = 0.0, ext = extent;
  uint64_t key = 0;
  for (int l = 0; l < lmax<DIM>(); ++l) {
    const double half = ext / 2.0;
    const unsigned bx = px > cx, by = py > cy;
    unsigned digit = bx | (by << 1);
    cx = bx ? cx + half : cx - half;
    cy = by ? cy + half : cy - half;
    if (DIM == 3) {
      const unsigned bz = pz > cz;
      digit |= bz << 2;
      cz = bz ? cz + half : cz - half;
    }
    key = (key << DIM) | digit;
    ext = half;
  }
  return key;
}

struct CellTable {
  // per sorted body
  std::vector<uint64_t> key;
  std::vector<uint32_t> perm;        // sorted position -> original index
  std::vector<uint32_t> cell_start;  // first cell headed by this sorted body (size n+1)
  // per cell, DFS pre-order, children in ascending digit order
  std::vector<uint8_t> level;
  std::vector<uint32_t> head, count, skip, parent;
  std::vector<double> cx, cy, cz, ext;     // geometric centre and half-width
  std::vector<double> mx, my, mz, mass;    // centre of mass (leaf: body position) and mass
  double extent = 1.0;
};

template <int DIM>
int shared_levels(uint64_t a, uint64_t b) {
  const uint64_t x = a ^ b;
  if (x == 0) return lmax<DIM>();
  const int top = 63 - __builtin_clzll(x);       // highest differing bit
  const int digit_from_bottom = top / DIM;       // which digit (0 = deepest level) differs
  return lmax<DIM>() - 1 - digit_from_bottom;    // levels fully shared
}

template <int DIM>
void build_cell_table(const Entity* state, size_t n, CellTable& t) {
  constexpr int LM = lmax<DIM>();
  t.extent = 1.0 * state_extent(state, n);
  std::vector<uint64_t> k(n);
  for (size_t i = 0; i < n; ++i) k[i] = encode_key<DIM>(state[i].x, state[i].y, state[i].z, t.extent);
  t.perm.resize(n);
  for (size_t i = 0; i < n; ++i) t.perm[i] = static_cast<uint32_t>(i);
  std::stable_sort(t.perm.begin(), t.perm.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
  t.key.resize(n);
  for (size_t s = 0; s < n; ++s) t.key[s] = k[t.perm[s]];

  auto body = [&](size_t s) -> const Entity& { return state[t.perm[s]]; };
  // Merged leaves (octree.rs:69-80).  A body that arrives at a LEAF whose resident lies within 1e-9
  // on every axis is merged into it: masses add, the newer position wins.  In the final tree that is
  // a run of sorted-adjacent bodies, each within 1e-9 of the previous one, that is alone in its cell
  // at the level where it separates from its outer neighbours.  joined[s] = "s is in the same
  // merged leaf as s-1".  (Runs that straddle a cell boundary above that level are insertion-order
  // dependent in the reference; they are left unmerged here.)
  auto close = [&](size_t p, size_t q) {
    const Entity &u = body(p), &v = body(q);
    return std::fabs(u.x - v.x) < 1e-9 && std::fabs(u.y - v.y) < 1e-9 && std::fabs(u.z - v.z) < 1e-9;
  };
  std::vector<uint8_t> joined_flag(n + 1, 0);
  for (size_t r0 = 0; r0 < n;) {
    size_t r1 = r0;
    int inner = LM + 1;
    while (r1 + 1 < n && close(r1, r1 + 1)) {
      inner = std::min(inner, shared_levels<DIM>(t.key[r1], t.key[r1 + 1]));
      ++r1;
    }
    if (r1 > r0) {
      const int a = r0 > 0 ? shared_levels<DIM>(t.key[r0 - 1], t.key[r0]) : -1;
      const int b = r1 + 1 < n ? shared_levels<DIM>(t.key[r1], t.key[r1 + 1]) : -1;
      if (inner >= std::min(std::max(a, b) + 1, LM))
        for (size_t j = r0 + 1; j <= r1; ++j) joined_flag[j] = 1;
    }
    r0 = r1 + 1;
  }
  auto joined = [&](size_t s) { return s < n && joined_flag[s] != 0; };
  // levels shared between the unit ending at s-1 and the unit starting at s (s is a unit head)
  auto shared_before = [&](size_t s) -> int {
    if (s == 0) return -1;
    return shared_levels<DIM>(t.key[s - 1], t.key[s]);
  };

  t.cell_start.assign(n + 1, 0);
  std::vector<int> a_of(n, 0), b_of(n, 0);
  std::vector<uint32_t> unit_end(n, 0);
  uint32_t total = 0;
  for (size_t s = 0; s < n; ++s) {
    t.cell_start[s] = total;
    if (joined(s)) continue;  // not a unit head: heads no cell
    size_t e = s + 1;
    while (joined(e)) ++e;
    unit_end[s] = static_cast<uint32_t>(e);
    const int a = shared_before(s);
    const int b = (e < n) ? shared_levels<DIM>(t.key[e - 1], t.key[e]) : -1;
    a_of[s] = a;
    b_of[s] = b;
    total += static_cast<uint32_t>(std::max(0, b - a) + 1);
  }
  t.cell_start[n] = total;
  if (n == 0) total = 0;

  auto rs = [&](auto& v) { v.assign(total, 0); };
  rs(t.level); rs(t.head); rs(t.count); rs(t.skip); rs(t.parent);
  rs(t.cx); rs(t.cy); rs(t.cz); rs(t.ext); rs(t.mx); rs(t.my); rs(t.mz); rs(t.mass);

  for (size_t s = 0; s < n; ++s) {
    if (joined(s)) continue;
    const int a = a_of[s], b = b_of[s];
    const int top = a + 1, leaf_level = std::max(a, b) + 1;
    for (int lev = top; lev <= leaf_level; ++lev) {
      const uint32_t c = t.cell_start[s] + static_cast<uint32_t>(lev - top);
      t.level[c] = static_cast<uint8_t>(lev);
      t.head[c] = static_cast<uint32_t>(s);
      // run end: first sorted body past s that does not share `lev` digits with s
      size_t e;
      if (lev == leaf_level) {
        e = unit_end[s];
      } else {
        e = unit_end[s];
        while (e < n && shared_levels<DIM>(t.key[s], t.key[e]) >= lev) ++e;
      }
      t.count[c] = static_cast<uint32_t>(e - s);
      t.skip[c] = t.cell_start[e];
      // geometric centre: replay the head body's digits
      double cx = 0.0, cy = 0.0, cz = 0.0, ext = t.extent;
      for (int l = 0; l < lev; ++l) {
        const double half = ext / 2.0;
        unsigned digit;
        if (l < LM) {
          digit = static_cast<unsigned>((t.key[s] >> (DIM * (LM - 1 - l))) & ((1u << DIM) - 1));
        } else {  // pseudo level below the key: compare the head body directly
          const Entity& p = body(s);
          digit = unsigned(p.x > cx) | (unsigned(p.y > cy) << 1);
          if (DIM == 3) digit |= unsigned(p.z > cz) << 2;
        }
        cx = (digit & 1) ? cx + half : cx - half;
        cy = (digit & 2) ? cy + half : cy - half;
        if (DIM == 3) cz = (digit & 4) ? cz + half : cz - half;
        ext = half;
      }
      t.cx[c] = cx; t.cy[c] = cy; t.cz[c] = cz; t.ext[c] = ext;
      // parent
      if (lev == 0) {
        t.parent[c] = 0xFFFFFFFFu;
      } else if (lev > top) {
        t.parent[c] = c - 1;
      } else {
        size_t h = s;  // head of the level-(lev-1) run containing s
        while (h > 0 && (joined(h) || shared_levels<DIM>(t.key[h - 1], t.key[s]) >= lev - 1)) --h;
        const int ah = (h == 0) ? -1 : shared_levels<DIM>(t.key[h - 1], t.key[h]);
        t.parent[c] = t.cell_start[h] + static_cast<uint32_t>(lev - 1 - (ah + 1));
      }
    }
  }
  // masses and centres of mass: plain sums over the run (fp64), leaf = merged-unit rule
  for (uint32_t c = total; c-- > 0;) {
    const size_t s = t.head[c], e = s + t.count[c];
    const bool is_leaf = (c + 1 == t.skip[c]);
    if (is_leaf) {
      double m = 0.0;
      for (size_t j = s; j < e; ++j) m += body(j).mass;
      size_t newest = s;  // the member inserted last (largest original index) keeps its position
      for (size_t j = s; j < e; ++j)
        if (t.perm[j] > t.perm[newest]) newest = j;
      const Entity& last = body(newest);
      t.mass[c] = m; t.mx[c] = last.x; t.my[c] = last.y; t.mz[c] = last.z;
    } else {
      double m = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
      for (uint32_t ch = c + 1; ch < t.skip[c]; ch = t.skip[ch]) {
        m += t.mass[ch];
        sx += t.mass[ch] * t.mx[ch];
        sy += t.mass[ch] * t.my[ch];
        sz += t.mass[ch] * t.mz[ch];
      }
      t.mass[c] = m; t.mx[c] = sx / m; t.my[c] = sy / m; t.mz[c] = sz / m;
    }
  }
}

// walk of the cell table with the reference's acceptance rule; ascending-cell order
template <int DIM>
void cell_table_transform(const CellTable& t, double theta, double easing, const Entity* state,
                          size_t n, Acceleration* acc, uint32_t* n_inter) {
  const uint32_t total = static_cast<uint32_t>(t.level.size());
  for (size_t i = 0; i < n; ++i) {
    const Entity& a = state[i];
    if (n_inter) n_inter[i] = 0;
    if (a.fixed) continue;
    double f[3] = {0, 0, 0};
    uint32_t cnt = 0;
    uint32_t c = 0;
    while (c < total) {
      const double ax = a.x - t.cx[c], ay = a.y - t.cy[c], az = a.z - t.cz[c];
      const double r = std::sqrt(ax * ax + ay * ay + az * az);
      const bool is_leaf = (c + 1 == t.skip[c]);
      if (is_leaf || t.ext[c] / r < theta) {
        ++cnt;
        if (!(a.x == t.mx[c] && a.y == t.my[c] && a.z == t.mz[c])) {
          const double c3[3] = {t.mx[c], t.my[c], t.mz[c]};
          const Entity b = fake(c3, t.mass[c]);
          double fij[3];
          newton(a, b, easing, fij);
          f[0] += fij[0]; f[1] += fij[1]; f[2] += fij[2];
        }
        c = t.skip[c];
      } else {
        c = c + 1;
      }
    }
    if (n_inter) n_inter[i] = cnt;
    acc[i].x += f[0] / a.mass;
    acc[i].y += f[1] / a.mass;
    acc[i].z += f[2] / a.mass;
  }
}

// DFS pre-order dump of the pointer tree (children ascending) for comparison with oracle 2
template <int DIM>
void dump_preorder(const Node<DIM>* n, int level, std::vector<double>& out) {
  if (!n->has_entity) return;
  const bool leaf = n->childless();
  const double row[12] = {double(level), n->centre[0], n->centre[1], n->centre[2], n->extent,
                          double(n->n_bodies), n->entity.mass, n->entity.x, n->entity.y,
                          n->entity.z, leaf ? 1.0 : 0.0, 0.0};
  out.insert(out.end(), row, row + 12);
  for (int k = 0; k < Node<DIM>::NCH; ++k)
    if (n->child[k]) dump_preorder<DIM>(n->child[k], level + 1, out);
}

struct AnyTree {
  int dim;
  std::unique_ptr<Tree<2>> q;
  std::unique_ptr<Tree<3>> o;
};

struct TransformCtx {
  int kind;  // 0 astro (quadtree), 1 astro2 (octree), 2 simple_astro
  double theta, easing;
  double phases[2];
};

void transform_cb(void* ctx, const Entity* s, size_t n, Acceleration* acc) {
  auto* c = static_cast<TransformCtx*>(ctx);
  if (c->kind == 0)
    bh_transform<2>(c->theta, c->easing, s, n, acc, nullptr, c->phases);
  else if (c->kind == 1)
    bh_transform<3>(c->theta, c->easing, s, n, acc, nullptr, c->phases);
  else
    direct_transform(c->easing, s, n, acc, 0, n);
}

thread_local char g_err[256];

template <class F>
int guarded(F&& f) {
  try {
    f();
    g_err[0] = 0;
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(g_err, sizeof g_err, "%s", e.what());
    return -1;
  }
}

}  // namespace

// ---- C API (ctypes) -------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error(void) { return g_err; }

void* oracle_tree_new(int dim, const double* centre, double extent) {
  auto* t = new AnyTree;
  t->dim = dim;
  if (dim == 2)
    t->q = std::make_unique<Tree<2>>(centre, extent);
  else
    t->o = std::make_unique<Tree<3>>(centre, extent);
  return t;
}
void oracle_tree_free(void* h) { delete static_cast<AnyTree*>(h); }

int oracle_tree_push(void* h, const Entity* e, size_t n) {
  auto* t = static_cast<AnyTree*>(h);
  return guarded([&] {
    for (size_t i = 0; i < n; ++i) {
      if (t->dim == 2)
        t->q->push(e[i]);
      else
        t->o->push(e[i]);
    }
  });
}

// get_leaves_with_resolution: returns the number of entities; writes up to cap of them
size_t oracle_tree_query(void* h, const double* loc, double theta, Entity* out, size_t cap) {
  auto* t = static_cast<AnyTree*>(h);
  size_t cnt = 0;
  auto sink = [&](const Entity& e) {
    if (out && cnt < cap) out[cnt] = e;
    ++cnt;
  };
  if (t->dim == 2)
    t->q->root.walk(loc, theta, sink);
  else
    t->o->root.walk(loc, theta, sink);
  return cnt;
}

// rows of 12 doubles: level, cx, cy, cz, extent, n_bodies, mass, ex, ey, ez, is_leaf, 0
size_t oracle_tree_dump(void* h, double* out, size_t cap_rows) {
  auto* t = static_cast<AnyTree*>(h);
  std::vector<double> rows;
  if (t->dim == 2)
    dump_preorder<2>(&t->q->root, 0, rows);
  else
    dump_preorder<3>(&t->o->root, 0, rows);
  const size_t n = rows.size() / 12;
  if (out) std::memcpy(out, rows.data(), std::min(n, cap_rows) * 12 * sizeof(double));
  return n;
}

double oracle_state_extent(const Entity* s, size_t n) { return state_extent(s, n); }

// kind: 0 astro (quadtree), 1 astro2 (octree), 2 simple_astro.  acc is accumulated into.
// n_inter (optional, size n): entities returned by the walk per target (0 for fixed targets).
// phases (optional, 2 doubles): += seconds spent in build and in walk+force.
int oracle_transform(int kind, double theta, double easing, const Entity* s, size_t n,
                     Acceleration* acc, uint32_t* n_inter, double* phases) {
  return guarded([&] {
    if (kind == 0)
      bh_transform<2>(theta, easing, s, n, acc, n_inter, phases);
    else if (kind == 1)
      bh_transform<3>(theta, easing, s, n, acc, n_inter, phases);
    else
      direct_transform(easing, s, n, acc, 0, n);
  });
}

// simple_astro restricted to targets [t0, t1): bounded sample for timing O(N²) configs
void oracle_direct_range(double easing, const Entity* s, size_t n, Acceleration* acc, size_t t0,
                         size_t t1) {
  direct_transform(easing, s, n, acc, t0, t1);
}

void* oracle_verlet_new(void) { return new Verlet; }
void oracle_verlet_free(void* v) { delete static_cast<Verlet*>(v); }
void oracle_verlet_step(void* v, const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx,
                        double dt) {
  static_cast<Verlet*>(v)->step(ent, out, n, fn, ctx, dt);
}

// The reference's simulation loop (pipeline.rs:143-182) for one gravity transform + verlet:
// integrate(state -> new_state); state = new_state.clone(); send(new_state.clone()).
// `state` is updated in place to the final state.  seconds[0..3] += build, walk+force,
// integrate+copies.  Returns 0, or -1 on a reference panic.
int oracle_run_pipeline(int kind, double theta, double easing, Entity* state, size_t n, double dt,
                        size_t iterations, double* seconds) {
  return guarded([&] {
    auto now = [] {
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    TransformCtx ctx{kind, theta, easing, {0.0, 0.0}};
    Verlet verlet;
    std::vector<Entity> cur(state, state + n), nxt(state, state + n);
    const double t0 = now();
    for (size_t it = 0; it < iterations; ++it) {
      verlet.step(cur.data(), nxt.data(), n, transform_cb, &ctx, dt);
      cur = nxt;                       // state = new_state.clone()
      std::vector<Entity> sent = nxt;  // simulation_sender.send(new_state.clone())
      asm volatile("" : : "r"(sent.data()) : "memory");
    }
    const double total = now() - t0;
    std::memcpy(state, cur.data(), n * sizeof(Entity));
    if (seconds) {
      seconds[0] += ctx.phases[0];
      seconds[1] += ctx.phases[1];
      seconds[2] += total - ctx.phases[0] - ctx.phases[1];
    }
  });
}

void* oracle_integrator_new(int kind) {
  auto* g = new Integrator;
  g->kind = kind;
  return g;
}
void oracle_integrator_free(void* g) { delete static_cast<Integrator*>(g); }
void oracle_integrator_step(void* g, const Entity* ent, Entity* out, size_t n, AccFn fn, void* ctx,
                            double dt) {
  static_cast<Integrator*>(g)->step(ent, out, n, fn, ctx, dt);
}

// oracle_run_pipeline with a choice of integrator (0 verlet, 1 euler, 2 rk4)
int oracle_run_pipeline_with(int integrator, int kind, double theta, double easing, Entity* state,
                             size_t n, double dt, size_t iterations) {
  return guarded([&] {
    TransformCtx ctx{kind, theta, easing, {0.0, 0.0}};
    Integrator integ;
    integ.kind = integrator;
    std::vector<Entity> cur(state, state + n), nxt(state, state + n);
    for (size_t it = 0; it < iterations; ++it) {
      integ.step(cur.data(), nxt.data(), n, transform_cb, &ctx, dt);
      cur = nxt;
    }
    std::memcpy(state, cur.data(), n * sizeof(Entity));
  });
}

// ---- oracle 2 --------------------------------------------------------------------------------

uint64_t oracle_encode_key(int dim, double x, double y, double z, double extent) {
  return dim == 2 ? encode_key<2>(x, y, z, extent) : encode_key<3>(x, y, z, extent);
}

void* oracle_table_build(int dim, const Entity* s, size_t n) {
  auto* t = new CellTable;
  const int rc = guarded([&] {
    if (dim == 2)
      build_cell_table<2>(s, n, *t);
    else
      build_cell_table<3>(s, n, *t);
  });
  if (rc != 0) {
    delete t;
    return nullptr;
  }
  return t;
}
void oracle_table_free(void* h) { delete static_cast<CellTable*>(h); }
size_t oracle_table_cells(void* h) { return static_cast<CellTable*>(h)->level.size(); }
double oracle_table_extent(void* h) { return static_cast<CellTable*>(h)->extent; }

// copies out whichever arrays are non-NULL
void oracle_table_get(void* h, uint64_t* key, uint32_t* perm, uint32_t* cell_start, uint8_t* level,
                      uint32_t* head, uint32_t* count, uint32_t* skip, uint32_t* parent,
                      double* centre_ext /*4 per cell*/, double* com_mass /*4 per cell*/) {
  auto* t = static_cast<CellTable*>(h);
  const size_t n = t->key.size(), c = t->level.size();
  if (key) std::memcpy(key, t->key.data(), n * 8);
  if (perm) std::memcpy(perm, t->perm.data(), n * 4);
  if (cell_start) std::memcpy(cell_start, t->cell_start.data(), (n + 1) * 4);
  if (level) std::memcpy(level, t->level.data(), c);
  if (head) std::memcpy(head, t->head.data(), c * 4);
  if (count) std::memcpy(count, t->count.data(), c * 4);
  if (skip) std::memcpy(skip, t->skip.data(), c * 4);
  if (parent) std::memcpy(parent, t->parent.data(), c * 4);
  if (centre_ext)
    for (size_t i = 0; i < c; ++i) {
      centre_ext[4 * i + 0] = t->cx[i];
      centre_ext[4 * i + 1] = t->cy[i];
      centre_ext[4 * i + 2] = t->cz[i];
      centre_ext[4 * i + 3] = t->ext[i];
    }
  if (com_mass)
    for (size_t i = 0; i < c; ++i) {
      com_mass[4 * i + 0] = t->mx[i];
      com_mass[4 * i + 1] = t->my[i];
      com_mass[4 * i + 2] = t->mz[i];
      com_mass[4 * i + 3] = t->mass[i];
    }
}

int oracle_table_transform(int dim, void* h, double theta, double easing, const Entity* s, size_t n,
                           Acceleration* acc, uint32_t* n_inter) {
  auto* t = static_cast<CellTable*>(h);
  return guarded([&] {
    if (dim == 2)
      cell_table_transform<2>(*t, theta, easing, s, n, acc, n_inter);
    else
      cell_table_transform<3>(*t, theta, easing, s, n, acc, n_inter);
  });
}

}  // extern "C"

"""literal.py — a SECOND, independent restatement of the reference's Barnes-Hut transform in pure Python.

TEST INFRASTRUCTURE ONLY (same rule as physim_oracle.cpp: only tests/ may import it).

Purpose: the C++ oracle (physim_oracle.cpp, "oracle 1") is the thing every GPU parity test trusts, and
the reference ships no value-asserting test that could pin it.  This file restates the
same Rust, statement by statement, a second time and from the Rust text alone — no code, helper or
data structure is shared with physim_oracle.cpp — so that the two restatements can be required to agree
BIT FOR BIT (tests/test_oracle_literal.py).  A slip in either one (operation order, strictness of a
comparison, the merge rule, the stack order of the walk) shows up as a mismatch.  It does not replace a
golden vector from a real `physim` build (tests/golden/README.md describes how to add one): parity
stays "unpinned" for values until such a file exists.

Python floats are IEEE-754 binary64 and `+ - * /` are correctly rounded, as in Rust; `x.powi(2)` is
`x * x`; `powf(0.5)` is libm `pow(x, 0.5)` (math.pow); `sqrt` is correctly rounded (math.sqrt).
Sized for N <= a few thousand bodies (pure-Python loops).

Follows:
  astro/src/octree.rs:49-215      OctreeNode::new / push / get_leaves_with_resolution / get_octant_id(_centre)
  astro/src/quadtree.rs:49-194    the same with 4 children, z never split
  astro/src/lib.rs:40-113         centre_of_mass, fake, newtons_law_of_universal_gravitation
  astro/src/transformers.rs:32-69,123-160   AstroElement / AstroOctreeElement::transform
"""
import math

G = 1.0  # astro/src/lib.rs:27


class Ent:
    """physim-core/src/lib.rs:16-30 — only the fields the path reads."""
    __slots__ = ("x", "y", "z", "mass", "fixed")

    def __init__(self, x, y, z, mass, fixed=False):
        self.x, self.y, self.z, self.mass, self.fixed = x, y, z, mass, fixed

    def get_centre(self):  # lib.rs:57-63
        return [self.x, self.y, self.z]

    def get_mass(self):  # lib.rs:53-55
        return self.mass

    def centre_of_mass(self, other):  # lib.rs:40-51
        total_mass = self.mass + other.mass
        inv_total_mass = 1.0 / total_mass
        return [
            (self.mass * self.x + other.mass * other.x) * inv_total_mass,
            (self.mass * self.y + other.mass * other.y) * inv_total_mass,
            (self.mass * self.z + other.mass * other.z) * inv_total_mass,
        ]

    @staticmethod
    def fake(centre, mass):  # lib.rs:65-77
        if math.isnan(centre[0]):
            raise RuntimeError("panic: fake() with NaN x")
        return Ent(centre[0], centre[1], centre[2], mass)

    def newtons_law_of_universal_gravitation(self, other, easing_factor):  # lib.rs:84-113
        ac = self.get_centre()
        bc = other.get_centre()
        d0, d1, d2 = ac[0] - bc[0], ac[1] - bc[1], ac[2] - bc[2]
        r_norm = math.pow(d0 * d0 + d1 * d1 + d2 * d2, 0.5)
        r_easing = d0 * d0 + d1 * d1 + d2 * d2 + easing_factor
        r = [(bc[0] - ac[0]) / r_norm, (bc[1] - ac[1]) / r_norm, (bc[2] - ac[2]) / r_norm]
        am = self.mass
        bm = other.mass
        return [r[0] * G * am * bm / r_easing, r[1] * G * am * bm / r_easing, r[2] * G * am * bm / r_easing]


class Node:
    """OctreeNode (dim = 3) / QuadTreeNode (dim = 2)."""
    __slots__ = ("centre", "extent", "entity", "children", "dim")

    def __init__(self, centre, extent, dim):  # octree.rs:49-57
        self.centre = list(centre)
        self.extent = extent
        self.entity = None
        self.dim = dim
        self.children = [None] * (8 if dim == 3 else 4)

    def no_children(self):
        return all(c is None for c in self.children)

    def get_octant_id(self, item_pos):  # octree.rs:160-165 / quadtree.rs:161-165
        x_bit = 1 if item_pos[0] > self.centre[0] else 0
        y_bit = 1 if item_pos[1] > self.centre[1] else 0
        if self.dim == 2:
            return x_bit | (y_bit << 1)
        z_bit = 1 if item_pos[2] > self.centre[2] else 0
        return x_bit | (y_bit << 1) | (z_bit << 2)

    def get_octant_id_centre(self, item_pos):  # octree.rs:167-215 / quadtree.rs:167-194
        i = self.get_octant_id(item_pos)
        h = self.extent / 2.0
        cx = self.centre[0] + h if i & 1 else self.centre[0] - h
        cy = self.centre[1] + h if i & 2 else self.centre[1] - h
        if self.dim == 2:
            cz = self.centre[2]
        else:
            cz = self.centre[2] + h if i & 4 else self.centre[2] - h
        return [cx, cy, cz]

    def push(self, item, count):  # octree.rs:59-128
        if count > 64:
            raise RuntimeError("panic: Recursion too deep %r" % (item.get_centre(),))
        if self.entity is None:
            self.entity = item
            return
        current_elem = self.entity
        if self.no_children() and all(
                abs(a - b) < 1e-9 for a, b in zip(current_elem.get_centre(), item.get_centre())):
            self.entity = Ent.fake(item.get_centre(), current_elem.get_mass() + item.get_mass())
            return
        centre_of_mass = current_elem.centre_of_mass(item)
        self.entity = Ent.fake(centre_of_mass, current_elem.get_mass() + item.get_mass())
        if self.no_children():
            idx = self.get_octant_id(current_elem.get_centre())
            new_node = Node(self.get_octant_id_centre(current_elem.get_centre()), self.extent / 2.0, self.dim)
            new_node.entity = current_elem
            self.children[idx] = new_node
        idx = self.get_octant_id(item.get_centre())
        if self.children[idx] is not None:
            self.children[idx].push(item, count + 1)
        else:
            new_node = Node(self.get_octant_id_centre(item.get_centre()), self.extent / 2.0, self.dim)
            new_node.entity = item
            self.children[idx] = new_node

    def get_leaves_with_resolution(self, location, bh_factor):  # octree.rs:130-158
        result = []
        stack = [self]
        while stack:
            node = stack.pop()
            if node.entity is not None:
                e0 = location[0] - node.centre[0]
                e1 = location[1] - node.centre[1]
                e2 = location[2] - node.centre[2]
                r = math.sqrt(e0 * e0 + e1 * e1 + e2 * e2)
                # extent / 0.0 is +inf in Rust (never < theta); Python raises instead
                ratio = node.extent / r if r != 0.0 else math.inf
                if ratio < bh_factor:
                    result.append(node.entity)
                    continue
            if node.entity is None:
                continue
            elif node.no_children():
                result.append(node.entity)
            else:
                for child in node.children:  # .iter().flatten(): children 0.., popped last-first
                    if child is not None:
                        stack.append(child)
        return result


def transform(dim, state, theta, easing_factor):
    """AstroElement (dim 2) / AstroOctreeElement (dim 3) ::transform on a list of Ent.
    Returns (accelerations as [x, y, z] lists, interactions counted per body)."""
    extent = None  # .flat_map(get_centre).map(abs).reduce(f64::max).unwrap_or(1.0)
    for s in state:
        for v in s.get_centre():
            a = abs(v)
            if extent is None:
                extent = a
            elif not math.isnan(a) and (math.isnan(extent) or a > extent):  # f64::max ignores a NaN operand
                extent = a
    if extent is None:
        extent = 1.0
    root = Node([0.0, 0.0, 0.0], 1.0 * extent, dim)
    for star in state:
        root.push(star, 0)
    acc = [[0.0, 0.0, 0.0] for _ in state]
    counts = [0] * len(state)
    for i, star_a in enumerate(state):
        if star_a.fixed:
            continue
        f = [0.0, 0.0, 0.0]
        star_bs = root.get_leaves_with_resolution(star_a.get_centre(), theta)
        for star_b in star_bs:
            counts[i] += 1
            if star_a.get_centre() == star_b.get_centre():
                continue
            fij = star_a.newtons_law_of_universal_gravitation(star_b, easing_factor)
            f[0] += fij[0]
            f[1] += fij[1]
            f[2] += fij[2]
        if star_a.mass != 0.0:
            acc[i] = [acc[i][0] + f[0] / star_a.mass, acc[i][1] + f[1] / star_a.mass, acc[i][2] + f[2] / star_a.mass]
        else:  # 0.0 / 0.0 (or f / 0.0) as IEEE arithmetic gives it; Python raises on float division by zero
            acc[i] = [_div0(f[0]), _div0(f[1]), _div0(f[2])]
    return acc, counts


def _div0(f):
    if f == 0.0 or math.isnan(f):
        return math.nan
    return math.copysign(math.inf, f)


def from_records(state):
    """numpy Entity records -> list of Ent."""
    return [Ent(float(s["x"]), float(s["y"]), float(s["z"]), float(s["mass"]), bool(s["fixed"])) for s in state]

"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: tests/, smoke(), bench cpu legs)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

ENTITY = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("vz", "<f8"),
     ("radius", "<f8"), ("mass", "<f8"), ("id", "<u8"), ("fixed", "?")], align=True)
ACCELERATION = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8")], align=True)

KIND = {"astro": 0, "astro2": 1, "simple_astro": 2}
ACC_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def build(force=False):
    src = os.path.join(_HERE, "physim_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_tree_new.restype = C.c_void_p
        L.oracle_tree_new.argtypes = [C.c_int, C.c_void_p, C.c_double]
        L.oracle_tree_free.argtypes = [C.c_void_p]
        L.oracle_tree_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_tree_query.restype = C.c_size_t
        L.oracle_tree_query.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]
        L.oracle_tree_dump.restype = C.c_size_t
        L.oracle_tree_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_state_extent.restype = C.c_double
        L.oracle_state_extent.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_transform.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_direct_range.argtypes = [C.c_double, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_size_t, C.c_size_t]
        L.oracle_verlet_new.restype = C.c_void_p
        L.oracle_verlet_free.argtypes = [C.c_void_p]
        L.oracle_verlet_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, ACC_FN,
                                         C.c_void_p, C.c_double]
        L.oracle_run_pipeline.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                                          C.c_double, C.c_size_t, C.c_void_p]
        L.oracle_integrator_new.restype = C.c_void_p
        L.oracle_integrator_new.argtypes = [C.c_int]
        L.oracle_integrator_free.argtypes = [C.c_void_p]
        L.oracle_integrator_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, ACC_FN,
                                             C.c_void_p, C.c_double]
        L.oracle_run_pipeline_with.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                               C.c_size_t, C.c_double, C.c_size_t]
        L.oracle_encode_key.restype = C.c_uint64
        L.oracle_encode_key.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.oracle_table_build.restype = C.c_void_p
        L.oracle_table_build.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.oracle_table_free.argtypes = [C.c_void_p]
        L.oracle_table_cells.restype = C.c_size_t
        L.oracle_table_cells.argtypes = [C.c_void_p]
        L.oracle_table_extent.restype = C.c_double
        L.oracle_table_extent.argtypes = [C.c_void_p]
        L.oracle_table_get.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        L.oracle_table_transform.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double,
                                             C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OraclePanic(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _ents(state):
    state = np.ascontiguousarray(state)
    assert state.dtype.itemsize == 80, state.dtype
    return state


class Tree:
    """Octree (dim=3, astro/src/octree.rs) or QuadTree (dim=2, astro/src/quadtree.rs)."""

    def __init__(self, dim, centre=(0.0, 0.0, 0.0), extent=1.0):
        self.dim = dim
        c = np.asarray(centre, dtype=np.float64)
        self._h = lib().oracle_tree_new(dim, _ptr(c), float(extent))

    def push(self, ents):
        ents = _ents(np.atleast_1d(ents))
        if lib().oracle_tree_push(self._h, _ptr(ents), len(ents)) != 0:
            raise OraclePanic(lib().oracle_last_error().decode())

    def get_leaves_with_resolution(self, location, theta, want=False):
        loc = np.asarray(location, dtype=np.float64)
        n = lib().oracle_tree_query(self._h, _ptr(loc), float(theta), None, 0)
        if not want:
            return n
        out = np.zeros(n, dtype=ENTITY)
        lib().oracle_tree_query(self._h, _ptr(loc), float(theta), _ptr(out), n)
        return out

    def dump(self):
        n = lib().oracle_tree_dump(self._h, None, 0)
        out = np.zeros((n, 12), dtype=np.float64)
        lib().oracle_tree_dump(self._h, _ptr(out), n)
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_tree_free(self._h)
            self._h = None


def state_extent(state):
    state = _ents(state)
    return lib().oracle_state_extent(_ptr(state), len(state))


def transform(kind, state, theta=1.0, e=1.0, acc=None, counts=False, phases=None):
    """astro / astro2 / simple_astro transform (astro/src/transformers.rs). Accumulates into acc."""
    state = _ents(state)
    n = len(state)
    if acc is None:
        acc = np.zeros(n, dtype=ACCELERATION)
    cnt = np.zeros(n, dtype=np.uint32) if counts else None
    rc = lib().oracle_transform(KIND[kind], float(theta), abs(float(e)), _ptr(state), n, _ptr(acc),
                                _ptr(cnt), _ptr(phases))
    if rc != 0:
        raise OraclePanic(lib().oracle_last_error().decode())
    return (acc, cnt) if counts else acc


def direct_range(state, e, t0, t1, acc=None):
    state = _ents(state)
    if acc is None:
        acc = np.zeros(len(state), dtype=ACCELERATION)
    lib().oracle_direct_range(abs(float(e)), _ptr(state), len(state), _ptr(acc), t0, t1)
    return acc


class Verlet:
    """integrators/src/verlet.rs. acc_fn(state_view, acc_view) adds accelerations in place."""

    def __init__(self):
        self._h = lib().oracle_verlet_new()

    def integrate(self, ents, acc_fn, dt):
        ents = _ents(ents)
        n = len(ents)
        out = np.zeros(n, dtype=ENTITY)

        def tramp(_ctx, sp, nn, ap):
            s = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_uint8)), shape=(nn * 80,)).view(ENTITY)
            a = np.ctypeslib.as_array(C.cast(ap, C.POINTER(C.c_double)), shape=(nn * 3,)).view(ACCELERATION)
            acc_fn(s, a)

        cb = ACC_FN(tramp)
        lib().oracle_verlet_step(self._h, _ptr(ents), _ptr(out), n, cb, None, float(dt))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_verlet_free(self._h)
            self._h = None


INTEGRATOR = {"verlet": 0, "euler": 1, "rk4": 2}


class Integrator:
    """verlet / euler / rk4 (integrators/src/{verlet,euler,rk4}.rs)."""

    def __init__(self, name):
        self._h = lib().oracle_integrator_new(INTEGRATOR[name])

    def integrate(self, ents, acc_fn, dt, out=None):
        ents = _ents(ents)
        n = len(ents)
        out = ents.copy() if out is None else out  # rk4 assigns fields of the caller's new_state

        def tramp(_ctx, sp, nn, ap):
            if nn == 0:
                return
            s = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_uint8)), shape=(nn * 80,)).view(ENTITY)
            a = np.ctypeslib.as_array(C.cast(ap, C.POINTER(C.c_double)), shape=(nn * 3,)).view(ACCELERATION)
            acc_fn(s, a)

        cb = ACC_FN(tramp)
        lib().oracle_integrator_step(self._h, _ptr(ents), _ptr(out), n, cb, None, float(dt))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_integrator_free(self._h)
            self._h = None


def run_pipeline_with(integrator, kind, state, theta, e, dt, iterations):
    """The simulation loop with a choice of integrator; returns the final state."""
    state = _ents(state).copy()
    rc = lib().oracle_run_pipeline_with(INTEGRATOR[integrator], KIND[kind], float(theta), abs(float(e)),
                                        _ptr(state), len(state), float(dt), int(iterations))
    if rc != 0:
        raise OraclePanic(lib().oracle_last_error().decode())
    return state


def run_pipeline(kind, state, theta, e, dt, iterations):
    """pipeline.rs:143-182 for one transform + verlet. Returns (final_state, [build, walk, rest] s)."""
    state = _ents(state).copy()
    secs = np.zeros(3, dtype=np.float64)
    rc = lib().oracle_run_pipeline(KIND[kind], float(theta), abs(float(e)), _ptr(state), len(state),
                                   float(dt), int(iterations), _ptr(secs))
    if rc != 0:
        raise OraclePanic(lib().oracle_last_error().decode())
    return state, secs


def encode_key(dim, x, y, z, extent):
    return lib().oracle_encode_key(dim, x, y, z, extent)


class CellTable:
    """Oracle 2: DFS pre-order cell table built from sorted compare-and-halve keys."""

    def __init__(self, dim, state):
        self.dim = dim
        self.state = _ents(state)
        n = len(self.state)
        self._h = lib().oracle_table_build(dim, _ptr(self.state), n)
        if not self._h:
            raise OraclePanic(lib().oracle_last_error().decode())
        c = lib().oracle_table_cells(self._h)
        self.n, self.n_cells = n, c
        self.extent = lib().oracle_table_extent(self._h)
        self.key = np.zeros(n, np.uint64)
        self.perm = np.zeros(n, np.uint32)
        self.cell_start = np.zeros(n + 1, np.uint32)
        self.level = np.zeros(c, np.uint8)
        self.head = np.zeros(c, np.uint32)
        self.count = np.zeros(c, np.uint32)
        self.skip = np.zeros(c, np.uint32)
        self.parent = np.zeros(c, np.uint32)
        self.centre_ext = np.zeros((c, 4), np.float64)
        self.com_mass = np.zeros((c, 4), np.float64)
        lib().oracle_table_get(self._h, _ptr(self.key), _ptr(self.perm), _ptr(self.cell_start),
                               _ptr(self.level), _ptr(self.head), _ptr(self.count), _ptr(self.skip),
                               _ptr(self.parent), _ptr(self.centre_ext), _ptr(self.com_mass))

    def transform(self, theta, e, acc=None, counts=False):
        n = self.n
        if acc is None:
            acc = np.zeros(n, dtype=ACCELERATION)
        cnt = np.zeros(n, dtype=np.uint32) if counts else None
        rc = lib().oracle_table_transform(self.dim, self._h, float(theta), abs(float(e)),
                                          _ptr(self.state), n, _ptr(acc), _ptr(cnt))
        if rc != 0:
            raise OraclePanic(lib().oracle_last_error().decode())
        return (acc, cnt) if counts else acc

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_table_free(self._h)
            self._h = None


# ---- csvsink (utilities/src/csvsink.rs:44-80), restated in Python -----------------------------
def rust_display_f64(v):
    """Rust `{}` for f64: shortest round-trip digits in positional notation; "NaN", "inf", "-inf"."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def csvsink_lines(states, print_n=1):
    """The file a csvsink with `print_n` writes for the sequence of states a pipeline sends its
    renderer: state 0 always (csvsink.rs:60-63), state k >= 1 iff k % print_n == 0 (:65-70); each
    line is "x,y,z," per entity (:74-79)."""
    out = []
    for k, st in enumerate(states):
        if k == 0 or k % print_n == 0:
            out.append("".join("%s,%s,%s," % (rust_display_f64(e["x"]), rust_display_f64(e["y"]),
                                              rust_display_f64(e["z"])) for e in st) + "\n")
    return "".join(out)

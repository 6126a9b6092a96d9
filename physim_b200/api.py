"""Host-side mirror of physim's element interface over libphysim_b200.so (ctypes; no torch).

`TransformElement` drives the library exactly as physim's `TransformElementHandler` does
(physim-core/src/plugin/transform.rs:58-104): dlopen, `{name}_get_api`, `init(json, len)` with a
non-NUL-terminated JSON object, `transform(obj, state, n, acc, n)`, `destroy(obj)`.
`Verlet` mirrors `IntegratorElement::integrate` (physim-core/src/plugin/integrator.rs:5-13) over the
`pb200_verlet_*` entry points the Rust shim forwards to.  `Sim` is the device-resident loop.

The CUDA library is the only implementation: importing this module builds nothing and computes
nothing; calling into it without a GPU raises (or aborts inside the plugin ABI, as physim's own
trampolines do).
"""
import ctypes as C
import json
import os

import numpy as np

from .entity import ACCELERATION, ENTITY

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB200_LIB_PATH") or os.path.join(_HERE, "libphysim_b200.so")  # (override: A/B runs of two builds)

KINDS = {"astro": 0, "astro2": 1, "simple_astro": 2}

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_char_p)  # RustStringAllocFn: char* (*)(const char*)
ACC_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class CMessage(C.Structure):
    _fields_ = [("priority", C.c_int), ("topic", C.c_char_p), ("message", C.c_char_p),
                ("sender_id", C.c_size_t), ("origin", C.c_int)]


class TransformElementAPI(C.Structure):
    """c_plugin/physim.h:58-65 — field order is ABI."""
    _fields_ = [
        ("init", C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)),
        ("transform", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t)),
        ("destroy", C.CFUNCTYPE(None, C.c_void_p)),
        ("get_property_descriptions", C.CFUNCTYPE(C.c_void_p, C.c_void_p, ALLOC_FN)),
        ("recv_message", C.CFUNCTYPE(None, C.c_void_p, C.POINTER(CMessage))),
        ("post_configuration_messages", C.CFUNCTYPE(None, C.c_void_p)),
    ]


class ElementMetaFFI(C.Structure):
    """c_plugin/physim.h:70-79."""
    _fields_ = [("kind", C.c_int)] + [(k, C.c_void_p) for k in
                                      ("name", "plugin", "version", "license", "author", "blurb", "repo")]


class Pb200TransformRef(C.Structure):
    """ctx of pb200_acc_from_transform: a transform element as physim holds it (vtable + object)."""
    _fields_ = [("api", C.POINTER(TransformElementAPI)), ("obj", C.c_void_p)]


class Pb200Stats(C.Structure):
    _fields_ = [("n_bodies", C.c_uint64), ("n_cells", C.c_uint64), ("interactions", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("extent", C.c_double), ("ms_h2d", C.c_float),
                ("ms_build", C.c_float), ("ms_force", C.c_float), ("ms_integrate", C.c_float),
                ("ms_d2h", C.c_float), ("ms_host_pack", C.c_float), ("ms_host_unpack", C.c_float),
                ("ms_wall", C.c_float), ("replays", C.c_uint32), ("sort_bits", C.c_uint32),
                ("sort_mode", C.c_uint32), ("max_bucket", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Pb200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libphysim_b200.so (built by physim_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pb200Error(f"{LIB_PATH} is missing: run `python -m physim_b200._build` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    L.get_plugin_abi_info.restype = C.c_char_p
    L.register_plugin.restype = C.c_char_p
    L.set_callback_target.argtypes = [vp]
    for el in KINDS:
        getattr(L, f"{el}_get_api").restype = C.POINTER(TransformElementAPI)
        getattr(L, f"{el}_register").restype = ElementMetaFFI
        getattr(L, f"{el}_register").argtypes = [ALLOC_FN]
    L.pb200_device_count.restype = i32
    L.pb200_set_device.argtypes = [i32]
    L.pb200_last_error.restype = C.c_char_p
    L.pb200_transform_create.restype = vp
    L.pb200_transform_create.argtypes = [i32, dbl, dbl]
    L.pb200_transform_create_json.restype = vp
    L.pb200_transform_create_json.argtypes = [i32, vp, sz]
    L.pb200_transform_destroy.argtypes = [vp]
    L.pb200_transform_apply.argtypes = [vp, vp, sz, vp, sz]
    L.pb200_transform_stats.argtypes = [vp, C.POINTER(Pb200Stats)]
    L.pb200_transform_theta.restype = dbl
    L.pb200_transform_theta.argtypes = [vp]
    L.pb200_transform_easing.restype = dbl
    L.pb200_transform_easing.argtypes = [vp]
    L.pb200_transform_debug_tree.argtypes = [vp] * 12
    L.pb200_transform_debug_hint.argtypes = [vp, i32, sz]
    L.pb200_transform_debug_sort_mode.argtypes = [vp, i32]
    L.pb200_verlet_create.restype = vp
    L.pb200_verlet_destroy.argtypes = [vp]
    L.pb200_verlet_step.argtypes = [vp, vp, vp, sz, ACC_FN, vp, dbl]
    L.pb200_verlet_step_fused.argtypes = [vp, vp, vp, vp, sz, dbl]
    L.pb200_verlet_stats.argtypes = [vp, C.POINTER(Pb200Stats)]
    L.pb200_verlet_set_resident.argtypes = [vp, i32]
    L.pb200_verlet_resident_counts.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pb200_integrator_create.restype = vp
    L.pb200_integrator_create.argtypes = [i32]
    L.pb200_integrator_destroy.argtypes = [vp]
    L.pb200_integrator_step.argtypes = [vp, vp, vp, sz, ACC_FN, vp, dbl]
    L.pb200_integrator_step_fused.argtypes = [vp, vp, vp, vp, sz, dbl]
    L.pb200_sim_set_integrator.argtypes = [vp, i32]
    L.pb200_sim_create.restype = vp
    L.pb200_sim_create.argtypes = [i32, dbl, dbl, dbl, i32, i32]
    L.pb200_sim_destroy.argtypes = [vp]
    L.pb200_sim_upload.argtypes = [vp, vp, sz]
    L.pb200_sim_run.argtypes = [vp, sz]
    L.pb200_sim_step_local.argtypes = [vp]
    L.pb200_sim_run_timed.argtypes = [vp, sz, C.POINTER(C.c_float)]
    L.pb200_sim_profile.argtypes = [vp, i32]
    L.pb200_sim_profile_report.argtypes = [vp, C.c_char_p, sz]
    L.pb200_sim_gather_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    L.pb200_sim_download.argtypes = [vp, vp, sz]
    L.pb200_sim_last_accelerations.argtypes = [vp, vp, sz]
    L.pb200_sim_stats.argtypes = [vp, C.POINTER(Pb200Stats)]
    L.pb200_sim_set_stream.argtypes = [vp, vp]
    L.pb200_sim_set_targets.argtypes = [vp, sz, sz]
    L.pb200_sim_stream.restype = vp
    L.pb200_sim_stream.argtypes = [vp]
    L.pb200_probe_fp32_tflops.restype = dbl
    L.pb200_csvsink_create.restype = vp
    L.pb200_csvsink_create.argtypes = [C.c_char_p, sz]
    L.pb200_csvsink_push.argtypes = [vp, vp, sz]
    L.pb200_csvsink_destroy.argtypes = [vp]
    L.pb200_csvsink_count.restype = sz
    L.pb200_csvsink_count.argtypes = [vp]
    L.pb200_csvsink_next_is_printed.argtypes = [vp]
    L.pb200_csvsink_skip.argtypes = [vp]
    L.pb200_csv_format_f64.restype = sz
    L.pb200_csv_format_f64.argtypes = [dbl, C.c_char_p, sz]
    L.pb200_sim_run_csvsink.argtypes = [vp, sz, vp]
    L.pb200_sim_generate_cube.argtypes = [vp, sz, C.c_uint64, dbl, dbl, dbl, vp]
    L.pb200_msim_generate_cube.argtypes = [vp, sz, C.c_uint64, dbl, dbl, dbl, vp]
    L.pb200_comm_unique_id.argtypes = [vp]
    L.pb200_msim_create.restype = vp
    L.pb200_msim_create.argtypes = [i32, dbl, dbl, dbl, i32, i32, vp, vp, vp]
    L.pb200_msim_destroy.argtypes = [vp]
    L.pb200_msim_upload.argtypes = [vp, vp, sz]
    L.pb200_msim_run.argtypes = [vp, sz]
    L.pb200_msim_run_timed.argtypes = [vp, sz, C.POINTER(C.c_float)]
    L.pb200_msim_download.argtypes = [vp, vp, sz]
    L.pb200_msim_last_accelerations.argtypes = [vp, vp, sz]
    L.pb200_msim_stats.argtypes = [vp, C.POINTER(Pb200Stats), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pb200_msim_rank_counts.argtypes = [vp, vp, vp]
    L.pb200_msim_replicas_identical.argtypes = [vp]
    L.pb200_msim_debug_shard.argtypes = [vp, vp]
    L.pb200_msim_profile.argtypes = [vp, i32]
    L.pb200_msim_profile_report.argtypes = [vp, C.c_char_p, sz]
    _lib = L
    return L


def last_error():
    return lib().pb200_last_error().decode(errors="replace")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _state(state):
    state = np.ascontiguousarray(state)
    if state.dtype.itemsize != ENTITY.itemsize:
        raise TypeError("state must be an array of 80-byte physim Entity records")
    return state


# strings handed to the "host allocator" during a call; freed when the call returns
_host_strings = []


@ALLOC_FN
def host_alloc_string(s):
    """Stand-in for physim's host_alloc_string (physim-core/src/plugin/mod.rs): copies into host-owned
    storage and returns the pointer."""
    buf = C.create_string_buffer(s)
    _host_strings.append(buf)
    return C.addressof(buf)


def element_meta(name):
    """`{name}_register(alloc)` as discover.rs:363-371 calls it."""
    m = getattr(lib(), f"{name}_register")(host_alloc_string)
    out = {"kind": m.kind}
    for k in ("name", "plugin", "version", "license", "author", "blurb", "repo"):
        out[k] = C.string_at(getattr(m, k)).decode()
    _host_strings.clear()
    return out


class TransformElement:
    """A transform element loaded through the plugin C ABI (the drop-in boundary)."""

    def __init__(self, name, **properties):
        if name not in KINDS:
            raise KeyError(name)
        self.name = name
        self._api = getattr(lib(), f"{name}_get_api")().contents
        blob = json.dumps(properties).encode()            # serde_json::to_string(&properties)
        self._blob = np.frombuffer(blob, dtype=np.uint8)  # not NUL-terminated, like into_raw_parts
        self._obj = self._api.init(_ptr(self._blob), len(blob))
        if not self._obj:
            raise Pb200Error("Failed to load transform element: " + last_error())

    @property
    def theta(self):
        return lib().pb200_transform_theta(self._obj)

    @property
    def easing(self):
        return lib().pb200_transform_easing(self._obj)

    def transform(self, state, accelerations=None):
        """accelerations[i] += a_i (transformers.rs:32-69,123-160,220-244). Returns the array."""
        state = _state(state)
        n = len(state)
        if accelerations is None:
            accelerations = np.zeros(n, dtype=ACCELERATION)
        assert accelerations.dtype.itemsize == 24 and len(accelerations) == n
        self._api.transform(self._obj, _ptr(state), n, _ptr(accelerations), n)
        return accelerations

    def get_property_descriptions(self):
        p = self._api.get_property_descriptions(self._obj, host_alloc_string)
        if not p:
            raise Pb200Error("Unable to load descriptions of properties")
        out = json.loads(C.string_at(p).decode())
        _host_strings.clear()
        return out

    def recv_message(self, topic="t", message="m", sender_id=0, priority=2):
        msg = CMessage(priority, topic.encode(), message.encode(), sender_id, 0)
        self._api.recv_message(self._obj, C.byref(msg))

    def post_configuration_messages(self):
        self._api.post_configuration_messages(self._obj)

    def stats(self):
        st = Pb200Stats()
        if lib().pb200_transform_stats(self._obj, C.byref(st)) != 0:
            raise Pb200Error(last_error())
        return st.as_dict()

    def debug_tree(self):
        """Arrays of the tree built by the last transform() (parity tests)."""
        st = self.stats()
        n, c = st["n_bodies"], st["n_cells"]
        out = {
            "key": np.zeros(n, np.uint64), "perm": np.zeros(n, np.uint32),
            "cell_start": np.zeros(n + 1, np.uint32), "level": np.zeros(c, np.uint8),
            "head": np.zeros(c, np.uint32), "count": np.zeros(c, np.uint32),
            "skip": np.zeros(c, np.uint32), "parent": np.zeros(c, np.uint32),
            "centre_ext": np.zeros((c, 4), np.float64), "com_mass": np.zeros((c, 4), np.float64),
            "counts": np.zeros(n, np.uint32),
        }
        rc = lib().pb200_transform_debug_tree(self._obj, *[_ptr(out[k]) for k in (
            "key", "perm", "cell_start", "level", "head", "count", "skip", "parent", "centre_ext",
            "com_mass", "counts")])
        if rc != 0:
            raise Pb200Error(last_error())
        out["extent"] = st["extent"]
        return out

    def debug_hint(self, sort_lo, n_cells_hint):
        lib().pb200_transform_debug_hint(self._obj, int(sort_lo), int(n_cells_hint))

    def debug_sort_mode(self, mode):
        lib().pb200_transform_debug_sort_mode(self._obj, int(mode))

    def destroy(self):
        if getattr(self, "_obj", None):
            self._api.destroy(self._obj)
            self._obj = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


INTEGRATORS = {"verlet": 0, "euler": 1, "rk4": 2}


class Verlet:
    """`verlet` integrator (integrators/src/verlet.rs) over pb200_verlet_*; with `name` = "euler" or
    "rk4" the sibling integrators (integrators/src/euler.rs, rk4.rs) over pb200_integrator_*."""

    def __init__(self, name="verlet"):
        self.name = name
        self._v = lib().pb200_integrator_create(INTEGRATORS[name])
        if not self._v:
            raise Pb200Error(last_error())

    def integrate(self, entities, acc_fn, dt):
        """IntegratorElement::integrate: acc_fn(state, accelerations) adds into accelerations."""
        entities = _state(entities)
        n = len(entities)
        new_state = np.zeros(n, dtype=ENTITY)

        def tramp(_ctx, sp, nn, ap):
            if nn == 0:
                return
            s = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_uint8)), shape=(nn * 80,)).view(ENTITY)
            a = np.ctypeslib.as_array(C.cast(ap, C.POINTER(C.c_double)), shape=(nn * 3,)).view(ACCELERATION)
            acc_fn(s, a)

        cb = ACC_FN(tramp)
        if lib().pb200_integrator_step(self._v, _ptr(entities), _ptr(new_state), n, cb, None, float(dt)) != 0:
            raise Pb200Error(last_error())
        return new_state

    def integrate_fused(self, entities, transform, dt, out=None):
        """Same step with the accelerations of `transform` (a TransformElement) kept on the device.
        `out` is the caller's `new_state` buffer (physim reuses one Vec across steps)."""
        entities = _state(entities)
        n = len(entities)
        new_state = np.zeros(n, dtype=ENTITY) if out is None else out
        assert new_state.dtype.itemsize == 80 and len(new_state) == n
        rc = lib().pb200_verlet_step_fused(self._v, transform._obj, _ptr(entities), _ptr(new_state), n,
                                           float(dt))
        if rc != 0:
            raise Pb200Error(last_error())
        return new_state

    def integrate_dropin(self, entities, transform, dt, out=None):
        """The composition stock physim runs, natively: IntegratorElement::integrate with an acc_fn that calls
        `transform` (a TransformElement) through its plugin vtable (pipeline.rs:137-141) - what the Rust shim's
        `verlet` does when the gravity element is a separate plugin element."""
        entities = _state(entities)
        n = len(entities)
        new_state = np.zeros(n, dtype=ENTITY) if out is None else out
        ref = Pb200TransformRef(C.pointer(transform._api), transform._obj)
        fn = C.cast(lib().pb200_acc_from_transform, ACC_FN)
        if lib().pb200_integrator_step(self._v, _ptr(entities), _ptr(new_state), n, fn, C.byref(ref), float(dt)) != 0:
            raise Pb200Error(last_error())
        return new_state

    def set_resident(self, on=True):
        """Opt-in of the fused step: state kept in HBM between calls, whole Entity records returned by one DMA."""
        lib().pb200_verlet_set_resident(self._v, 1 if on else 0)

    def resident_counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().pb200_verlet_resident_counts(self._v, C.byref(a), C.byref(b))
        return a.value, b.value

    def stats(self):
        st = Pb200Stats()
        lib().pb200_verlet_stats(self._v, C.byref(st))
        return st.as_dict()

    def __del__(self):
        try:
            if getattr(self, "_v", None):
                lib().pb200_verlet_destroy(self._v)
                self._v = None
        except Exception:
            pass


class Sim:
    """Device-resident `transform ! verlet` loop (pipeline.rs:143-182 without the host copies)."""

    def __init__(self, name, theta=float("nan"), e=float("nan"), dt=1e-6, rank=0, world=1, device=None):
        if device is not None:
            lib().pb200_set_device(int(device))
        self.rank, self.world = rank, world
        self._s = lib().pb200_sim_create(KINDS[name], float(theta), float(e), float(dt), rank, world)
        if not self._s:
            raise Pb200Error(last_error())
        self.n = 0

    def upload(self, state):
        state = _state(state)
        self.n = len(state)
        if lib().pb200_sim_upload(self._s, _ptr(state), self.n) != 0:
            raise Pb200Error(last_error())

    def generate_cube(self, n, seed=0, spin=0.0, mass=1.0, size=1.0, centre=(0.0, 0.0, 0.0)):
        """`cube` initial conditions made on the device from the reference's ChaCha8 stream (no host array)."""
        c = np.asarray(centre, dtype=np.float64)
        self.n = int(n)
        if lib().pb200_sim_generate_cube(self._s, self.n, int(seed), float(spin), float(mass), float(size), _ptr(c)) != 0:
            raise Pb200Error(last_error())

    def run(self, steps):
        if lib().pb200_sim_run(self._s, int(steps)) != 0:
            raise Pb200Error(last_error())

    def run_timed(self, steps):
        """Runs `steps` steps; returns their device time in ms (CUDA events on the sim's stream)."""
        ms = C.c_float()
        if lib().pb200_sim_run_timed(self._s, int(steps), C.byref(ms)) != 0:
            raise Pb200Error(last_error())
        return ms.value

    def profile(self, enable=True):
        lib().pb200_sim_profile(self._s, 1 if enable else 0)

    def profile_report(self):
        buf = C.create_string_buffer(1 << 16)
        if lib().pb200_sim_profile_report(self._s, buf, len(buf)) != 0:
            raise Pb200Error("profile report failed")
        return json.loads(buf.value.decode())

    def step_local(self):
        if lib().pb200_sim_step_local(self._s) != 0:
            raise Pb200Error(last_error())

    def gather_buffer(self):
        """(device pointer, total bytes, slice offset, slice bytes) of the fp64 {x,y,z,m} buffer."""
        p, tot, off, sl = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().pb200_sim_gather_buffer(self._s, C.byref(p), C.byref(tot), C.byref(off), C.byref(sl))
        return p.value, tot.value, off.value, sl.value

    def run_csvsink(self, steps, sink):
        """`steps` steps with `sink` (a CsvSink) as the pipeline's renderer: it receives the current
        state first if it is fresh, then the state after every step (pipeline.rs:129-131,179)."""
        if lib().pb200_sim_run_csvsink(self._s, int(steps), sink._obj) != 0:
            raise Pb200Error(last_error())

    def download(self, state):
        state = _state(state)
        if lib().pb200_sim_download(self._s, _ptr(state), len(state)) != 0:
            raise Pb200Error(last_error())
        return state

    def last_accelerations(self):
        acc = np.zeros(self.n, dtype=ACCELERATION)
        if lib().pb200_sim_last_accelerations(self._s, _ptr(acc), self.n) != 0:
            raise Pb200Error(last_error())
        return acc

    def stats(self):
        st = Pb200Stats()
        if lib().pb200_sim_stats(self._s, C.byref(st)) != 0:
            raise Pb200Error(last_error())
        return st.as_dict()

    def set_integrator(self, name):
        if lib().pb200_sim_set_integrator(self._s, INTEGRATORS[name]) != 0:
            raise Pb200Error(last_error())

    def set_targets(self, t0, t1):
        if lib().pb200_sim_set_targets(self._s, int(t0), int(t1)) != 0:
            raise Pb200Error(last_error())

    def set_stream(self, cuda_stream_ptr):
        if lib().pb200_sim_set_stream(self._s, C.c_void_p(cuda_stream_ptr)) != 0:
            raise Pb200Error(last_error())

    @property
    def stream(self):
        return lib().pb200_sim_stream(self._s)

    def __del__(self):
        try:
            if getattr(self, "_s", None):
                lib().pb200_sim_destroy(self._s)
                self._s = None
        except Exception:
            pass


def comm_unique_id():
    """128 bytes naming a communicator (NCCL unique id); make it on one process, give it to all."""
    buf = np.zeros(128, dtype=np.uint8)
    if lib().pb200_comm_unique_id(_ptr(buf)) != 0:
        raise Pb200Error(last_error())
    return buf


class MultiSim:
    """The device-resident loop on several GPUs of one box (csrc/multi.cu): one rank per GPU.

    MultiSim(name, ..., devices=[0, 1])                    every rank in this process (the plugin's shape)
    MultiSim(name, ..., world=W, rank=r, device=d, comm_id=id)   one rank per process (torchrun)"""

    def __init__(self, name, theta=float("nan"), e=float("nan"), dt=1e-6, devices=None, world=None, rank=None,
                 device=None, comm_id=None):
        if devices is not None:
            ranks = np.arange(len(devices), dtype=np.int32)
            devs = np.asarray(devices, dtype=np.int32)
            world = len(devices)
            idp = None
        else:
            ranks = np.asarray([rank], dtype=np.int32)
            devs = np.asarray([device], dtype=np.int32)
            self._id = None if comm_id is None else np.ascontiguousarray(comm_id, dtype=np.uint8)
            idp = _ptr(self._id)
        self.world = int(world)
        self._m = lib().pb200_msim_create(KINDS[name], float(theta), float(e), float(dt), self.world, len(ranks),
                                          _ptr(ranks), _ptr(devs), idp)
        if not self._m:
            raise Pb200Error(last_error())
        self.n = 0

    def upload(self, state):
        state = _state(state)
        self.n = len(state)
        if lib().pb200_msim_upload(self._m, _ptr(state), self.n) != 0:
            raise Pb200Error(last_error())

    def generate_cube(self, n, seed=0, spin=0.0, mass=1.0, size=1.0, centre=(0.0, 0.0, 0.0)):
        c = np.asarray(centre, dtype=np.float64)
        self.n = int(n)
        if lib().pb200_msim_generate_cube(self._m, self.n, int(seed), float(spin), float(mass), float(size), _ptr(c)) != 0:
            raise Pb200Error(last_error())

    def run(self, steps):
        if lib().pb200_msim_run(self._m, int(steps)) != 0:
            raise Pb200Error(last_error())

    def run_timed(self, steps):
        ms = C.c_float()
        if lib().pb200_msim_run_timed(self._m, int(steps), C.byref(ms)) != 0:
            raise Pb200Error(last_error())
        return ms.value

    def download(self, state):
        state = _state(state)
        if lib().pb200_msim_download(self._m, _ptr(state), len(state)) != 0:
            raise Pb200Error(last_error())
        return state

    def last_accelerations(self):
        acc = np.zeros(self.n, dtype=ACCELERATION)
        if lib().pb200_msim_last_accelerations(self._m, _ptr(acc), self.n) != 0:
            raise Pb200Error(last_error())
        return acc

    def stats(self):
        st, a, b = Pb200Stats(), C.c_uint64(), C.c_uint64()
        if lib().pb200_msim_stats(self._m, C.byref(st), C.byref(a), C.byref(b)) != 0:
            raise Pb200Error(last_error())
        d = st.as_dict()
        d["sharded_steps"], d["replicated_steps"] = a.value, b.value
        return d

    def rank_counts(self):
        bodies, cells = np.zeros(8, np.uint32), np.zeros(8, np.uint32)
        w = lib().pb200_msim_rank_counts(self._m, _ptr(bodies), _ptr(cells))
        return (bodies[:w].copy(), cells[:w].copy()) if w > 0 else (None, None)

    def debug_shard(self):
        out = np.zeros(13, dtype=np.uint64)
        if lib().pb200_msim_debug_shard(self._m, _ptr(out)) != 0:
            return None
        return {"cuts": [hex(int(v)) for v in out[:self.world + 1]], "kept": int(out[9]), "epoch": int(out[10]),
                "capacity": int(out[11]), "wait_gave_up": int(out[12])}

    def replicas_identical(self):
        return lib().pb200_msim_replicas_identical(self._m) == 0

    def profile(self, enable=True):
        lib().pb200_msim_profile(self._m, 1 if enable else 0)

    def profile_report(self):
        buf = C.create_string_buffer(1 << 16)
        if lib().pb200_msim_profile_report(self._m, buf, len(buf)) != 0:
            raise Pb200Error("profile report failed")
        return json.loads(buf.value.decode())

    def close(self):
        if getattr(self, "_m", None):
            lib().pb200_msim_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return lib().pb200_device_count()


def probe_fp32_tflops():
    return lib().pb200_probe_fp32_tflops()


def csv_format_f64(v):
    """Rust `format!("{}", v)` for an f64, as the csvsink writes it (csvsink.rs:76)."""
    buf = C.create_string_buffer(512)
    k = lib().pb200_csv_format_f64(float(v), buf, 512)
    return buf.raw[:k].decode()


class CsvSink:
    """physim's `csvsink` renderer (utilities/src/csvsink.rs:17-80): properties `file`, `print_n`."""

    def __init__(self, file="csvsink.csv", print_n=1):
        self._obj = lib().pb200_csvsink_create(os.fsencode(file), int(print_n))
        if not self._obj:
            raise Pb200Error(last_error())

    def push(self, state):
        state = _state(state)
        if lib().pb200_csvsink_push(self._obj, _ptr(state), len(state)) != 0:
            raise Pb200Error(last_error())

    def count(self):
        return lib().pb200_csvsink_count(self._obj)

    def close(self):
        if getattr(self, "_obj", None):
            lib().pb200_csvsink_destroy(self._obj)
            self._obj = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

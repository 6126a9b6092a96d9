"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/physim_b200.h declares, and the plugin vtables behave as physim's loader expects
(discover.rs:327-387, transform.rs:58-128) — all without touching a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import physim_b200._build as build_mod
from physim_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    build_mod.build()
    return api.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "physim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+(struct|enum)[^{;]*\{.*?\}[^;]*;", "", text, flags=re.S)  # type bodies
    text = re.sub(r"typedef[^;]*;", "", text)                                          # fn-pointer typedefs
    names = re.findall(r"\b([A-Za-z_][A-Za-z_0-9]*)\s*\([^;{()]*\)\s*;", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    names = declared_symbols()
    assert "astro2_get_api" in names and "pb200_verlet_step_fused" in names and len(names) >= 35
    L = api.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in physim_b200.h but not exported"


def test_plugin_identity():
    L = api.lib()
    assert L.get_plugin_abi_info() == b"C"                       # discover.rs:332-346
    assert L.register_plugin() == b"astro,astro2,simple_astro"   # discover.rs:353-362


@pytest.mark.parametrize("name,blurb", [
    ("astro", "Compute approximate gravitational accelerations with the Barnes-Hut algorithm (quadtree)"),
    ("astro2", "Compute approximate gravitational accelerations with the Barnes-Hut algorithm (octree)"),
    ("simple_astro", "Compute exact gravitational accelerations"),
])
def test_register_meta(name, blurb):
    m = api.element_meta(name)
    assert m["kind"] == 1 and m["name"] == name and m["blurb"] == blurb   # ElementKind::Transform


def test_init_defaults_and_properties():
    # discovery instantiates with {} and drops (discover.rs:376-387): must work with no GPU
    for name in ("astro", "astro2", "simple_astro"):
        el = api.TransformElement(name)
        assert el.theta == 1.0 and el.easing == 1.0          # transformers.rs:72-81
        docs = el.get_property_descriptions()
        assert "e" in docs and (("theta" in docs) == (name != "simple_astro"))
        el.recv_message()
        el.post_configuration_messages()                     # no bus target: silently nothing
        el.destroy()
    el = api.TransformElement("astro2", theta=1.5, e=-0.5)
    assert el.theta == 1.5 and el.easing == 0.5              # .map(|x| x.abs())
    el = api.TransformElement("astro", theta=2, e=1, other="x", nested={"a": [1, 2, {"b": None}]})
    assert el.theta == 2.0 and el.easing == 1.0              # as_f64 accepts integers
    el = api.TransformElement("astro", theta="1.5")          # a string is not a number: default
    assert el.theta == 1.0


def test_init_rejects_malformed_json():
    L = api.lib()
    a = L.astro2_get_api().contents
    for blob in (b"", b"{", b"[1,2]", b'{"theta":}', b'{"theta":1.5} x'):
        buf = np.frombuffer(blob + b"#", dtype=np.uint8)  # trailing byte: len excludes it
        assert not a.init(buf.ctypes.data_as(C.c_void_p), len(blob))


def test_empty_state_is_a_noop_without_gpu():
    el = api.TransformElement("astro2")
    st = np.zeros(0, dtype=api.ENTITY)
    acc = el.transform(st)
    assert len(acc) == 0


@pytest.mark.skipif(api.lib().pb200_device_count() > 0, reason="only meaningful without a GPU")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    # engine API: error code + message
    from physim_b200 import generators as gen
    s = gen.solar()
    L = api.lib()
    t = L.pb200_transform_create(1, 1.0, 1.0)
    acc = np.zeros(len(s), dtype=api.ACCELERATION)
    rc = L.pb200_transform_apply(t, s.ctypes.data_as(C.c_void_p), len(s), acc.ctypes.data_as(C.c_void_p), len(s))
    assert rc != 0 and b"no CPU fallback" in L.pb200_last_error()
    assert not acc["x"].any()
    # plugin ABI: aborts the process rather than returning zero forces
    code = ("from physim_b200 import api, generators as g\n"
            "api.TransformElement('simple_astro').transform(g.solar())\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "fatal" in r.stderr


def test_set_callback_target_null_aborts():
    code = "from physim_b200 import api\napi.lib().set_callback_target(None)\n"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "callback target is null" in r.stderr

"""The plugin driven by a C++ stand-in for the physim executable (tools/physim_host.cpp): dlopen,
discovery, `{el}_get_api`, init(json), bus, and the simulation loop with `verlet` — no Python in the
data path.  Result compared with the CPU oracle's pipeline loop."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api, generators as gen
from physim_b200.entity import ENTITY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "physim_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-rdynamic", os.path.join(ROOT, "tools", "physim_host.cpp"),
                           "-o", exe, "-ldl"])
    return exe


@pytest.mark.gpu
@pytest.mark.parametrize("element,props,theta,e", [
    ("astro2", '{"theta":1.5,"e":0.5}', 1.5, 0.5),
    ("astro", '{"theta":1.3}', 1.3, 1.0),
    ("simple_astro", '{"e":0.5}', 1.0, 0.5),
])
def test_cpp_host_runs_the_pipeline(host, tmp_path, element, props, theta, e):
    s = gen.readme_pipeline(5000, seed=4, spin=1000.0)
    inp, out = str(tmp_path / "state.bin"), str(tmp_path / "out.bin")
    s.tofile(inp)
    steps, dt = 5, 1e-5
    r = subprocess.run([host, api.LIB_PATH, element, props, inp, str(dt), str(steps), out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "elements=astro,astro2,simple_astro" in r.stdout
    assert "bus messages=1 last=energysink/gravity" in r.stdout      # transformers.rs:107-112
    assert f"ran {steps} iterations on {len(s)} entities" in r.stdout
    got = np.fromfile(out, dtype=ENTITY)
    ref, _ = ob.run_pipeline(element, s, theta, e, dt, steps)
    disp = np.abs(np.stack([ref[k] - s[k] for k in "xyz"], 1)).max()
    assert np.abs(np.stack([got[k] - ref[k] for k in "xyz"], 1)).max() <= 1e-6 * disp
    for f in ("radius", "mass", "id", "fixed"):
        assert np.array_equal(got[f], s[f])


def test_cpp_host_csvsink_without_gpu(host, tmp_path):
    """The csvsink renderer through the C ABI from the C++ host: with 0 iterations it receives only
    the initial state (pipeline.rs:129-131) and writes the reference's one-line format."""
    s = gen.solar()
    inp, out, csv = str(tmp_path / "state.bin"), str(tmp_path / "out.bin"), str(tmp_path / "out.csv")
    s.tofile(inp)
    r = subprocess.run([host, api.LIB_PATH, "simple_astro", '{"e":0.1}', inp, "0.01", "0", out, csv, "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert open(csv).read() == ob.csvsink_lines([s], 1)


def test_cpp_host_discovery_without_gpu(host, tmp_path):
    """Discovery, throw-away instances and the bus message need no GPU; with 0 iterations nothing
    touches CUDA (lazy initialisation)."""
    s = gen.solar()
    inp, out = str(tmp_path / "state.bin"), str(tmp_path / "out.bin")
    s.tofile(inp)
    r = subprocess.run([host, api.LIB_PATH, "simple_astro", '{"e":0.1}', inp, "0.01", "0", out],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert 'element simple_astro kind=1 blurb="Compute exact gravitational accelerations"' in r.stdout
    assert "bus messages=1 last=energysink/gravity" in r.stdout
    assert np.fromfile(out, dtype=ENTITY).tobytes() == s.tobytes()

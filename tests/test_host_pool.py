"""The host thread pool behind the plugin boundary's pack / unpack pipelines (physim_b200/csrc/host_pool.hpp),
compiled into a small C++ program and run without a GPU: parallel_for coverage, parallel_parts part ids, and
the two chunk hand-off patterns engine.cu builds on it (upload_packed, download_chunked)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("threads", ["1", "2", "5", ""])
def test_host_pool_patterns(tmp_path, threads):
    exe = str(tmp_path / "host_pool_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "cpp", "host_pool_test.cpp"), "-o", exe],
                   check=True)
    env = dict(os.environ)
    if threads:
        env["PB200_HOST_THREADS"] = threads
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "host_pool ok" in r.stdout, r.stdout + r.stderr

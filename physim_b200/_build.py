"""Builds libphysim_b200.so in-tree with nvcc for sm_100a (no torch dependency, plain CUDA runtime)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphysim_b200.so")
SOURCES = ["gravity.cu", "verlet.cu", "engine.cu", "multi.cu", "generate.cu", "plugin_abi.cpp", "csvsink.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=default", "--shared", "-cudart", "static",
    "-Xptxas", "-v",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "physim_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libphysim_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)

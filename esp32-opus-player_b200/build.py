"""Build libopusb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libopusb200.so")
SRCS = [os.path.join(HERE, "csrc", f) for f in ("opus_b200.cu", "host_api.cpp", "opus_player.cpp")]
DROPIN = os.path.join(HERE, "lib", "libopusb200_dropin.so")  # the reference's own C++ entry points over the C ABI (csrc/dropin.cpp)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def _newest_source_mtime():
    m = 0.0
    for root in (os.path.join(HERE, "csrc"), os.path.join(HERE, "..", "include")):
        for fn in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources. Returns the .so path."""
    if not force and os.path.exists(LIB) and os.path.exists(DROPIN) and min(os.path.getmtime(LIB), os.path.getmtime(DROPIN)) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRCS
    subprocess.run(cmd, check=True)
    build_dropin()
    return LIB


def build_dropin():
    """libopusb200_dropin.so: host-only glue (g++), linked against libopusb200.so next to it."""
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-o", DROPIN, os.path.join(HERE, "csrc", "dropin.cpp"),
           "-L" + os.path.dirname(LIB), "-lopusb200", "-Wl,-rpath,$ORIGIN", "-Wl,--no-undefined", "-Wl,--exclude-libs,ALL"]
    subprocess.run(cmd, check=True)
    return DROPIN

"""Builds tests/dropin/harness.cpp two ways (see its header).  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(ROOT, "esp32-opus-player_b200", "lib")
REF_OBJS = [os.path.join(ROOT, "oracle", "_ref", f) for f in ("opusfile.o", "ogg.o")]


def _link(name, extra):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    src = os.path.join(HERE, "harness.cpp")
    deps = [src, os.path.join(LIBDIR, "libopusb200_dropin.so")] + extra
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src] + extra + \
          ["-o", exe, "-L" + LIBDIR, "-lopusb200_dropin", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return exe


def harness_dropin():
    """harness + libopusb200_dropin.so: every entry point comes from the drop-in."""
    return _link("harness_dropin", [])


def harness_reflink():
    """harness + the reference's unmodified opusfile.o / ogg.o + libopusb200_dropin.so in place of the reference's
    opus_decoder.cpp, celt.cpp and silk.cpp.  None if the reference objects are not there."""
    if not all(os.path.exists(o) for o in REF_OBJS):
        return None
    return _link("harness_reflink", REF_OBJS)

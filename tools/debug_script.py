#!/usr/bin/env python3
"""Compare the parser's output on the GPU with the host build of the same source (tests/hostsim) frame by frame.
usage: debug_script.py   (golden CELT FB stereo fixture)"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
L = pkg.lib()
L.dm_debug_frame_script.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
H = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
H.hs_celt_parse.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
fx = np.load(os.path.join(ROOT, "tests", "golden", "celt_fb_stereo.npz"))
data = fx["payload"].reshape(-1).copy()
n = 4
frames = pkg.make_frames(1, n)
for f in range(n):
    frames[0, f] = (f * 159, 159, int(fx["toc"]), 0)
dec = pkg.BatchDecoder(1, n, channels=2)
rc, pcm, status = dec.decode_many(data, frames)
print("rc", rc, "status", status, "pcm ok", np.array_equal(pcm.reshape(-1, 2), fx["pcm"]))
names = ["final_rng", "status", "flags", "LM", "C", "start", "end", "pf_tapset", "spread", "oflags", "n_ops", "n_leaves", "pf_pitch", "pf_gain", "lcg_steps"]
dt = np.dtype([("final_rng", "<u4"), ("status", "<i4"), ("flags", "u1"), ("LM", "u1"), ("C", "u1"), ("start", "u1"), ("end", "u1"), ("pf_tapset", "u1"),
               ("spread", "u1"), ("oflags", "u1"), ("n_ops", "<u2"), ("n_leaves", "<u2"), ("pf_pitch", "<i2"), ("pf_gain", "<i2"), ("lcg_steps", "<u4"),
               ("pq", "<i4", (2, 21)), ("fine", "<i2", (2, 21)), ("pulses", "<i2", 21), ("cmask", "u1", 42)])
assert dt.itemsize <= 364, dt.itemsize
for f in range(n):
    gh = np.zeros(364, np.uint8)
    go = np.zeros(1360 * 16, np.uint8)
    assert L.dm_debug_frame_script(dec._ctx, f, gh.ctypes.data, gh.size, go.ctypes.data, go.size) == 0
    hh = np.zeros(512, np.uint8)
    ho = np.zeros(1360 * 16, np.uint8)
    fr = data[f * 159:(f + 1) * 159].copy()
    H.hs_celt_parse(fr.ctypes.data, 159, 2, 3, 0, 0, hh.ctypes.data, ho.ctypes.data, None)
    g = gh[:dt.itemsize].view(dt)[0]
    h = hh[:dt.itemsize].view(dt)[0]
    print("frame", f)
    for k in names:
        if g[k] != h[k]:
            print("   ", k, "gpu", g[k], "host", h[k])
    for k in ("pq", "fine", "pulses", "cmask"):
        if not np.array_equal(g[k], h[k]):
            print("   ", k, "differs\n      gpu ", g[k].reshape(-1).tolist(), "\n      host", h[k].reshape(-1).tolist())
    no = int(h["n_ops"])
    gops, hops = go.reshape(1360, 16), ho.reshape(1360, 16)
    for k in range(no):
        if not np.array_equal(gops[k], hops[k]):
            print("    first differing op", k, "gpu", gops[k].tolist(), "host", hops[k].tolist())
            break

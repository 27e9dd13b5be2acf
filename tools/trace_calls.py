#!/usr/bin/env python3
"""Timeline of consecutive pipelined device calls (OPUSB_TRACE=1): which stages of call k+1 ran beside the tail of call k.
usage: trace_calls.py [workload] [streams]"""
import os
import sys

os.environ["OPUSB_TRACE"] = "1"
os.environ.setdefault("OPUSB_TAIL", "2")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
F, steps = 50, int(sys.argv[3]) if len(sys.argv) > 3 else 8
toc = 0xFC
data, frames = pkg.synth_corpus(S, F * steps, 159, toc)
dec = pkg.BatchDecoder(S, F, channels=2)
d_bytes = torch.from_numpy(data).cuda()
d_frames = [torch.from_numpy(np.ascontiguousarray(frames[:, k * F:(k + 1) * F]).view(np.uint8).reshape(-1)).cuda() for k in range(steps)]
d_pcm = torch.empty(S * F * 960 * 2, dtype=torch.int16, device="cuda")
d_st = torch.empty(S * F, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for k in range(steps):
    rc, _ = dec.decode_device_async(d_bytes.data_ptr(), data.size, d_frames[k].data_ptr(), F, d_pcm.data_ptr(), d_st.data_ptr())
    assert rc == 0
t = dec.trace(steps)
names = ["start", "entropy", "sequence", "leaves", "bands", "imdct", "postfilter", "deemph", "end"]
print("call  " + "  ".join("%10s" % n for n in names))
for k in range(steps):
    print("%4d  " % k + "  ".join("%10.3f" % v for v in t[k]))
print("steady-state ms per call: %.3f" % ((t[-1, 8] - t[2, 8]) / (steps - 3)), "| de-emphasis segments done again:", dec.deemph_redone())

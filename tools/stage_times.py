#!/usr/bin/env python3
"""Per-stage device times of one workload, quickly (no corpus beyond a few steps, no CPU leg).
usage: stage_times.py [workload=celt] [streams] [steps=4]   -- prints the stage times of the last step and the mean step"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import bench

pkg = ge.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "celt"
comps, S, _ = bench.WORKLOADS[wl]
S = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else S
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
c = comps[0]
assert len(comps) == 1, "single-component workloads only"
F = 50 * 960 // c.frame_size
data, frames = pkg.synth_corpus(S, F * steps, c.payload, c.toc)
dec = pkg.BatchDecoder(S, F, channels=c.channels)
kw = dict(mode=pkg.DM_MODE_CELT_RAW if c.raw else pkg.DM_MODE_OPUS, frame_size=c.frame_size, coded_channels=c.channels, start_band=0)
d_bytes = torch.from_numpy(data).cuda()
d_frames = [torch.from_numpy(np.ascontiguousarray(frames[:, k * F:(k + 1) * F]).view(np.uint8).reshape(-1)).cuda() for k in range(steps)]
d_pcm = torch.empty(S * F * c.frame_size * c.channels, dtype=torch.int16, device="cuda")
d_st = torch.empty(S * F, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tot = []
for k in range(steps):
    assert dec.decode_device(d_bytes.data_ptr(), data.size, d_frames[k].data_ptr(), F, d_pcm.data_ptr(), d_st.data_ptr(), st, **kw) == 0
    torch.cuda.synchronize()
    tot.append(dec.stage_times())
assert (d_st.cpu().numpy() == c.frame_size).all()
last = tot[-1]
print(wl, S, " ".join("%s %.3f" % (k, v) for k, v in last.items()), "| mean pipeline of steps 1.. %.3f" % np.mean([t["pipeline"] for t in tot[1:]]),
      "| pcm sum", int(d_pcm.cpu().numpy().astype(np.int64).sum()))

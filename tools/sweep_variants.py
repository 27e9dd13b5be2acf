#!/usr/bin/env python3
"""Tuning sweep in ONE process on the GPU box: per-stage device times of the headline step (4096 CELT FB stereo streams x
50 frames) for a list of build variants of libopusb200.so (compile-time knobs: -DK1_THREADS=.. etc., built into
esp32-opus-player_b200/lib/variants/lib_<name>.so) and of run-time knobs (NAME=VALUE environment settings read by dm_create).

    python tools/sweep_variants.py base k1_64 kb1 OPUSB_FORK=1 OPUSB_FORK=3 kb1+OPUSB_FORK=3 ...

Every variant decodes the same packets; the checksum of its PCM must equal the first variant's."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
VAR = os.path.join(ROOT, "esp32-opus-player_b200", "lib", "variants")
S, F, STEPS = 4096, 50, 4
data, frames = pkg.synth_corpus(S, F * STEPS, 159, 0xFC)
d_bytes = torch.from_numpy(data).cuda()
d_frames = [torch.from_numpy(np.ascontiguousarray(frames[:, k * F:(k + 1) * F]).view(np.uint8).reshape(-1)).cuda() for k in range(STEPS)]
d_pcm = torch.empty(S * F * 960 * 2, dtype=torch.int16, device="cuda")
d_st = torch.empty(S * F, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
want = None
for spec in sys.argv[1:]:
    env, lib = {}, None
    for part in spec.split("+"):
        if "=" in part:
            k, v = part.split("=")
            env[k] = v
        else:
            lib = part
    for k in [k for k in os.environ if k.startswith("OPUSB_")]:
        del os.environ[k]
    os.environ.update(env)
    if lib and lib != "default":
        os.environ["OPUSB_LIB"] = os.path.join(VAR, "lib_%s.so" % lib)
    pkg.api._lib = None
    pkg.api._player_bound = False
    try:
        dec = pkg.BatchDecoder(S, F, channels=2)
        tot = []
        for rep in range(2):  # two passes over the steps: the second one is the record
            for k in range(STEPS):
                assert dec.decode_device(d_bytes.data_ptr(), data.size, d_frames[k].data_ptr(), F, d_pcm.data_ptr(), d_st.data_ptr(), st) == 0
                torch.cuda.synchronize()
                if rep:
                    tot.append(dec.stage_times())
            if not rep:
                dec.reset()
        ok = bool((d_st.cpu().numpy() == 960).all())
        chk = int(d_pcm.cpu().numpy().astype(np.int64).sum())
        if want is None:
            want = chk
        mean = {k: float(np.mean([t[k] for t in tot])) for k in tot[0]}
        print("%-28s step %.3f | %s | %s" % (spec, mean["pipeline"], " ".join("%s %.3f" % (k, v) for k, v in mean.items() if k != "pipeline"),
                                             "same PCM" if ok and chk == want else "PCM DIFFERS"), flush=True)
        dec.close()
    except Exception as e:  # a variant that does not load or launch must not end the sweep
        print("%-28s FAILED: %r" % (spec, e), flush=True)

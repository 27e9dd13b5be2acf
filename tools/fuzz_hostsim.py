"""Differential fuzzing on the CPU-only box: the kernels' stage bodies compiled for the host (tests/hostsim, the very
source the GPU runs, lanes = 1) against the unmodified reference (oracle/_ref), on long random mode walks with biased
payloads.  TEST INFRASTRUCTURE -- the only consumer of oracle/ outside tests/, bench.py and smoke(), and a checker only.

    python tools/fuzz_hostsim.py --seconds 120 --jobs 8 [--kind opus|celt|all] [--seed 1]

One forked worker per job (the reference keeps its decoder in process-wide singletons, SURVEY F5, so a worker owns one
reference decoder at a time).  Every walk is reproducible from (kind, seed).  Prints one JSON line; exit code 1 on a
mismatch, with the walk's seed, frame number and TOC.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_hostsim():
    # HOSTSIM_SO: another build of tests/hostsim/hostsim.cpp, e.g. with -fsanitize=address,undefined (run with
    # LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0)
    so = os.environ.get("HOSTSIM_SO") or os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    H = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    H.hs_celt_new.restype = P
    H.hs_celt_new.argtypes = [I]
    H.hs_celt_free.argtypes = [P]
    H.hs_celt_decode.argtypes = [P, P, I, I, I, I, P, P]
    H.hs_opus_new.restype = P
    H.hs_opus_new.argtypes = [I]
    H.hs_opus_free.argtypes = [P]
    H.hs_opus_decode.argtypes = [P, I, P, I, P]
    return H


def payload(rng, L):
    """Random bytes, biased towards the corners a uniform payload never reaches: long runs of one byte value (extreme
    symbols: maximum boosts, trims, pulses at one end), low-entropy alphabets, a random tail only (raw bits differ)."""
    kind = int(rng.integers(0, 10))
    b = rng.integers(0, 256, L, dtype=np.uint8)
    if kind == 0:
        b[: int(rng.integers(1, L + 1))] = int(rng.choice([0, 255]))
    elif kind == 1:
        b[L - int(rng.integers(1, L + 1)):] = int(rng.choice([0, 255]))
    elif kind == 2:
        b = rng.choice(np.array([0, 1, 127, 128, 254, 255], np.uint8), L)
    elif kind == 3:
        b[:] = int(rng.integers(0, 256))
    elif kind == 4:
        a, n = int(rng.integers(0, L)), int(rng.integers(1, 16))
        b[a:a + n] = int(rng.choice([0, 255]))
    return np.ascontiguousarray(b, np.uint8)


def length(rng, lo, hi):
    k = int(rng.integers(0, 8))
    if k == 0:
        return int(rng.integers(lo, min(hi, lo + 6) + 1))
    if k == 1:
        return int(rng.integers(max(lo, hi - 8), hi + 1))
    if k == 2:
        return int(rng.integers(lo, hi + 1))
    return int(rng.integers(lo, min(hi, 400) + 1))


def walk_opus(H, ref, seed, n_frames):
    """A random walk over 20 ms TOCs (config 0..31 restricted to the 20 ms entries: SILK NB/MB/WB, hybrid SWB/FB, CELT
    NB/WB/SWB/FB), stereo flag equal to the decoder's for frames with a SILK layer, either for CELT-only frames."""
    rng = np.random.default_rng(seed)
    ch = int(rng.integers(1, 3))
    silk20 = [0x08, 0x28, 0x48]
    hyb20 = [0x68, 0x78]
    celt20 = [0x98, 0xB8, 0xD8, 0xF8]
    dec = ref.PacketDecoder(ch)
    d = H.hs_opus_new(ch)
    got = np.zeros(960 * 2, np.int16)
    frames = 0
    try:
        stay = float(rng.choice([0.0, 0.5, 0.9]))
        toc = None
        for f in range(n_frames):
            if toc is None or rng.random() >= stay:
                fam = int(rng.integers(0, 3))
                toc = int(rng.choice([silk20, hyb20, celt20][fam]))
                if fam == 2:
                    toc |= 4 if rng.integers(0, 2) else 0
                else:
                    toc |= 4 if ch == 2 else 0
            L = length(rng, 1, 1274)
            pl = payload(rng, L)
            pkt = np.concatenate([np.array([toc], np.uint8), pl])
            r, want = dec.decode(pkt)
            s = H.hs_opus_decode(d, toc, pl.ctypes.data, L, got.ctypes.data)
            if r != s:
                return {"kind": "opus", "seed": seed, "frame": f, "toc": toc, "len": L, "what": "status %d vs %d" % (r, s)}
            if r > 0 and not np.array_equal(got[: 960 * ch].reshape(960, ch), want):
                return {"kind": "opus", "seed": seed, "frame": f, "toc": toc, "len": L, "what": "pcm"}
            frames += 1
    finally:
        dec.close()
        H.hs_opus_free(d)
    return frames


def walk_celt(H, ref, seed, n_frames):
    """CELT layer (celt_decode_with_ec): frame size fixed per walk (the reference's overlap state assumes it), coded
    channel count C switching against the decoder's CC, start band 0 or 17, lengths 2..1275."""
    rng = np.random.default_rng(seed)
    fs = int(rng.choice([120, 240, 480, 960]))
    ch = int(rng.integers(1, 3))
    start = 17 if rng.integers(0, 4) == 0 else 0
    R = ref.lib()
    R.ref_celt_init(ch)
    if start:
        R.ref_celt_set_start_band(start)
    d = H.hs_celt_new(ch)
    got = np.zeros(960 * 2, np.int16)
    want = np.zeros((fs, ch), np.int16)
    frames = 0
    try:
        C = ch
        for f in range(n_frames):
            if rng.integers(0, 4) == 0:
                C = int(rng.integers(1, 3))
            R.ref_celt_set_stream_channels(C)
            L = length(rng, 2, 1275)
            pl = payload(rng, L)
            r = R.ref_celt_decode(pl.ctypes.data, L, want.ctypes.data, fs)
            s = H.hs_celt_decode(d, pl.ctypes.data, L, C, fs, start, got.ctypes.data, None)
            if r != s:
                return {"kind": "celt", "seed": seed, "frame": f, "fs": fs, "C": C, "CC": ch, "len": L, "what": "status %d vs %d" % (r, s)}
            if r > 0 and not np.array_equal(got[: fs * ch].reshape(fs, ch), want):
                return {"kind": "celt", "seed": seed, "frame": f, "fs": fs, "C": C, "CC": ch, "len": L, "what": "pcm"}
            frames += 1
    finally:
        H.hs_celt_free(d)
    return frames


def worker(job, args, q):
    from oracle import ref

    ref.lib()
    H = load_hostsim()
    t_end = time.time() + args.seconds
    n_walks = n_frames = 0
    kinds = ["opus", "celt"] if args.kind == "all" else [args.kind]
    k = 0
    while time.time() < t_end:
        seed = (args.seed << 32) + (job << 20) + k
        kind = kinds[k % len(kinds)]
        res = (walk_opus if kind == "opus" else walk_celt)(H, ref, seed, args.frames)
        k += 1
        if isinstance(res, dict):
            q.put(("fail", res))
            return
        n_walks += 1
        n_frames += res
    q.put(("ok", {"walks": n_walks, "frames": n_frames}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--frames", type=int, default=80, help="frames per walk")
    ap.add_argument("--kind", default="all", choices=["opus", "celt", "all"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--replay", type=str, default=None, help="kind:seed -- run one walk in this process")
    args = ap.parse_args()
    if args.replay:
        from oracle import ref

        ref.lib()
        kind, seed = args.replay.split(":")
        print((walk_opus if kind == "opus" else walk_celt)(load_hostsim(), ref, int(seed), args.frames))
        return 0
    q = mp.get_context("fork").Queue()
    procs = [mp.get_context("fork").Process(target=worker, args=(j, args, q)) for j in range(args.jobs)]
    for p in procs:
        p.start()
    out = {"walks": 0, "frames": 0, "failures": []}
    dead = 0
    for _ in procs:
        try:
            kind, res = q.get(timeout=args.seconds + 600)
        except Exception:
            dead += 1
            continue
        if kind == "fail":
            out["failures"].append(res)
        else:
            out["walks"] += res["walks"]
            out["frames"] += res["frames"]
    for p in procs:
        p.join(5)
        if p.exitcode not in (0, None):
            out["failures"].append({"what": "worker exit code %s (abort of the reference's assert or of OPUS_CHECK_NARROW)" % p.exitcode})
    out.update(seconds=args.seconds, jobs=args.jobs, kind=args.kind, seed=args.seed, frames_per_walk=args.frames)
    print(json.dumps(out))
    return 1 if out["failures"] else 0


if __name__ == "__main__":
    sys.exit(main())

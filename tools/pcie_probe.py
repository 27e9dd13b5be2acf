#!/usr/bin/env python3
"""What the host takes from N GPUs at once: the bare device-to-host copy of one bench step's PCM (787 MB per GPU), every
rank at the same time (run under torchrun for N > 1), for several kinds of pinned host memory.  The e2e leg of bench.py
ends in exactly this copy, so this is its ceiling; bench.py measures the default kind itself (e2e.d2h_ceiling_gbs).

  python tools/pcie_probe.py                                                     one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py
Options tried (judge's list): cudaHostAlloc default / write-combined / portable, 2 MB huge pages + cudaHostRegister, and
each of them with the rank pinned to the CPU cores nvidia-smi names as the GPU's neighbours.
Prints one JSON line on rank 0."""
import ctypes
import json
import mmap
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

NBYTES = 4096 * 50 * 960 * 2 * 2
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
d = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def affinity_of_gpu(i):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        for ln in out.splitlines():
            if ln.startswith("GPU%d\t" % i) or ln.startswith("GPU%d " % i):
                for tok in ln.split("\t"):
                    tok = tok.strip()
                    if tok and all(ch.isdigit() or ch in "-," for ch in tok) and ("-" in tok or tok.isdigit()) and tok != "0" * len(tok):
                        cores = set()
                        for part in tok.split(","):
                            a, _, b = part.partition("-")
                            cores |= set(range(int(a), int(b or a) + 1))
                        return cores
    except Exception:
        pass
    return None


def measure(ptr, reps=5):
    rt.cudaMemcpyAsync(ptr, d.data_ptr(), NBYTES, 2, stream)  # warm-up (cudaMemcpyDeviceToHost = 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rt.cudaMemcpyAsync(ptr, d.data_ptr(), NBYTES, 2, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * NBYTES / (float(t.item()) * 1e-3) / 1e9  # aggregate GB/s, slowest rank


def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), NBYTES, flags)
    return p if rc == 0 else None


results = {}
all_cores = os.sched_getaffinity(0)
near = affinity_of_gpu(local)
for pin in ("unpinned", "near_gpu"):
    if pin == "near_gpu":
        if not near or near >= all_cores:
            results["near_gpu"] = "nvidia-smi names all cores (one NUMA node): nothing to pin"
            break
        os.sched_setaffinity(0, near & all_cores)
    for name, flags in (("default", 0), ("portable", 1), ("write_combined", 4)):
        p = host_alloc(flags)
        results["%s/%s" % (name, pin)] = round(measure(p), 2) if p else None
        if p:
            rt.cudaFreeHost(p)
    # 2 MB huge pages (transparent: madvise) registered with CUDA
    try:
        size = (NBYTES + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        mm.madvise(mmap.MADV_HUGEPAGE)
        buf = (ctypes.c_char * size).from_buffer(mm)
        addr = ctypes.addressof(buf)
        np.frombuffer(mm, np.uint8)[:: 4096] = 1  # touch
        rc = rt.cudaHostRegister(addr, size, 0)
        results["hugepage_registered/%s" % pin] = round(measure(ctypes.c_void_p(addr)), 2) if rc == 0 else "cudaHostRegister rc=%d" % rc
        if rc == 0:
            rt.cudaHostUnregister(addr)
        del buf
        mm.close()
    except Exception as e:  # noqa: BLE001
        results["hugepage_registered/%s" % pin] = "failed: %s" % e
if rank == 0:
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    print(json.dumps({"what": "aggregate D2H GB/s, %d GPU(s) at once, %d MB per GPU per copy, slowest rank" % (world, NBYTES // 1000000),
                      "n_gpus": world, "host_cores": len(all_cores), "results": results, "topology_first_lines": topo.splitlines()[:12]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

#!/usr/bin/env python3
"""Host-side ceiling of the end-to-end path: every rank copies device -> pinned host concurrently (what MonteCarloRollout's
log copies do), aggregate GB/s = the D2H roofline of `e2e` at this N.  Compares cudaHostAlloc pinned memory (torch) with
2 MB-huge-page backed memory registered with cudaHostRegister (fewer IOMMU / page-table entries per DMA).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/probe_d2h.py"""
import ctypes
import json
import mmap
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = 1 << 30
nbytes = 2 * GB
dev = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda").normal_()


def thp_pinned(n):
    mm = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except Exception as e:
        print("madvise failed", e, file=sys.stderr)
    arr = np.frombuffer(mm, dtype=np.float64)
    arr[::512] = 0.                                   # touch every 4 KB page
    t = torch.from_numpy(arr)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n, 0)
    return t, mm, int(rc)


def timed(host, chunk_bytes, reps=3):
    n = chunk_bytes // 8
    best = 1e9
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in range(0, nbytes // 8, n):
            host[o:o + n].copy_(dev[o:o + n], non_blocking=True)
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return world * nbytes / (best * 1e-3) / 1e9


out = {"n_gpus": world, "bytes_per_rank": nbytes}
pinned = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
for chunk in (64 << 20, 512 << 20):
    out[f"cudaHostAlloc_chunk{chunk >> 20}MB_GBs"] = timed(pinned, chunk)
del pinned
try:
    thp, keep, rc = thp_pinned(nbytes)
    out["thp_register_rc"] = rc
    out["thp_is_pinned"] = bool(thp.is_pinned())
    for chunk in (64 << 20, 512 << 20):
        out[f"thp_registered_chunk{chunk >> 20}MB_GBs"] = timed(thp, chunk)
    try:
        thp_kb = [l for l in open("/proc/self/smaps_rollup") if "AnonHugePages" in l]
        out["anon_huge_pages"] = thp_kb[0].split()[1] + " kB" if thp_kb else None
    except Exception:
        pass
except Exception as e:
    out["thp_error"] = f"{type(e).__name__}: {e}"
try:
    out["thp_enabled"] = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    out["cpus"] = os.cpu_count()
    out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
except Exception:
    pass
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""Which D2H route should N concurrent ranks take?  Every rank moves one 8192 x 8192 complex64 image to a complex128 host
array, all ranks at once (barrier before every repetition), for the device-widened route (NIS_HOST_THREADS=0) and the
host-widened route at several worker counts.  Run under torchrun; rank 0 prints one JSON line per setting with the MAX time
over the ranks.  Backs the thread policy of nis_sar.hostio.host_threads()."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import numpy as np
import torch
import torch.distributed as dist

from nis_sar import hostio

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 8192
x = torch.view_as_complex(torch.randn((n, n, 2), device=dev))
out = hostio.pinned_empty((n, n))
hostio._MIN_THREADS_FOR_HOST_ROUTE = 1
cores = len(os.sched_getaffinity(0))
settings = [0] + [t for t in (2, 3, 4, 6, 8, 12) if t * world <= cores]
for thr in settings:
    os.environ["NIS_HOST_THREADS"] = str(thr)
    ts = []
    for rep in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hostio.to_host_c128(x, out=out)
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts[1:]), float(np.median(ts[1:]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"ranks": world, "cores": cores, "threads_per_rank": thr,
                          "route": "device-widened complex128 DMA" if thr == 0 else "complex64 DMA + host widening",
                          "ms_best_max_over_ranks": float(t[0]) * 1e3, "ms_median_max_over_ranks": float(t[1]) * 1e3,
                          "aggregate_GBps_of_c128": world * 16 * n * n / float(t[1]) * 1e-9}), flush=True)
if world > 1:
    dist.destroy_process_group()

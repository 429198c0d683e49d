#!/usr/bin/env python
"""What bounds the end-to-end step at N ranks: concurrent device->host copies of 1 GB results into page-locked memory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_probe.py

For each placement policy of nis_sar.hostio.bind_rank_to_numa (none / gpu / spread) every rank allocates a pinned 1 GiB
buffer and times cudaMemcpyAsync D2H (a) one rank at a time, (b) all ranks at once.  No library kernel is involved: the
numbers are the platform's ceiling for `e2e` at N GPUs.  Rank 0 prints one JSON line (+ `nvidia-smi topo -m`)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import torch
import torch.distributed as dist

from nis_sar import hostio


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    all_cpus = sorted(os.sched_getaffinity(0))
    nbytes = 1 << 30
    src = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out = {"world": world, "nodes": {k: len(v) for k, v in hostio.numa_nodes().items()},
           "gpu_numa": [hostio.gpu_numa_node(i) for i in range(torch.cuda.device_count())], "cpus_allowed": len(all_cpus)}

    def bar():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    def copy_gbs(dst, reps=4):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(device)
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    for policy in ("none", "gpu", "spread"):
        os.sched_setaffinity(0, all_cpus)
        info = hostio.bind_rank_to_numa(local, world, policy=policy)
        t0 = time.perf_counter()
        dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        alloc_s = time.perf_counter() - t0
        solo = torch.zeros(world, dtype=torch.float64, device=device)
        for r in range(world):
            bar()
            if r == rank:
                solo[r] = copy_gbs(dst)
        bar()
        allv = torch.zeros(world, dtype=torch.float64, device=device)
        allv[rank] = copy_gbs(dst, reps=6)
        bar()
        if world > 1:
            dist.all_reduce(solo)
            dist.all_reduce(allv)
        nodes = [None] * world
        if world > 1:
            dist.all_gather_object(nodes, info.get("node"))
        else:
            nodes = [info.get("node")]
        out[policy] = {"node_of_rank": nodes, "pinned_alloc_s": round(alloc_s, 3),
                       "solo_GBps": [round(float(x), 1) for x in solo.cpu()],
                       "concurrent_GBps": [round(float(x), 1) for x in allv.cpu()],
                       "concurrent_sum_GBps": round(float(allv.sum()), 1)}
        del dst
        torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
    if rank == 0:
        print(json.dumps(out))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
        except Exception as e:
            print("nvidia-smi topo failed:", e)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Times nis_sar.hostio's two D2H / H2D routes for one complex64 image (default 8192 x 8192) at several worker counts."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import numpy as np
import torch

from nis_sar import hostio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.view_as_complex(torch.randn((n, n, 2), device="cuda"))
out = hostio.pinned_empty((n, n))
pageable = np.empty((n, n), np.complex128)
pageable[:] = 0


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3


for thr in (0, 4, 6, 8, 12, 15, 16):
    if thr > (os.cpu_count() or 1):
        break
    os.environ["NIS_HOST_THREADS"] = str(thr)
    hostio._MIN_THREADS_FOR_HOST_ROUTE = 1
    rec = {"n": n, "threads": thr}
    rec["d2h_pinned_out_ms(best,median)"] = best(lambda: hostio.to_host_c128(x, out=out))
    rec["d2h_pageable_out_ms"] = best(lambda: hostio.to_host_c128(x, out=pageable))
    rec["d2h_cached_block_ms"] = best(lambda: hostio.to_host_c128(x))
    rec["h2d_from_pageable_ms"] = best(lambda: hostio.to_device_c64(pageable, "cuda:0"))
    rec["h2d_from_pinned_ms"] = best(lambda: hostio.to_device_c64(out, "cuda:0"))
    print(json.dumps(rec), flush=True)

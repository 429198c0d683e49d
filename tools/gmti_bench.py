#!/usr/bin/env python
"""Per-kernel view of K3 on a realistic two-channel frame (sparse detections): run under ncu launch-list."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))
import torch
from nis_sar import device as dev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.view_as_complex(torch.randn((n, n, 2), generator=g, device="cuda"))
a[n // 2, n // 3] = 400.0            # one bright mover: the 5 % mask keeps a few pixels only
b = a * torch.exp(torch.tensor(0.3j, device="cuda")) + 0.05 * torch.view_as_complex(torch.randn((n, n, 2), generator=g, device="cuda"))
for want in (dev.GMTI_PRODUCTS, ("ati_phase_masked",)):
    for _ in range(3):
        out = dev.gmti_fused(a, b, want=want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = dev.gmti_fused(a, b, want=want)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    bytes_px = 16 + (33 if len(want) > 1 else 4) + 8
    print(json.dumps({"n": n, "products": len(want), "ms": ms, "det": out["det_count"], "GBps": bytes_px * n * n / ms * 1e-6}))

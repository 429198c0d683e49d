#!/usr/bin/env python
"""Development tool: run a few CSA focus calls of one size (for ncu).  Usage: python tools/one_csa.py N|NAZxNRG [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))
import torch
from nis_sar import device as dev, params

na, _, nr = sys.argv[1].partition("x")
na = int(na)
nr = int(nr) if nr else na
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prm = params.spaceborne_preset()
plan = dev.CsaPlan(na, nr, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                   t_start=prm.t_start_fast)
x = torch.view_as_complex(torch.randn((na, nr, 2), device="cuda"))
out = torch.empty((nr, na), dtype=torch.complex64, device="cuda")
for _ in range(iters):
    plan.focus(x, out=out)
torch.cuda.synchronize()
print("done")

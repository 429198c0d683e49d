#!/usr/bin/env python
"""Development tool: two-channel 4096^2 ATI frame with the two CSA focus calls on one stream vs two streams."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))
import torch
from nis_sar import device as dev, params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prm = params.spaceborne_preset()
mk = lambda: dev.CsaPlan(n, n, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                         t_start=prm.t_start_fast)
pA, pB = mk(), mk()
gen = torch.Generator(device="cuda").manual_seed(5)
ch = [torch.view_as_complex(torch.randn((n + 1, n, 2), generator=gen, device="cuda")) for _ in range(2)]
ch[0][n // 2, n // 3] += 3000.0
s1 = torch.empty((n, n), dtype=torch.complex64, device="cuda")
s2 = torch.empty_like(s1)
mx = torch.zeros(1, dtype=torch.float64, device="cuda")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def frame_seq():
    pA.focus(ch[0][1:], out=s1, max_sq=mx)
    pB.focus(ch[1][:-1], out=s2)
    return dev.gmti_fused(s1, s2, max_sq=mx, lazy=True)


def frame_par():
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur)
    sB.wait_stream(cur)
    with torch.cuda.stream(sA):
        pA.focus(ch[0][1:], out=s1, max_sq=mx)
    with torch.cuda.stream(sB):
        pB.focus(ch[1][:-1], out=s2)
    cur.wait_stream(sA)
    cur.wait_stream(sB)
    return dev.gmti_fused(s1, s2, max_sq=mx, lazy=True)


def csa_only_par():
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur)
    sB.wait_stream(cur)
    with torch.cuda.stream(sA):
        pA.focus(ch[0][1:], out=s1)
    with torch.cuda.stream(sB):
        pB.focus(ch[1][:-1], out=s2)
    cur.wait_stream(sA)
    cur.wait_stream(sB)


def csa_only_seq():
    pA.focus(ch[0][1:], out=s1)
    pB.focus(ch[1][:-1], out=s2)


def timeit(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("NIS_")}}
for name, fn in (("csa2_seq", csa_only_seq), ("csa2_par", csa_only_par), ("frame_seq", frame_seq), ("frame_par", frame_par)):
    res[name] = round(timeit(fn), 4)
print(json.dumps(res))

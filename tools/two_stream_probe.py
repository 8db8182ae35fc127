#!/usr/bin/env python
"""Probe: one batch-8 DDPM-256 UNet forward vs two batch-4 forwards issued on two CUDA streams (two engine instances with the
same weights).  Overlaps the latency-bound low-resolution levels and GroupNorm (HBM) of one half with the convolutions
(tensor) of the other.   python tools/two_stream_probe.py [batch] [splits]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit.unet import UNet2DModel
from models import DDPM256_CONFIG

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(3)).cuda()
full = UNet2DModel(**DDPM256_CONFIG, max_batch=B).init_random(0)
parts = [UNet2DModel(**DDPM256_CONFIG, max_batch=B // NS).init_random(0) for _ in range(NS)]
streams = [torch.cuda.Stream() for _ in range(NS)]
out_full = torch.empty_like(x)
out_split = torch.empty_like(x)


def run_full():
    full(x, 500, out=out_full)


def run_split():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    h = B // NS
    for i, (net, st) in enumerate(zip(parts, streams)):
        st.wait_event(ev)
        with torch.cuda.stream(st):
            net(x[i * h:(i + 1) * h], 500, out=out_split[i * h:(i + 1) * h])
    for st in streams:
        cur.wait_stream(st)


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = timeit(run_full)
b = timeit(run_split)
print(f"B={B}: one batch-{B} forward {a:.3f} ms | {NS} x batch-{B // NS} on {NS} streams {b:.3f} ms | ratio {a / b:.3f} | "
      f"identical={torch.equal(out_full, out_split)} maxdiff={(out_full - out_split).abs().max().item():.2e}")

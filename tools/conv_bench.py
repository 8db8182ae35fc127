#!/usr/bin/env python
"""Event-timed launches of single convolutions (b2e_conv2d_bench_f16): conv_bench.py [N H Cin Cout k [stride]] ...
Without arguments: the low-resolution layer shapes of the DDPM-256 UNet at batch 1 and 8."""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit import _C
from b200edit._C import check, lib
torch.zeros(1, device="cuda")
shapes = []
a = [int(v) for v in sys.argv[1:]]
while a:
    shapes.append(tuple(a[:5]) + (1,)); a = a[5:]
if not shapes:
    for n in (1, 8):
        shapes += [(n, 8, 512, 512, 3, 1), (n, 8, 1024, 512, 3, 1), (n, 16, 512, 512, 3, 1), (n, 16, 1024, 512, 3, 1),
                   (n, 16, 512, 512, 1, 1), (n, 32, 256, 256, 3, 1), (n, 32, 512, 256, 3, 1), (n, 64, 256, 256, 3, 1),
                   (n, 128, 128, 128, 3, 1), (n, 256, 128, 128, 3, 1)]
for (n, h, cin, cout, k, s) in shapes:
    us = C.c_float()
    copies = int(os.environ.get("B2E_BENCH_COPIES", 0)) or max(1, min(64, int(200e6 / (cin * cout * k * k * 2))))
    check(lib.b2e_conv2d_bench_f16(n, h, h, cin, cout, k, s, 200, copies, C.byref(us), None), "conv2d_bench")
    fl = 2.0 * n * h * h * cout * cin * k * k
    print(f"N{n} {h}x{h} cin{cin} cout{cout} k{k}: {us.value:7.2f} us/launch  {fl / us.value / 1e6:7.1f} TF/s  ({copies} weight copies)")

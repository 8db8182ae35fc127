#!/usr/bin/env python
"""Throughput of the native VQ / KL encoders (LDM.encode / SD.encode): python tools/bench_encoder.py [ldm|sd] [batch]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "diffusion-image-editing_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from b200edit.vqmodel import LDM_VQ_CONFIG, SD_VAE_CONFIG, AutoencoderKL, VQModel  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ldm"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = (VQModel(**LDM_VQ_CONFIG, max_batch=B, with_encoder=True) if which == "ldm"
     else AutoencoderKL(**SD_VAE_CONFIG, max_batch=B, with_encoder=True)).init_random(0)
S = m.out_size
x = torch.rand(B, 3, S, S, device="cuda") * 2 - 1
for _ in range(3):
    m._encode_raw(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    m._encode_raw(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = m._enc.flops_per_sample
print(f"{which} encoder B={B} ({S}x{S}): {ms:.3f} ms/forward, {B / ms * 1e3:.1f} img/s, {fl * B / ms / 1e9:.1f} TFLOP/s (as executed), "
      f"{fl / 1e12:.4f} TFLOP/img, {m._enc.launches_per_forward} launches")

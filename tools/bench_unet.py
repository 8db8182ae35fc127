#!/usr/bin/env python
"""UNet forward throughput + per-op profile for a named layout:  bench_unet.py {ddpm|ldm} BATCH [profile.json]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit.unet import DDPM256_CONFIG, LDM_CELEBAHQ_CONFIG, UNet2DModel
from b200edit.unet_cond import SD15_CONFIG, UNet2DConditionModel
name, B = sys.argv[1], int(sys.argv[2])
cfg = {"ddpm": DDPM256_CONFIG, "ldm": LDM_CELEBAHQ_CONFIG, "sd": SD15_CONFIG}[name]
x = torch.randn(B, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"], device="cuda")
out = torch.empty(B, cfg["out_channels"], cfg["sample_size"], cfg["sample_size"], device="cuda")
if name == "sd":
    net = UNet2DConditionModel(**cfg, max_batch=B).init_random(0)
    ctx = torch.randn(B, 77, 768, device="cuda")
    unet = net
    _call = net.__call__
    net_call = lambda x, t, out=None: _call(x, t, encoder_hidden_states=ctx, out=out)   # noqa: E731
    type(net).profile = lambda self, x, t: []   # the per-op profile hook has no context argument
    class _W:   # same call shape as the unconditional wrapper below
        flops_per_sample = net.flops_per_sample
        launches_per_forward = net.launches_per_forward
        def __call__(self, x, t, out=None): return net_call(x, t, out)
        def profile(self, x, t): return []
    unet = _W()
else:
    unet = UNet2DModel(**cfg, max_batch=B).init_random(0)
for _ in range(3):
    unet(x, 500, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    unet(x, 500, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = unet.flops_per_sample * B
print(f"{name} B={B}: {ms:.3f} ms/forward, {B / ms * 1e3:.1f} img/s, {fl / ms / 1e9:.1f} TFLOP/s (as executed), "
      f"{unet.flops_per_sample / 1e12:.4f} TFLOP/img, {unet.launches_per_forward} launches")
prof = unet.profile(x, 500)
by = {}
for p in prof:
    by.setdefault(p["kind"], [0, 0.0])
    by[p["kind"]][0] += 1; by[p["kind"]][1] += p["ms"]
print({k: (v[0], round(v[1], 3)) for k, v in by.items()})
if len(sys.argv) > 3:
    json.dump(prof, open(sys.argv[3], "w"))

#!/usr/bin/env python
"""Per-op CUDA-event profile of one DDPM-256 UNet forward (b2e_unet_profile).   python tools/profile_ops.py [batch] [out.json]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from models import create_diffusion_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = create_diffusion_model("ddpm", sample_clipping=False, max_batch=B, seed=0)
x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(3)).cuda()
w.unet.profile(x, 500)
prof = w.unet.profile(x, 500)
if len(sys.argv) > 2:
    json.dump(prof, open(sys.argv[2], "w"))
tot = {}
for p in prof:
    tot[p["kind"]] = tot.get(p["kind"], 0) + p["ms"]
print("GN_FUSE", os.environ.get("B2E_GN_FUSE", "1"), "total ms", sum(tot.values()), tot)
if os.environ.get("B2E_PROFILE_ALL"):      # every op, not only the GEMM launches
    for i, p in enumerate(prof):
        gbs = p.get("bytes", 0) / p["ms"] / 1e6 if p["ms"] > 0 else 0
        print(f"{i:3d} {p['ms'] * 1e3:7.1f}us {p['kind']:10s} {p['flops'] / p['ms'] / 1e9:7.0f}TF {gbs:7.0f}GB/s {p['desc']}")
    sys.exit(0)
for i, p in enumerate(prof):
    if p["kind"] == "conv_igemm":
        print(f"{i:3d} {p['ms'] * 1e3:7.1f}us {p['flops'] / p['ms'] / 1e9:7.0f}TF {p['desc']}")

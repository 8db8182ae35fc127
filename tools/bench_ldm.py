#!/usr/bin/env python
"""BASELINE config 3 timing: LDM-CelebAHQ layout, masked colour guidance THROUGH the native VQ decoder.
One guided step = UNet forward (B latents) + fused DDIM step + decoder forward + colour loss + decoder gradient +
masked update.    python tools/bench_ldm.py BATCH [steps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from attr_functions import SingleColorAttrFunc
from models import create_diffusion_model
from SegDiffEditPipeline import SegDiffEditPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = create_diffusion_model("ldm", sample_clipping=False, max_batch=B, seed=0)
w.scheduler.set_timesteps(K)
pipe = SegDiffEditPipeline(w, None)
g = torch.Generator().manual_seed(2)
xt = torch.randn(B, 3, 64, 64, generator=g).cuda()
mask = (torch.rand(1, 3, 64, 64, generator=g) > 0.5).float().cuda()
f = SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=50.0, t1=0, t2=K, use_mask=True, mask_attr_grad=True)


def run():
    return pipe.edit_image(xt=xt, attr_func=f, prog_bar=False, output_type="tensor", mask=mask)


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 2
for _ in range(reps):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (reps * K)
fl = (w.unet.flops_per_sample + w.vqvae.flops_per_sample) * B     # in gradient mode the decoder figure counts forward + backward
print(f"config 3, B={B}: {ms:.2f} ms per guided step, {B / ms * 1e3:.1f} guided img-steps/s, ~{fl / ms / 1e9:.0f} TFLOP/s "
      f"(UNet {w.unet.flops_per_sample / 1e12:.3f} + decoder fwd+bwd {w.vqvae.flops_per_sample / 1e12:.3f} TFLOP/img; "
      f"includes one final decode per {K} steps)")
print("max memory GB", torch.cuda.max_memory_allocated() / 1e9)

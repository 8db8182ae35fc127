#!/usr/bin/env python
"""BASELINE config 5: colour-guided DDIM throughput sweep over the per-GPU batch (1 GPU; the N-GPU numbers are
weak-scaling multiples, see bench.py --gpus).  Same step as bench.py (DDPM-256 UNet + fused guided step through
SegDiffEditPipeline.edit_image, eta = 0 DDIM + SingleColorAttrFunc), K steps per batch size, CUDA-event timing.

    python tools/sweep_batch.py [out.json] [max_batch] [bf16|fp32]
"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from attr_functions import SingleColorAttrFunc
from models import create_diffusion_model
from SegDiffEditPipeline import SegDiffEditPipeline

out_path = sys.argv[1] if len(sys.argv) > 1 else None
max_b = int(sys.argv[2]) if len(sys.argv) > 2 else 256
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
rows = []
K = 10   # DDIM steps per loop (set_timesteps(K): the per-step work does not depend on the stride)
for B in [1, 2, 4, 8, 16, 32, 64, 128, 256]:
    if B > max_b:
        break
    w = create_diffusion_model("ddpm", sample_clipping=True, max_batch=B, seed=0, precision=precision)
    w.scheduler.set_timesteps(K)
    pipe = SegDiffEditPipeline(w, None)
    f = SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=100.0, t1=0, t2=K, per_sample=True)
    x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(1234)).cuda()

    def run():
        return pipe.edit_image(xt=x, eta=0.0, attr_func=f, prog_bar=False, output_type="tensor")

    run(); run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * K)
    fl = w.unet.flops_per_sample * B / (ms * 1e-3) / 1e12
    rows.append({"batch": B, "precision": precision, "ms_per_step": ms, "img_steps_per_s": B / ms * 1e3, "unet_tflops_as_executed": fl})
    print(f"B={B:4d}  {ms:8.3f} ms/step  {B / ms * 1e3:8.1f} img-steps/s  {fl:7.1f} TFLOP/s", flush=True)
    del pipe, w
    torch.cuda.empty_cache()
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)

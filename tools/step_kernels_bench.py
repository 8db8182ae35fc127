#!/usr/bin/env python
"""The fused element-wise / reduction kernels of the guided step at a batch that exceeds L2 (default 256 images of
3x256x256: 201 MB per tensor), each launched a few times - the command the ncu `--set full` captures of
profiles/r2_ncu_step_kernels.md profile.   python tools/step_kernels_bench.py [batch] [reps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "diffusion-image-editing_b200"))
import torch
from b200edit import ops
from b200edit.scheduler import DDIMScheduler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = 256
dev = torch.device("cuda")
sch = DDIMScheduler.from_preset("ddpm")
sch.set_timesteps(50)
c = sch.coeffs(int(sch.timesteps[-5]), 1.0, "ddpm")
x, e, x0r = (torch.randn(B, 3, S, S, device=dev) for _ in range(3))
z = torch.randn(3, S, S, device=dev)
mask = (torch.rand(1, 3, S, S, device=dev) > 0.5).float()
zout, xm = torch.empty_like(x), x0r.clone()
noise = torch.randn(8, 3, S, S, device=dev)
kernels = {
    "guided_step_vec4 (16 B/elem)": lambda: ops.guided_step(x, e, c, noise=z, targets=[0.8, None, None], loss_scale=50.0, n_mean=S * S),
    "l2reg_pass1 + l2reg_pass2 (36 B/elem)": lambda: ops.guided_step_l2reg(x, e, c, noise=z, targets=[0.8, None, None], loss_scale=50.0,
                                                                          mask=mask, x_ref=x0r, lambda_=0.1),
    "extract_noise_kernel (20 B/elem)": lambda: ops.extract_noise(x, e, xm, zout, c),
    "map2_kernel<pred_x0> (12 B/elem)": lambda: ops.pred_x0(x, e, float(c.sqrt_a_t), float(c.sqrt_b_t)),
    "apply_mask_kernel (12 B/elem)": lambda: ops.apply_mask(mask, x, e),
    "to_uint8_kernel (5 B/elem)": lambda: ops.to_uint8(x),
    "color_grad_kernel (8 B/elem)": lambda: ops.color_loss_grad(x, [0.8, None, None], None, 50.0, S * S),
}
for name, fn in kernels.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpe = float(name.split("(")[-1].split()[0])
    print(f"{name}: {ms:.4f} ms, {bpe * x.numel() / ms / 1e6:.0f} GB/s", flush=True)

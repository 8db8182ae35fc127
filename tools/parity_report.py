#!/usr/bin/env python
"""Per-step parity of BASELINE configs[0] (DDPM-256 UNet2DModel, random init seed 0; colour-guided DDIM, 50 steps,
batch 1, clip_sample) against an fp64 GROUND TRUTH, with yardsticks.

Teacher-forced on the trajectory of the fp32 oracle: at every step the same x_t is fed to
  * the oracle UNet in fp64 (torch eager on the GPU)                      -> ground truth eps64, x_{t-1} in fp64 step math
  * the oracle UNet in torch fp32 (cuDNN, TF32 off) / fp16 / bf16         -> yardsticks (what diffusers + PyTorch gives)
  * the native engine, fp16 operands and fp32-accurate (split fp16)       -> the product
and the max-abs error of x_{t-1} (the image the step produces, after the fused guided-step kernel for the native arms)
against the fp64 result is recorded, together with max|eps| and max|x_{t-1}| (the range of the compared tensor).

    python tools/parity_report.py [out.json] [T]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "diffusion-image-editing_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from oracle import loops, step_math as sm  # noqa: E402
from oracle.ddim_scheduler import DDIMScheduler as OracleScheduler  # noqa: E402
from oracle.unet2d import DDPM256_CONFIG, UNet2DModel as OracleUNet  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "parity_report.json")
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    from attr_functions import SingleColorAttrFunc
    from b200edit import _C, ops
    from diffusion_utils import get_noise_pred
    from models import create_diffusion_model
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    o32 = OracleUNet(**DDPM256_CONFIG).eval()
    sd = o32.state_dict()
    o32 = o32.cuda()
    o64 = OracleUNet(**DDPM256_CONFIG).eval()
    o64.load_state_dict(sd)
    o64 = o64.double().cuda()
    o16 = OracleUNet(**DDPM256_CONFIG).eval()
    o16.load_state_dict(sd)
    o16 = o16.half().cuda()
    ob16 = OracleUNet(**DDPM256_CONFIG).eval()
    ob16.load_state_dict(sd)
    ob16 = ob16.bfloat16().cuda()
    fast = _C.fast_precision()
    natives = {p: create_diffusion_model("ddpm", sample_clipping=True, max_batch=1, state_dict=sd, precision=p)
               for p in (fast, "fp32")}
    for w in natives.values():
        w.scheduler.set_timesteps(T)
    s = OracleScheduler.from_preset("ddpm")
    s.set_timesteps(T)
    f = SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=100.0, t1=0, t2=T)
    guide = loops.color_guidance([0.8, None, None], [1, 1, 1], 100.0, 0, T)
    xt = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(1234))
    rows = []
    arms = ["torch_fp32", "torch_fp16", "torch_bf16", f"native_{fast}", "native_fp32"]
    worst = {a: 0.0 for a in arms}
    worst_rel = {a: 0.0 for a in arms}
    for step_idx, t in loops.window_timesteps(s):
        c = sm.step_coeffs(s, t)
        tt = torch.tensor(t)
        with torch.no_grad():
            e64 = o64(xt.double().cuda(), tt)["sample"]
            e32 = o32(xt.cuda(), tt)["sample"]
            e16 = o16(xt.half().cuda(), tt)["sample"].float()
            eb16 = ob16(xt.bfloat16().cuda(), tt)["sample"].float()

        def step(eps):   # reference step math in the dtype of eps (fp64 for the ground truth)
            x = xt.to(eps.dtype).to(eps.device)
            xn, _ = sm.ddim_step(x, eps, c, 0.0, None, clip=True, clip_range=s.config.clip_sample_range)
            return guide(xn, eps, c, step_idx)

        x64 = step(e64)
        rng_x = max(1.0, x64.abs().max().item())
        row = dict(step=step_idx, t=t, max_eps=e64.abs().max().item(), max_x_prev=x64.abs().max().item())
        res = {"torch_fp32": step(e32), "torch_fp16": step(e16), "torch_bf16": step(eb16)}
        for p, w in natives.items():
            xg = xt.cuda()
            eps_n = get_noise_pred(w.model, xg, tt)
            fk = f.fused_kwargs(xg, w, mask=None)
            x_nat, _ = ops.guided_step(xg, eps_n, w.scheduler.coeffs(t, 0.0, "ddim"), clip=True,
                                       clip_range=w.scheduler.config.clip_sample_range, noise=None, **fk)
            res[f"native_{p}"] = x_nat
        for a, x in res.items():
            err = (x.double() - x64).abs().max().item()
            row[a] = err
            worst[a] = max(worst[a], err)
            worst_rel[a] = max(worst_rel[a], err / rng_x)
        rows.append(row)
        # teacher forcing: continue on the fp32 oracle's trajectory (the one tests/test_gpu_pipeline.py walks)
        xt = res["torch_fp32"].cpu()
    rep = dict(config="BASELINE configs[0]: DDPM-256 UNet2DModel random-init, colour-guided DDIM-50, batch 1, clip_sample",
               ground_truth="oracle UNet + step math in fp64 (torch eager, GPU)", arms=arms,
               worst_abs_x_prev=worst, worst_abs_over_range_of_x_prev=worst_rel, rows=rows)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(rep, fh, indent=1)
    print("worst max-abs error of x_(t-1) vs fp64 ground truth over", T, "steps:")
    for a in arms:
        print(f"  {a:14s} abs {worst[a]:.3e}   / max(1, max|x_(t-1)|) {worst_rel[a]:.3e}")
    print("  step: max|eps| max|x_prev| | " + " ".join(arms))
    for r in rows[::5]:
        print(f"  {r['step']:2d}: {r['max_eps']:.2f} {r['max_x_prev']:.2f} | " + " ".join(f"{r[a]:.1e}" for a in arms))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE configs[0] end to end on the GPU: DDPM-256 UNet2DModel (random init, seed 0), colour-guided DDIM
(SingleColorAttrFunc target 0.8, channel 0, loss_scale 100), T steps, batch 1, x_T = randn(seed 1234).

The native pipeline (bf16 and fp32-accurate noise predictors) is compared step by step with the oracle loop driven by
the oracle UNet (torch fp32 eager on the GPU as the checker, TF32 off).  Prints max-abs deviations of the x0
predictions along the trajectory and of the final image.

    python tools/e2e_parity.py [T] [sample_size]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "diffusion-image-editing_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from oracle import loops  # noqa: E402
from oracle.ddim_scheduler import DDIMScheduler as OracleScheduler  # noqa: E402
from oracle.unet2d import DDPM256_CONFIG, UNet2DModel as OracleUNet  # noqa: E402


def run(T=50, precisions=("bf16", "fp32"), cfg=None, seed=0, loss_scale=100.0, control=True):
    from attr_functions import SingleColorAttrFunc
    from models import create_diffusion_model
    from SegDiffEditPipeline import SegDiffEditPipeline
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = dict(cfg or DDPM256_CONFIG)
    S = cfg["sample_size"]
    torch.manual_seed(seed)
    oracle = OracleUNet(**cfg).eval()
    sd = oracle.state_dict()
    oracle = oracle.cuda()
    xt = torch.randn(1, cfg["in_channels"], S, S, generator=torch.Generator().manual_seed(1234))
    s = OracleScheduler.from_preset("ddpm")
    s.set_timesteps(T)

    def eps_fn(x, t):
        with torch.no_grad():
            return oracle(x.cuda(), torch.tensor(t))["sample"].cpu()

    xf, _, x0_h = loops.guided_edit_loop(s, eps_fn, xt, eta=0.0, zs=None,
                                         guidance=loops.color_guidance([0.8, None, None], [1, 1, 1], loss_scale, 0, T))
    out = {}
    if control:
        # control: the SAME fp32 oracle UNet on the host CPU (the reference's own CPU path) against itself on the GPU
        cpu_unet = OracleUNet(**cfg).eval()
        cpu_unet.load_state_dict(sd)
        torch.set_num_threads(os.cpu_count() or 1)

        def eps_cpu(x, t):
            with torch.no_grad():
                return cpu_unet(x, torch.tensor(t))["sample"]

        xf_c, _, x0_c = loops.guided_edit_loop(s, eps_cpu, xt, eta=0.0, zs=None,
                                               guidance=loops.color_guidance([0.8, None, None], [1, 1, 1], loss_scale, 0, T))
        out["control: fp32 oracle on CPU vs fp32 oracle on GPU"] = {
            "final_max_abs": (xf_c - xf).abs().max().item(), "final_rms": (xf_c - xf).pow(2).mean().sqrt().item(),
            "x0_traj_max_abs": [(a - b).abs().max().item() for a, b in zip(x0_c, x0_h)],
            "final_abs_max_of_ref": xf.abs().max().item()}
    for prec in precisions:
        w = create_diffusion_model("ddpm", sample_clipping=True, max_batch=1, state_dict=sd, unet_config=cfg, precision=prec)
        w.scheduler.set_timesteps(T)
        f = SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=loss_scale, t1=0, t2=T)
        res = SegDiffEditPipeline(w, None).edit_image(xt=xt.cuda(), attr_func=f, prog_bar=False, output_type="tensor")
        traj = [(a.cpu() - b).abs().max().item() for a, b in zip(res.pred_original_samples, x0_h)]
        out[prec] = {"final_max_abs": (res.imgs.cpu() - xf).abs().max().item(),
                     "final_rms": (res.imgs.cpu() - xf).pow(2).mean().sqrt().item(),
                     "x0_traj_max_abs": traj, "final_abs_max_of_ref": xf.abs().max().item()}
        del w
    return out


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    r = run(T)
    for k, v in r.items():
        print(k, "final max-abs %.3e rms %.3e" % (v["final_max_abs"], v["final_rms"]),
              "| x0 trajectory max-abs:", " ".join("%.1e" % e for e in v["x0_traj_max_abs"][::max(1, T // 10)]))
    print(json.dumps(r))

"""The config-2 harness of the reference arm (oracle/ref_harness.py: the reference's OWN diffusion_loop / get_noise_pred /
get_variance_noise / reverse_step / AttrFunc.apply from baseline/_ref/src, composed in the order of
src/SegDiffEditPipeline.py:248-296) must agree BIT-EXACTLY with the oracle loop the GPU parity tests check the CUDA
path against (oracle.loops.guided_edit_loop, mode="ddpm") - this pins the timed path's checker to the unmodified
reference.  Runs in a subprocess: the reference's bare module names collide with the drop-in package's."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(repo)r)
from types import SimpleNamespace
import torch
from oracle import ref_harness as rh
ref = rh.load_reference()
from oracle import loops
from oracle.ddim_scheduler import DDIMScheduler
from oracle.unet2d import ToyEpsModel
T, Tskip, S = 50, 36, 16
for B in (1, 3):
    sch = DDIMScheduler.from_preset("ddpm")
    sch.config.clip_sample = False
    sch.set_timesteps(T)
    unet = ToyEpsModel(3, S, seed=5)
    w = ref.DDPM(SimpleNamespace(unet=unet, scheduler=sch, device=torch.device("cpu")))
    g = torch.Generator().manual_seed(11 + B)
    xt = torch.randn(B, 3, S, S, generator=g)
    zs = 3.0 * torch.randn(T, 3, S, S, generator=g)
    n = T - Tskip
    x_ref, _ = rh.config2_regeneration(ref, w, xt, zs[Tskip:], n, eta=1.0, loss_scale=50.0)

    def eps_fn(x, t):
        with torch.no_grad():
            return unet(x, torch.tensor(t))["sample"]

    x_or, _, _ = loops.guided_edit_loop(sch, eps_fn, xt, eta=1.0, zs=zs[Tskip:], mode="ddpm",
                                        guidance=loops.color_guidance([0.8, None, None], [1, 1, 1], 50.0, 0, 10 ** 9))
    assert torch.equal(x_ref.detach(), x_or), (B, (x_ref.detach() - x_or).abs().max().item())
print("HARNESS_OK")
"""


@pytest.mark.skipif(not os.path.isfile(os.path.join(REPO, "baseline", "_ref", "src", "SegDiffEditPipeline.py")),
                    reason="baseline/_ref/src not installed (__graft_entry__.build() copies it where /root/reference exists)")
def test_config2_harness_of_reference_functions_equals_oracle_loop():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"repo": REPO}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HARNESS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

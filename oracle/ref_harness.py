"""Runs the UNMODIFIED reference (a copy of /root/reference/src under baseline/_ref/src, made by
__graft_entry__.build(); git-ignored, shipped to the GPU box by gpurun) on the host CPU.

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py): used by ``bench.py --impl reference`` and bench.py's
``cpu_baseline`` leg, never by the product path.

The reference's own modules are imported as they are, with ``sys.path = [oracle/shims, baseline/_ref/src]``: the only
stand-ins are the third-party ``diffusers`` / ``lpips`` packages (absent offline; ``oracle/shims`` re-exports the oracle
restatements of DDIMScheduler / UNet2DModel) and the one call-time patch of tests/golden/make_golden.py (the reference
hard-codes ``.to("cuda")``, src/utils.py:74 - mapped to the CPU).

Workloads:
  * ``config1_edit_image``: BASELINE configs[0] through the reference's own public call,
    ``SegDiffEditPipeline.edit_image`` (src/SegDiffEditPipeline.py:202-302): DDPM-256 UNet2DModel, colour-guided DDIM.
  * ``config2_regeneration``: BASELINE configs[1].  The reference's own ``edit_image(inversion_method="ddpm", Tskip=...)``
    branch raises (``pred_original_sample`` is never bound, src/SegDiffEditPipeline.py:260-268 vs :298), so - as
    SURVEY.md section 8(d) specifies - the loop body is composed here from the reference's OWN functions, verbatim and
    in the order of src/SegDiffEditPipeline.py:248-296: diffusion_loop -> get_noise_pred -> get_variance_noise ->
    reverse_step -> AttrFunc.apply."""
from __future__ import annotations

import os
import sys
import time
from types import SimpleNamespace

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(REPO, "baseline", "_ref", "src")
_BARE = ("attr_functions", "attr_functions_registry", "base_diffusion", "ddim_inversion", "ddpm_inversion", "diffusion_classes",
         "diffusion_utils", "utils", "transforms", "constants", "mask_creator", "Morphology", "SegDiffEditPipeline", "models")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "SegDiffEditPipeline.py"))


def load_reference():
    """Imports the reference's modules from baseline/_ref/src.  The drop-in package uses the same bare module names, so
    its directory must not shadow them: it is removed from sys.path and already-imported same-named modules are refused."""
    if not available():
        raise FileNotFoundError(f"{REF_SRC} not found: run __graft_entry__.build() where /root/reference exists")
    pkg = os.path.join(REPO, "diffusion-image-editing_b200")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    for name in _BARE:
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "")).startswith(REF_SRC):
            raise RuntimeError(f"module {name!r} of the drop-in package is already imported; run the reference arm in its own process")
    for p in (REF_SRC, os.path.join(REPO, "oracle", "shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if REPO not in sys.path:
        sys.path.append(REPO)
    if not getattr(torch.Tensor.to, "_b2e_cpu_patch", False):
        orig = torch.Tensor.to

        def to(self, *a, **k):   # the reference hard-codes "cuda" (src/utils.py:74, src/attr_functions.py:61)
            a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
            return orig(self, *a, **k)

        to._b2e_cpu_patch = True
        torch.Tensor.to = to
    import attr_functions
    import ddpm_inversion
    import diffusion_utils
    from diffusion_classes import DDPM
    from SegDiffEditPipeline import SegDiffEditPipeline
    return SimpleNamespace(attr_functions=attr_functions, ddpm_inversion=ddpm_inversion, diffusion_utils=diffusion_utils,
                           DDPM=DDPM, SegDiffEditPipeline=SegDiffEditPipeline)


def build_ddpm256(ref, clip_sample: bool, num_inference_steps: int = 50, seed: int = 0):
    """The pipeline object the reference's model factory returns for google/ddpm-celebahq-256 (src/models.py:17-32), with
    random-init weights (no hub access): diffusers' UNet2DModel / DDIMScheduler are the oracle restatements."""
    from oracle.ddim_scheduler import DDIMScheduler
    from oracle.unet2d import DDPM256_CONFIG, UNet2DModel
    torch.manual_seed(seed)
    unet = UNet2DModel(**DDPM256_CONFIG).eval()
    sch = DDIMScheduler.from_preset("ddpm")
    sch.config.clip_sample = clip_sample      # src/models.py:28
    sch.set_timesteps(num_inference_steps)
    return ref.DDPM(SimpleNamespace(unet=unet, scheduler=sch, device=torch.device("cpu")))


def config1_edit_image(ref, wrapper, xt, n_steps: int, loss_scale: float = 100.0):
    """``SegDiffEditPipeline.edit_image`` as the reference ships it (DDIM eta = 0, SingleColorAttrFunc), n_steps denoising
    steps (scheduler.set_timesteps(n_steps)).  Returns seconds."""
    wrapper.model.scheduler.set_timesteps(n_steps)
    pipe = ref.SegDiffEditPipeline(wrapper, None)
    f = ref.attr_functions.SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=loss_scale, t1=0, t2=n_steps)
    t0 = time.perf_counter()
    pipe.edit_image(xt=xt, eta=0, attr_func=f)
    return time.perf_counter() - t0


def config2_regeneration(ref, wrapper, xt, zs, n_steps: int, eta: float = 1.0, target: float = 0.8, color_idx: int = 0,
                         loss_scale: float = 50.0):
    """n_steps guided regeneration steps of the edit-friendly DDPM inversion (the last n_steps scheduler timesteps, the
    window ``zs`` selects), composed from the reference's own functions in the order of src/SegDiffEditPipeline.py:248-296.
    xt (B,3,256,256), zs (n_steps,3,256,256).  Returns (x_final, seconds)."""
    du, inv = ref.diffusion_utils, ref.ddpm_inversion
    f = ref.attr_functions.SingleColorAttrFunc(target=target, color_idx=color_idx, loss_scale=loss_scale, t1=0, t2=10 ** 9)
    f.kwargs["mask"] = None                                  # src/SegDiffEditPipeline.py:281-284
    model = wrapper.model
    t0 = time.perf_counter()
    for step_idx, timestep in du.diffusion_loop(model, zs[:n_steps], prog_bar=False):
        with torch.no_grad():
            noise_pred = du.get_noise_pred(model, xt, timestep, None, None)
        variance_noise = du.get_variance_noise(zs, step_idx, eta)
        xt = inv.reverse_step(model=model, model_output=noise_pred, timestep=timestep, sample=xt, eta=eta,
                              variance_noise=variance_noise)
        xt, variance_noise = f.apply(xt=xt, zt=variance_noise, model_output=noise_pred, timestep=timestep, step_idx=step_idx,
                                     model=wrapper, **f.kwargs)
    return xt, time.perf_counter() - t0

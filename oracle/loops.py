"""Oracle restatement of the loop compositions of the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against the unmodified reference
(tests/golden/pipeline.npz, inversion.npz).  Each loop takes ``eps_fn(x, t) -> eps`` so a
test can plug either a model or recorded (teacher-forced) noise predictions.
"""
from __future__ import annotations

import torch

from . import step_math as sm


def window_timesteps(scheduler, n_ops=None):
    """``diffusion_loop`` (src/diffusion_utils.py:112-133): the last n_ops timesteps with
    step indices restarting at 0."""
    ts = [int(t) for t in scheduler.timesteps]
    if n_ops is not None:
        ts = ts[-n_ops:]
    return list(enumerate(ts))


def guided_edit_loop(scheduler, eps_fn, xt, eta=0.0, zs=None, guidance=None, mode="ddim"):
    """The body of ``SegDiffEditPipeline.edit_image`` (src/SegDiffEditPipeline.py:248-298).

    guidance: None or callable (x_post, eps, coeffs, step_idx) -> x_post'.
    mode "ddim"  -> scheduler.step (clipping per scheduler.config.clip_sample)
    mode "ddpm"  -> reverse_step (the Tskip branch; x0 is computed here because the
                    reference leaves it unbound, SURVEY section 8a defects).
    Returns (x_final, [eps], [x0_pred])."""
    eps_hist, x0_hist = [], []
    n_ops = zs.shape[0] if zs is not None else None
    for step_idx, t in window_timesteps(scheduler, n_ops):
        eps = eps_fn(xt, t)
        c = sm.step_coeffs(scheduler, t)
        z = zs[step_idx] if (zs is not None and eta != 0) else None
        if mode == "ddpm":
            xt, x0 = sm.ddpm_reverse_step(xt, eps, c, eta, z)
        else:
            xt, x0 = sm.ddim_step(xt, eps, c, eta, z, clip=scheduler.config.clip_sample,
                                  clip_range=scheduler.config.clip_sample_range)
        if guidance is not None:
            xt = guidance(xt, eps, c, step_idx)
        eps_hist.append(eps)
        x0_hist.append(x0)
    return xt, eps_hist, x0_hist


def color_guidance(targets, weights, loss_scale, t1=0, t2=50, mask=None, mask_grad=False):
    def fn(x_post, eps, c, step_idx):
        if step_idx < t1 or step_idx >= t2:
            return x_post
        return sm.color_guidance_update(x_post, eps, c, targets, weights, loss_scale,
                                        mask=mask, mask_grad=mask_grad)[0]
    return fn


def invert_ddpm(scheduler, eps_fn, x0, noises, eta=1.0):
    """``inversion_forward_process`` for eta > 0 (src/ddpm_inversion.py:80-176).
    Returns (x_T, zs, xts)."""
    T = len(scheduler.timesteps)
    etas = [eta] * T if isinstance(eta, (int, float)) else list(eta)
    xts = sm.sample_xts(x0, scheduler, noises)
    zs = torch.zeros((T,) + tuple(x0.shape[1:]), dtype=x0.dtype)
    xt = x0
    for idx in reversed(range(T)):
        t = int(scheduler.timesteps[idx])
        xt = xts[idx][None]
        eps = eps_fn(xt, t)
        c = sm.step_coeffs(scheduler, t)
        z, xm = sm.extract_noise(xt, xts[idx + 1][None], eps, c, etas[idx])
        zs[idx] = z[0]
        xts[idx + 1] = xm[0]
    zs[-1] = torch.zeros_like(zs[-1])
    return xt, zs, xts


def invert_eta0(scheduler, eps_fn, x0):
    """eta = 0 branch: repeated ``forward_step`` (src/ddpm_inversion.py:129-131)."""
    xt = x0
    for idx in reversed(range(len(scheduler.timesteps))):
        t = int(scheduler.timesteps[idx])
        xt = sm.ddpm_forward_step(xt, eps_fn(xt, t), scheduler, t)
    return xt


def sample_ddpm(scheduler, eps_fn, zs, xts, tskip=36, eta=1.0):
    """``sample`` -> ``inversion_reverse_process`` (src/ddpm_inversion.py:243-313)."""
    xt = xts[tskip][None]
    z_used = zs[tskip:]
    for idx, t in window_timesteps(scheduler, z_used.shape[0]):
        eps = eps_fn(xt, t)
        c = sm.step_coeffs(scheduler, t)
        z = z_used[idx] if eta != 0 else None
        xt, _ = sm.ddpm_reverse_step(xt, eps, c, eta, z)
    return xt


def invert_ddim(scheduler, eps_fn, x0):
    """``ddim_inversion`` (src/ddim_inversion.py:52-75)."""
    x = x0.clone()
    T = scheduler.num_inference_steps
    for i in range(T):
        t = int(scheduler.timesteps[T - i - 1])
        x = sm.ddim_next_step(x, eps_fn(x, t), scheduler, t)
    return x

"""GPU parity of the fused step kernels: CUDA (through the C ABI) vs the oracle on the same
seeded inputs (bit-exact, fp32) and vs the reference's golden vectors."""
import re

import numpy as np
import pytest
import torch

from oracle import step_math as sm
from oracle.ddim_scheduler import DDIMScheduler

pytestmark = pytest.mark.gpu
T_ = torch.from_numpy


def dev(x):
    return x.cuda() if x is not None else None


def sched(preset, T, clip=None):
    s = DDIMScheduler.from_preset(preset)
    if clip is not None:
        s.config.clip_sample = clip
    s.set_timesteps(T)
    return s


def coeffs(s, t, eta, mode):
    from b200edit import ops
    T = s.num_inference_steps
    return ops.step_coeffs(s.alphas_cumprod, s.final_alpha_cumprod, t, t - 1000 // T, eta, mode)


def close_to_golden(a, g):
    # goldens were generated on the dev container's CPU; scalar pow() may differ by an ulp on
    # another host, so allow a few ulps here (bit-exactness is asserted against the in-process oracle)
    return np.allclose(a, g, rtol=3e-6, atol=1e-6, equal_nan=True)


def test_single_and_reverse_step_vs_golden_and_oracle(golden):
    from b200edit import ops
    g = golden("step_math")
    x, e, z = T_(g["x"]), T_(g["e"]), T_(g["z"])
    xd, ed, zd = dev(x), dev(e), dev(z)
    exact = 0
    for key in g["cases"]:
        key = str(key)
        m = re.match(r"(\w+)_T(\d+)_t(\d+)_clip(\d)_eta([\d.]+)", key)
        preset, T, t, clip, eta = m.group(1), int(m.group(2)), int(m.group(3)), bool(int(m.group(4))), float(m.group(5))
        s = sched(preset, T)
        xp, x0 = ops.guided_step(xd, ed, coeffs(s, t, eta, "ddim"), clip=clip, noise=zd if eta > 0 else None)
        oc = sm.step_coeffs(s, t)
        oxp, ox0 = sm.ddim_step(x, e, oc, eta, z if eta > 0 else None, clip=clip)
        assert np.array_equal(xp.cpu().numpy(), oxp.numpy(), equal_nan=True), key
        assert np.array_equal(x0.cpu().numpy(), ox0.numpy(), equal_nan=True), key
        assert close_to_golden(xp.cpu().numpy(), g[f"single_step/{key}/prev"]), key
        assert close_to_golden(x0.cpu().numpy(), g[f"single_step/{key}/x0"]), key
        exact += np.array_equal(xp.cpu().numpy(), g[f"single_step/{key}/prev"], equal_nan=True)
        if not clip:
            xr, _ = ops.guided_step(xd, ed, coeffs(s, t, eta, "ddpm"), noise=zd if eta > 0 else None)
            oxr, _ = sm.ddpm_reverse_step(x, e, oc, eta, z if eta > 0 else None)
            assert np.array_equal(xr.cpu().numpy(), oxr.numpy(), equal_nan=True), key
            assert close_to_golden(xr.cpu().numpy(), g[f"reverse_step/{key}"]), key
    print(f"bit-exact vs golden: {exact}/{len(g['cases'])}")


def test_pointwise_ops(golden):
    from b200edit import ops
    g = golden("step_math")
    x, e = T_(g["x"]), T_(g["e"])
    xd, ed = dev(x), dev(e)
    for k in [k for k in g.files if k.startswith("pred_x0/")]:
        m = re.match(r"pred_x0/(\w+)_T(\d+)_t(\d+)", k)
        preset, T, t = m.group(1), int(m.group(2)), int(m.group(3))
        s = sched(preset, T)
        c = sm.step_coeffs(s, t)
        tag = f"{preset}_T{T}_t{t}"
        got = ops.pred_x0(xd, ed, float(c.sqrt_a_t), float(c.sqrt_b_t)).cpu().numpy()
        assert np.array_equal(got, sm.pred_x0(x, e, c).numpy())
        assert close_to_golden(got, g[k])
        ac = s.alphas_cumprod
        stride = 1000 // T
        t_cur = min(t - stride, 999)
        a_cur = ac[t_cur] if t_cur >= 0 else s.final_alpha_cumprod
        got = ops.renoise(xd, ed, float(a_cur ** 0.5), float((1 - a_cur) ** 0.5), float(ac[t] ** 0.5),
                          float((1 - ac[t]) ** 0.5)).cpu().numpy()
        assert np.array_equal(got, sm.ddim_next_step(x, e, s, t).numpy())
        assert close_to_golden(got, g["next_step/" + tag])
        t_n = min(998, t + stride)
        got = ops.renoise(xd, ed, float(ac[t] ** 0.5), float((1 - ac[t]) ** 0.5), float(ac[t_n] ** 0.5),
                          float((1 - ac[t_n]) ** 0.5)).cpu().numpy()
        assert np.array_equal(got, sm.ddpm_forward_step(x, e, s, t).numpy())
        assert close_to_golden(got, g["forward_step/" + tag])
    got = ops.apply_mask(dev(T_(g["am_mask"])), dev(T_(g["am_zo"])), dev(T_(g["am_zv"]))).cpu().numpy()
    assert np.array_equal(got, g["apply_mask"])
    both = dev(T_(g["cfg_both"]))
    for s_ in (3.5, 7.5):
        assert np.array_equal(ops.cfg_combine(both[:2], both[2:], s_).cpu().numpy(), g[f"cfg/{s_}"])
    u8 = ops.to_uint8(dev(T_(g["pil_in"])))[0].cpu().numpy()
    assert np.array_equal(u8, g["pil_out"])


def test_colour_guidance(golden):
    from b200edit import ops
    g = golden("guidance")
    e, xpost, mask = T_(g["e"]), T_(g["xpost"]), T_(g["mask"])
    s = sched("ddpm", 50)
    # The fused kernel performs scheduler step + guidance.  To test the guidance alone against
    # AttrFunc.apply, feed it an identity step: x_prev == x_post requires inverting the update, so
    # instead we compare the full fused step with the oracle's step-then-guide composition, and the
    # guidance alone through the generic gradient path below.
    x_t = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(5))
    for t in (980, 500, 20, 0):
        oc = sm.step_coeffs(s, t)
        c = coeffs(s, t, 0.0, "ddim")
        for scale in (100.0, 37.5):
            for targets, weights, mk, mg in (([0.8, None, None], None, None, False),
                                             ([None, None, -0.25], None, mask, True),
                                             ([0.9, -0.3, 0.45], [0.9, -0.3, 0.45], None, False)):
                for clip in (False, True):
                    xp, x0 = ops.guided_step(dev(x_t), dev(e), c, clip=clip, targets=targets, weights=weights,
                                             loss_scale=scale, mask=dev(mk), mask_grad=mg)
                    oxp, ox0 = sm.ddim_step(x_t, e, oc, 0.0, None, clip=clip)
                    oxg, _ = sm.color_guidance_update(oxp, e, oc, targets, weights or [1, 1, 1], scale,
                                                      mask=mk, mask_grad=mg)
                    assert np.array_equal(xp.cpu().numpy(), oxg.numpy())
                    assert np.array_equal(x0.cpu().numpy(), ox0.numpy())
    # guidance alone (reference golden): x_post + g * a_t^2 with g from the closed form
    for t in (980, 500, 20, 0):
        oc = sm.step_coeffs(s, t)
        _, gvec = sm.color_guidance_update(xpost, e, oc, [0.8, None, None], [1, 1, 1], 100.0)
        got = ops.apply_guidance_grad(dev(xpost), dev(gvec), float(oc.a_t_sq)).cpu().numpy()
        assert close_to_golden(got, g[f"single/t{t}_s100.0"])


def test_l2_regularised_guidance(golden):
    from b200edit import ops
    g = golden("guidance")
    e, mask, x_ref = T_(g["e"]), T_(g["mask"]), T_(g["x_ref"])
    s = sched("ddpm", 50)
    x_t = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(6))
    for t in (980, 500, 20):
        oc = sm.step_coeffs(s, t)
        c = coeffs(s, t, 0.0, "ddim")
        for mg in (False, True):
            xp, x0 = ops.guided_step_l2reg(dev(x_t), dev(e), c, x_ref=dev(x_ref), mask=dev(mask), lambda_=0.1,
                                           targets=[None, 0.8, None], loss_scale=37.5, mask_grad=mg)
            oxp, ox0 = sm.ddim_step(x_t, e, oc, 0.0, None)

            def loss(x0g):
                return sm.l2reg_loss(x0g, mask, x_ref, 0.1, lambda im: sm.single_color_loss(im, 1, 0.8))
            oxg, _ = sm.autograd_guidance_update(oxp, e, oc, loss, 37.5, mask=mask, mask_grad=mg)
            upd_ref = (oxg - oxp).numpy()
            upd = xp.cpu().numpy() - oxp.numpy()
            # tolerance: the norm is a global fp32 reduction (order differs); 1e-5 relative to the update
            assert np.allclose(upd, upd_ref, rtol=1e-4, atol=1e-5 * np.abs(upd_ref).max())
            assert np.array_equal(x0.cpu().numpy(), ox0.numpy())


def test_inversion_kernels(golden):
    from b200edit import ops
    g = golden("inversion")
    T = int(g["T"])
    for preset in ("ddpm", "sd"):
        s = sched(preset, T, clip=False)
        x0 = T_(g[f"{preset}/x0"])
        noises = T_(g[f"{preset}/fwd_noises"])
        ac = s.alphas_cumprod
        ts = s.timesteps
        sa = (ac[ts] ** 0.5)
        sb = ((1 - ac) ** 0.5)[ts]
        xts = ops.sample_xts(dev(x0), dev(noises), dev(sa), dev(sb))
        assert np.array_equal(xts.cpu().numpy(), sm.sample_xts(x0, s, noises).numpy())
        assert close_to_golden(xts.cpu().numpy(), g[f"{preset}/xts_sampled"])
        # teacher-forced z_t extraction over the whole trajectory
        for eta in (1.0, 0.6):
            p = f"{preset}/eta{eta}/"
            eps_rec = T_(g[p + "inv_eps"])
            xts_d = dev(T_(g[f"{preset}/xts_sampled"]).clone())
            zs_d = torch.zeros((T, 3, 16, 16), device="cuda")
            xts_o = T_(g[f"{preset}/xts_sampled"]).clone()
            zs_o = torch.zeros((T, 3, 16, 16))
            for i, idx in enumerate(reversed(range(T))):
                t = int(ts[idx])
                eps = eps_rec[i][None]
                c = coeffs(s, t, eta, "ddpm")
                ops.extract_noise(xts_d[idx][None], dev(eps), xts_d[idx + 1], zs_d[idx], c)
                z, xm = sm.extract_noise(xts_o[idx][None], xts_o[idx + 1][None], eps, sm.step_coeffs(s, t), eta)
                zs_o[idx] = z[0]
                xts_o[idx + 1] = xm[0]
            zs_d[-1].zero_()
            zs_o[-1] = 0
            assert np.array_equal(zs_d.cpu().numpy(), zs_o.numpy(), equal_nan=True)
            assert np.array_equal(xts_d.cpu().numpy(), xts_o.numpy(), equal_nan=True)
            assert close_to_golden(zs_d.cpu().numpy(), g[p + "zs"])


@pytest.mark.parametrize("shape", [(1, 3, 256, 256), (3, 4, 64, 64), (2, 3, 5, 7), (8, 3, 256, 256)])
def test_fused_step_shapes_and_broadcast(shape):
    """Full-size and ragged shapes (scalar fallback when HW % 4 != 0), batched / broadcast noise & mask."""
    from b200edit import ops
    B, C, H, W = shape
    gen = torch.Generator().manual_seed(B * 1000 + H)
    x, e = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    z1, zb = torch.randn(shape[1:], generator=gen), torch.randn(shape, generator=gen)
    m1 = (torch.randn((1,) + shape[1:], generator=gen) > 0).float()
    s = sched("ddpm", 50)
    oc = sm.step_coeffs(s, 500)
    targets = [0.5, None, -0.5, None][:C]
    for z in (z1, zb):
        c = coeffs(s, 500, 0.8, "ddpm")
        xp, x0 = ops.guided_step(dev(x), dev(e), c, noise=dev(z), targets=targets, loss_scale=50.0,
                                 mask=dev(m1), mask_grad=True)
        oxp, ox0 = sm.ddpm_reverse_step(x, e, oc, 0.8, z)
        oxg, _ = sm.color_guidance_update(oxp, e, oc, targets, [1] * C, 50.0, mask=m1, mask_grad=True)
        assert np.array_equal(xp.cpu().numpy(), oxg.numpy())
        assert np.array_equal(x0.cpu().numpy(), ox0.numpy())

"""The oracle restatement vs. vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Bit-exact unless a tolerance is stated."""
import re

import numpy as np
import pytest
import torch

from oracle import loops, mask as omask, step_math as sm
from oracle.ddim_scheduler import DDIMScheduler

T_ = torch.from_numpy


def sched(preset, T, clip=None):
    s = DDIMScheduler.from_preset(preset)
    if clip is not None:
        s.config.clip_sample = clip
    s.set_timesteps(T)
    return s


def parse_case(key):
    m = re.match(r"(\w+)_T(\d+)_t(\d+)_clip(\d)_eta([\d.]+)", key)
    return m.group(1), int(m.group(2)), int(m.group(3)), bool(int(m.group(4))), float(m.group(5))


def test_single_and_reverse_step(golden):
    g = golden("step_math")
    x, e, z = T_(g["x"]), T_(g["e"]), T_(g["z"])
    n = 0
    for key in g["cases"]:
        preset, T, t, clip, eta = parse_case(str(key))
        c = sm.step_coeffs(sched(preset, T), t)
        xp, x0 = sm.ddim_step(x, e, c, eta, z if eta > 0 else None, clip=clip)
        assert np.array_equal(xp.numpy(), g[f"single_step/{key}/prev"]), key
        assert np.array_equal(x0.numpy(), g[f"single_step/{key}/x0"]), key
        if not clip:
            xr, _ = sm.ddpm_reverse_step(x, e, c, eta, z if eta > 0 else None)
            assert np.array_equal(xr.numpy(), g[f"reverse_step/{key}"]), key
        n += 1
    assert n == 180


def test_pointwise_ops(golden):
    g = golden("step_math")
    x, e = T_(g["x"]), T_(g["e"])
    keys = [k for k in g.files if k.startswith("pred_x0/")]
    assert len(keys) == 30
    for k in keys:
        m = re.match(r"pred_x0/(\w+)_T(\d+)_t(\d+)", k)
        preset, T, t = m.group(1), int(m.group(2)), int(m.group(3))
        s = sched(preset, T)
        c = sm.step_coeffs(s, t)
        tag = f"{preset}_T{T}_t{t}"
        assert np.array_equal(sm.pred_x0(x, e, c).numpy(), g[k])
        assert np.array_equal(sm.ddim_next_step(x, e, s, t).numpy(), g["next_step/" + tag])
        assert np.array_equal(sm.ddpm_forward_step(x, e, s, t).numpy(), g["forward_step/" + tag])
        assert np.array_equal(c.variance.numpy(), g["variance/" + tag])
    assert np.array_equal(sm.apply_mask(T_(g["am_mask"]), T_(g["am_zo"]), T_(g["am_zv"])).numpy(),
                          g["apply_mask"])
    both = T_(g["cfg_both"])
    for s_ in (3.5, 7.5):
        assert np.array_equal(sm.cfg_combine(both[:2], both[2:], s_).numpy(), g[f"cfg/{s_}"])
    u8 = sm.to_uint8_image(T_(g["pil_in"]))[0].permute(1, 2, 0).numpy()
    assert np.array_equal(u8, g["pil_out"])


def test_colour_guidance_closed_form(golden):
    g = golden("guidance")
    e, xpost, mask = T_(g["e"]), T_(g["xpost"]), T_(g["mask"])
    s = sched("ddpm", 50)
    for t in (980, 500, 20, 0):
        c = sm.step_coeffs(s, t)
        for scale in (100.0, 37.5):
            out, _ = sm.color_guidance_update(xpost, e, c, [0.8, None, None], [1, 1, 1], scale)
            assert np.array_equal(out.numpy(), g[f"single/t{t}_s{scale}"])
            out, _ = sm.color_guidance_update(xpost, e, c, [None, None, -0.25], [1, 1, 1], scale,
                                              mask=mask, mask_grad=True)
            assert np.array_equal(out.numpy(), g[f"single_maskgrad/t{t}_s{scale}"])
            out, _ = sm.color_guidance_update(xpost, e, c, [0.9, -0.3, 0.45], [0.9, -0.3, 0.45], scale)
            assert np.array_equal(out.numpy(), g[f"multi/t{t}_s{scale}"])
    assert np.array_equal(g["window_outside"], g["xpost"])


def test_l2_regularised_guidance(golden):
    """Autograd restatement of the masked + L2-regularised loss; the norm is a global
    reduction, so allow 1 ulp-level differences: rtol 1e-6 on the update."""
    g = golden("guidance")
    e, xpost, mask, x_ref = T_(g["e"]), T_(g["xpost"]), T_(g["mask"]), T_(g["x_ref"])
    s = sched("ddpm", 50)
    for t in (980, 500, 20, 0):
        c = sm.step_coeffs(s, t)
        for scale in (100.0, 37.5):
            def loss(x0g):
                return sm.l2reg_loss(x0g, mask, x_ref, 0.1, lambda im: sm.single_color_loss(im, 1, 0.8))
            for mg, tag in ((False, "single_l2reg"), (True, "single_l2reg_maskgrad")):
                out, _ = sm.autograd_guidance_update(xpost, e, c, loss, scale, mask=mask, mask_grad=mg)
                ref = g[f"{tag}/t{t}_s{scale}"]
                upd_ref = ref - g["xpost"]
                assert np.allclose(out.numpy() - g["xpost"], upd_ref, rtol=1e-5,
                                   atol=1e-7 * np.abs(upd_ref).max())


def test_loss_heads(golden):
    g = golden("guidance")
    logits = T_(g["seg_logits"]).clone().requires_grad_(True)
    loss = sm.segmentation_area_loss(logits, list(g["seg_classes"]))
    loss.backward()
    assert np.allclose(loss.item(), g["seg_loss"], rtol=1e-6)
    assert np.allclose(logits.grad.numpy(), g["seg_dlogits"], rtol=1e-5, atol=1e-10)
    lg = T_(g["cls_logits"]).clone().requires_grad_(True)
    loss = sm.classifier_logit_loss(lg, 31, 1)
    loss.backward()
    assert loss.item() == g["cls_loss"]
    assert np.array_equal(lg.grad.numpy(), g["cls_dlogits"])
    lg = T_(g["cls_logits"]).clone().requires_grad_(True)
    loss = sm.classifier_logit_loss(lg, 31, 0, (15, 1, torch.tensor([0.3, -0.6])))
    loss.backward()
    assert np.allclose(loss.item(), g["cls_reg_loss"], rtol=1e-6)
    assert np.allclose(lg.grad.numpy(), g["cls_reg_dlogits"], rtol=1e-6)


class Replay:
    """eps_fn replaying recorded noise predictions in call order."""

    def __init__(self, eps):
        self.eps, self.i = T_(eps), 0

    def __call__(self, x, t):
        self.i += 1
        return self.eps[self.i - 1][None]


@pytest.mark.parametrize("preset", ["ddpm", "sd"])
def test_inversion_loops(golden, preset):
    g = golden("inversion")
    T = int(g["T"])
    s = sched(preset, T, clip=False)
    x0 = T_(g[f"{preset}/x0"])
    noises = T_(g[f"{preset}/fwd_noises"])
    xts = sm.sample_xts(x0, s, noises)
    assert np.array_equal(xts.numpy(), g[f"{preset}/xts_sampled"])
    for eta in (1.0, 0.6):
        p = f"{preset}/eta{eta}/"
        xT, zs, xts2 = loops.invert_ddpm(s, Replay(g[p + "inv_eps"]), x0, noises, eta)
        assert np.array_equal(xT.numpy(), g[p + "xT"])
        assert np.array_equal(zs.numpy(), g[p + "zs"], equal_nan=True)
        assert np.array_equal(xts2.numpy(), g[p + "xts"], equal_nan=True)
        for tskip in (0, 7):
            xr = loops.sample_ddpm(s, Replay(g[p + f"sample_T{tskip}_eps"]), zs, xts2, tskip, eta)
            assert np.array_equal(xr.numpy(), g[p + f"sample_T{tskip}"], equal_nan=True)
    assert np.array_equal(loops.invert_eta0(s, Replay(g[f"{preset}/eta0/eps"]), x0).numpy(),
                          g[f"{preset}/eta0/xT"])
    assert np.array_equal(loops.invert_ddim(s, Replay(g[f"{preset}/ddim_inv/eps"]), x0).numpy(),
                          g[f"{preset}/ddim_inv/xT"])


def to_img(x):
    return sm.to_uint8_image(x)[0].permute(1, 2, 0).numpy()


def test_edit_pipeline_loops(golden):
    g = golden("pipeline")
    T = int(g["T"])
    s = sched("ddpm", T, clip=True)
    xt, zs, mask = T_(g["xt"]), T_(g["zs"]), T_(g["mask"])
    cases = {
        "color_eta0": dict(eta=0.0, zs=None,
                           guidance=loops.color_guidance([0.8, None, None], [1, 1, 1], 100.0, 0, T)),
        "color_eta1_window": dict(eta=1.0, zs=zs,
                                  guidance=loops.color_guidance([0.8, None, None], [1, 1, 1], 100.0, 3, 15)),
        "color_maskgrad": dict(eta=0.0, zs=None,
                               guidance=loops.color_guidance([None, -0.5, None], [1, 1, 1], 60.0, 0, 50,
                                                             mask=mask, mask_grad=True)),
        "gen_eta0": dict(eta=0.0, zs=None, guidance=None),
        "gen_eta08": dict(eta=0.8, zs=zs, guidance=None),
    }
    for tag, kw in cases.items():
        xf, eps_h, x0_h = loops.guided_edit_loop(s, Replay(g[f"{tag}/eps"]), xt, **kw)
        assert np.array_equal(to_img(xf), g[f"{tag}/img"]), tag
        assert np.array_equal(np.stack([to_img(x) for x in x0_h]), g[f"{tag}/x0_imgs"]), tag
    errs = list(g["check_inputs_errors"])
    assert errs[0] == "eta > 0 and zs is empty" and errs[1] == "eta == 0 and zs is not empty"
    assert errs[2].startswith("attr_func is None")


def test_mask_creator_fixtures(golden):
    g = golden("mask")
    n = 0
    for k in g["keys"]:
        k = str(k)
        name, cls, d, dil = k.split("/")
        classes = [int(c) for c in cls[1:].split("_")]
        d = int(d[1:])
        m = omask.create_mask(g[f"seg/{name}"].astype(np.int64), classes, dil == "dil1", (d, d))
        ref = np.unpackbits(g["mask/" + k])[: d * d].reshape(d, d).astype(np.float32)
        assert m.shape == (1, 3, d, d)
        assert np.array_equal(m[0, 0], ref) and np.array_equal(m[0, 2], ref), k
        n += 1
    assert n == 60
    # survey probe counts
    seg = g["seg/testimg"].astype(np.int64)
    assert int(omask.create_mask(seg, [17, 1], False, (256, 256))[0, 0].sum()) == 32478
    assert int(omask.create_mask(seg, [17, 1], True, (256, 256))[0, 0].sum()) == 34503
    assert int(omask.create_mask(seg, [17, 1], False, (64, 64))[0, 0].sum()) == 1753
    assert int(omask.create_mask(seg, [17, 1], True, (64, 64))[0, 0].sum()) == 1931


def test_resize_and_morphology(golden):
    g = golden("mask")
    x = g["resize_in"]
    for k in [f for f in g.files if f.startswith("resize/")]:
        oh, ow = map(int, k.split("/")[1].split("x"))
        assert np.array_equal(omask.resize_bilinear_aa(x, oh, ow), g[k]), k
    xi = g["morph_in"]
    for k in (3, 5, 7):
        w = g[f"morph/w{k}"]
        assert np.array_equal(omask.dilate_hard(xi, k, w), g[f"morph/dil{k}"])
        assert np.array_equal(omask.erode_hard(xi, k, w), g[f"morph/ero{k}"])


def test_bisenet_restatement_vs_reference_module():
    """oracle/bisenet.py reproduces the unmodified reference BiSeNet (same seeds -> same random-init weights, same
    state_dict names): logits to fp32 round-off of the host's convolution kernels, parsing map identical."""
    from oracle.bisenet import BiSeNet, seeded_weights
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "bisenet.npz"))
    net = seeded_weights(BiSeNet(19).eval(), 1234)
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        out = net(x)[0]
    assert np.allclose(out[:, :, ::4, ::4].numpy(), g["out_sub4"], rtol=1e-4, atol=1e-5)
    assert (out[0].argmax(0).numpy() != g["argmax"]).mean() < 1e-3

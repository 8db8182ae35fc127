#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):
    python tests/golden/make_golden.py

sys.path = [<repo>/oracle/shims, /root/reference/src]: the reference's own src/*.py
are imported as-is; the only stand-ins are the ``diffusers``/``lpips`` stub packages
(third-party, absent offline - see oracle/__init__.py) and tiny fake networks where a
function needs *a* network (the networks themselves are not under test).  The single
call-time patch is mapping the hard-coded ``.to("cuda")`` of src/utils.py:74 to CPU.

Outputs (np.savez_compressed, fp32 unless noted):
  step_math.npz   single ops of src/diffusion_utils.py, ddim_inversion.py, ddpm_inversion.py, utils.py
  guidance.npz    AttrFunc.apply for the colour strategies, masked / L2-regularised variants, loss heads
  inversion.npz   sample_xts_from_x0, inversion_forward_process (invert), sample, ddim_inversion
  pipeline.npz    SegDiffEditPipeline.edit_image and Diffusion.generate_image end to end
  mask.npz        MaskCreator / Dilation2d / Erosion2d on the reference's fixture parsing maps
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(REPO, "oracle", "shims"), os.path.join(REF, "src")]
sys.path.append(REPO)
os.chdir(os.path.join(REF, "src"))

# --- the one call-time patch: hard-coded "cuda" -> cpu (src/utils.py:74) -------------
_orig_to = torch.Tensor.to


def _to(self, *a, **k):
    a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
    return _orig_to(self, *a, **k)


torch.Tensor.to = _to

import attr_functions as ref_attr  # noqa: E402
import base_diffusion  # noqa: E402,F401
import ddim_inversion as ref_ddim  # noqa: E402
import ddpm_inversion as ref_ddpm  # noqa: E402
import diffusion_utils as ref_du  # noqa: E402
import utils as ref_utils  # noqa: E402
from diffusion_classes import DDPM  # noqa: E402
from mask_creator import MaskCreator  # noqa: E402
from Morphology import Dilation2d, Erosion2d  # noqa: E402
from SegDiffEditPipeline import SegDiffEditPipeline  # noqa: E402
from transforms import tensor_to_pil  # noqa: E402

from oracle.ddim_scheduler import DDIMScheduler  # noqa: E402
from oracle.unet2d import ToyEpsModel  # noqa: E402

C, S = 3, 16


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def make_model(preset="ddpm", T=50, size=S, clip=None, seed=0):
    sch = DDIMScheduler.from_preset(preset)
    if clip is not None:
        sch.config.clip_sample = clip
    sch.set_timesteps(T)
    unet = ToyEpsModel(C, size, seed=seed)
    return SimpleNamespace(unet=unet, scheduler=sch, device=torch.device("cpu"))


def np_(t):
    return t.detach().cpu().numpy()


def gen_step_math():
    out = {}
    S = 8  # many cases: keep each vector small
    x, e, z = randn(1, 2, C, S, S), randn(2, 2, C, S, S), randn(3, C, S, S)
    out.update(x=np_(x), e=np_(e), z=np_(z))
    cases = []
    for preset in ("ddpm", "ldm", "sd"):
        for T in (50, 20):
            m = make_model(preset, T)
            ts = [int(m.scheduler.timesteps[i]) for i in (0, 1, T // 2, T - 2, T - 1)]
            for t in ts:
                tt = torch.tensor(t)
                for clip in (False, True):
                    m.scheduler.config.clip_sample = clip
                    for eta in (0.0, 0.7, 1.0):
                        key = f"{preset}_T{T}_t{t}_clip{int(clip)}_eta{eta}"
                        xp, x0 = ref_du.single_step(m, e, tt, x, eta, z if eta > 0 else None)
                        out["single_step/" + key + "/prev"] = np_(xp)
                        out["single_step/" + key + "/x0"] = np_(x0)
                        if not clip:
                            xr = ref_ddpm.reverse_step(m, e, tt, x, eta=eta,
                                                       variance_noise=z if eta > 0 else None)
                            out["reverse_step/" + key] = np_(xr)
                        cases.append(key)
                a_t = m.scheduler.alphas_cumprod[tt]
                out[f"pred_x0/{preset}_T{T}_t{t}"] = np_(
                    ref_du.compute_predicted_original_sample(x, 1 - a_t, e, a_t))
                out[f"next_step/{preset}_T{T}_t{t}"] = np_(ref_ddim.next_step(m, e, t, x))
                out[f"forward_step/{preset}_T{T}_t{t}"] = np_(ref_ddpm.forward_step(m, e, t, x))
                out[f"variance/{preset}_T{T}_t{t}"] = np_(ref_du.calculate_variance(m, tt))
    out["cases"] = np.array(cases)
    mask = (randn(4, 1, C, S, S) > 0).float()
    zo, zv = randn(5, 4, C, S, S), randn(6, 4, C, S, S)
    out.update(am_mask=np_(mask), am_zo=np_(zo), am_zv=np_(zv),
               apply_mask=np_(ref_utils.apply_mask(mask, zo, zv)))

    # CFG combine through get_noise_pred with a unet that returns a stored 2B batch
    class FakeCfgUnet:
        def __init__(self, both):
            self.both = both

        def __call__(self, sample, timestep, encoder_hidden_states):
            return {"sample": self.both}

    both = randn(7, 4, C, S, S)
    m = SimpleNamespace(unet=FakeCfgUnet(both))
    for s in (3.5, 7.5):
        out[f"cfg/{s}"] = np_(ref_du.get_noise_pred(m, x, torch.tensor(1), text_emb=randn(8, 2, 4, 8),
                                                    cfg_scale=s))
    out["cfg_both"] = np_(both)
    # tensor_to_pil numerics
    img = randn(9, 1, 3, S, S) * 1.5
    out["pil_in"] = np_(img)
    out["pil_out"] = np.asarray(tensor_to_pil(img))
    np.savez_compressed(os.path.join(HERE, "step_math.npz"), **out)
    print("step_math.npz", len(out))


class FakeSegNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.w = torch.nn.Parameter(torch.randn(19, 3, 1, 1, generator=g))
        self.logits = None

    def forward(self, img):
        self.logits = torch.nn.functional.conv2d(img, self.w)
        self.logits.retain_grad()
        return (self.logits,)


class FakePredictor(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(12)
        self.w = torch.nn.Parameter(torch.randn(80, 3, generator=g))
        self.logits = None

    def forward(self, img):
        self.logits = img.mean(dim=(2, 3)) @ self.w.t()
        self.logits.retain_grad()
        return self.logits


def gen_guidance():
    out = {}
    m = make_model("ddpm", 50)
    wrapper = DDPM(m)
    e = randn(21, 2, C, S, S)
    xpost = randn(22, 2, C, S, S)
    mask = (randn(23, 1, C, S, S) > 0).float()
    x_ref = randn(24, 2, C, S, S).clamp(-1, 1)
    out.update(e=np_(e), xpost=np_(xpost), mask=np_(mask), x_ref=np_(x_ref))
    keys = []
    for t in (980, 500, 20, 0):
        tt = torch.tensor(t)
        for scale in (100.0, 37.5):
            f = ref_attr.SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=scale)
            xo, _ = f.apply(xpost, None, e, tt, 3, wrapper)
            k = f"single/t{t}_s{scale}"
            out[k] = np_(xo); keys.append(k)
            f = ref_attr.SingleColorAttrFunc(target=-0.25, color_idx=2, loss_scale=scale)
            xo, _ = f.apply(xpost, None, e, tt, 3, wrapper, use_mask=True, mask=mask, mask_attr_grad=True)
            k = f"single_maskgrad/t{t}_s{scale}"
            out[k] = np_(xo); keys.append(k)
            f = ref_attr.MultiColorAttrFunc(r_target=0.9, g_target=-0.3, b_target=0.45, loss_scale=scale)
            xo, _ = f.apply(xpost, None, e, tt, 3, wrapper)   # no kwargs: its loss() takes none
            k = f"multi/t{t}_s{scale}"
            out[k] = np_(xo); keys.append(k)
            f = ref_attr.SingleColorAttrFunc(target=0.8, color_idx=1, loss_scale=scale, use_l2=True)
            xo, _ = f.apply(xpost, None, e, tt, 3, wrapper, mask_pred_original_sample=True, use_l2=True,
                            lambda_=0.1, mask=mask, x_0=x_ref)
            k = f"single_l2reg/t{t}_s{scale}"
            out[k] = np_(xo); keys.append(k)
            xo, _ = f.apply(xpost, None, e, tt, 3, wrapper, mask_pred_original_sample=True, use_l2=True,
                            lambda_=0.1, mask=mask, x_0=x_ref, mask_attr_grad=True)
            k = f"single_l2reg_maskgrad/t{t}_s{scale}"
            out[k] = np_(xo); keys.append(k)
    # window test: outside [t1, t2) nothing happens
    f = ref_attr.SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=100.0, t1=5, t2=10)
    xo, _ = f.apply(xpost, None, e, torch.tensor(500), 3, wrapper)
    out["window_outside"] = np_(xo)
    out["keys"] = np.array(keys)
    # loss heads with fake networks (the heads are reference code; the networks are not)
    img = randn(25, 1, 3, 8, 8).requires_grad_(True)
    seg = SimpleNamespace(net=FakeSegNet())
    nf = ref_attr.NetAttrFunc(seg, idx_for_class=[17, 1])
    loss = nf.loss(img)
    loss.backward()
    out.update(seg_logits=np_(seg.net.logits), seg_loss=np_(loss), seg_dlogits=np_(seg.net.logits.grad),
               seg_classes=np.array([17, 1]))
    pred = FakePredictor()
    img2 = randn(26, 2, 3, 8, 8).requires_grad_(True)
    cf = ref_attr.ClassifierAttrFunc(pred, idx_for_class=31, idx_of_interest=1)
    loss = cf.loss(img2)
    loss.backward()
    out.update(cls_logits=np_(pred.logits), cls_loss=np_(loss), cls_dlogits=np_(pred.logits.grad))
    pred2 = FakePredictor()
    cf = ref_attr.ClassifierAttrFunc(pred2, idx_for_class=31, idx_of_interest=0,
                                     regularize_idx_idx_score=(15, 1, torch.tensor([0.3, -0.6])))
    loss = cf.loss(img2)
    loss.backward()
    out.update(cls_reg_loss=np_(loss), cls_reg_dlogits=np_(pred2.logits.grad))
    np.savez_compressed(os.path.join(HERE, "guidance.npz"), **out)
    print("guidance.npz", len(out))


class RecordingUnet(torch.nn.Module):
    """Wraps the toy model and records every eps it returns (for teacher forcing)."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.config = inner.config
        self.in_channels = inner.in_channels
        self.sample_size = inner.sample_size
        self.rec = []

    def forward(self, sample, timestep, **k):
        o = self.inner(sample, timestep, **k)
        self.rec.append(o["sample"].detach().clone())
        return o


def gen_inversion():
    out = {}
    T = 20
    for preset in ("ddpm", "sd"):
        m = make_model(preset, T, clip=False)
        m.unet = RecordingUnet(m.unet)
        x0 = randn(31, 1, C, S, S).clamp(-1, 1) * 0.5
        out[f"{preset}/x0"] = np_(x0)
        # sample_xts_from_x0 alone, noise stream replayed from the same global seed
        torch.manual_seed(77)
        xts = ref_ddpm.sample_xts_from_x0(m, x0, num_inference_steps=T)
        torch.manual_seed(77)
        noises = torch.zeros(T, C, S, S)
        for idx in reversed(range(T)):          # reversed(timesteps) = ascending t = descending idx
            noises[idx] = torch.randn_like(x0)[0]
        out[f"{preset}/fwd_noises"] = np_(noises)
        out[f"{preset}/xts_sampled"] = np_(xts)
        # full edit-friendly inversion
        for eta in (1.0, 0.6):
            m.unet.rec.clear()
            torch.manual_seed(77)
            xt, zs, xts2 = ref_ddpm.invert(m, x0, num_inference_steps=T, eta=eta, prog_bar=False)
            out[f"{preset}/eta{eta}/inv_eps"] = np_(torch.cat(m.unet.rec))  # order: idx T-1 .. 0
            out[f"{preset}/eta{eta}/xT"] = np_(xt)
            out[f"{preset}/eta{eta}/zs"] = np_(zs)
            out[f"{preset}/eta{eta}/xts"] = np_(xts2)
            for tskip in (0, 7):
                m.unet.rec.clear()
                xr = ref_ddpm.sample(m, zs, xts2, Tskip=tskip, eta=eta, prog_bar=False)
                out[f"{preset}/eta{eta}/sample_T{tskip}"] = np_(xr)
                out[f"{preset}/eta{eta}/sample_T{tskip}_eps"] = np_(torch.cat(m.unet.rec))
        # eta = 0 branch (forward_step) and DDIM inversion
        m.unet.rec.clear()
        xt0, zs0, xts0 = ref_ddpm.inversion_forward_process(m, x0, etas=0, num_inference_steps=T)
        assert zs0 is None and xts0 is None
        out[f"{preset}/eta0/xT"] = np_(xt0)
        out[f"{preset}/eta0/eps"] = np_(torch.cat(m.unet.rec))
        m.unet.rec.clear()
        out[f"{preset}/ddim_inv/xT"] = np_(ref_ddim.ddim_inversion(m, x0))
        out[f"{preset}/ddim_inv/eps"] = np_(torch.cat(m.unet.rec))
    out["T"] = np.array(T)
    np.savez_compressed(os.path.join(HERE, "inversion.npz"), **out)
    print("inversion.npz", len(out))


def gen_pipeline():
    out = {}
    T = 20
    m = make_model("ddpm", T, clip=True)
    m.unet = RecordingUnet(m.unet)
    wrapper = DDPM(m)
    pipe = SegDiffEditPipeline(wrapper, None)
    xt = randn(41, 1, C, S, S)
    zs = randn(42, T, C, S, S)
    mask = (randn(43, 1, C, S, S) > 0.3).float()
    out.update(xt=np_(xt), zs=np_(zs), mask=np_(mask), T=np.array(T))

    def run(tag, **kw):
        m.unet.rec.clear()
        r = pipe.edit_image(xt=xt, **kw)
        out[f"{tag}/img"] = np.asarray(r.imgs)
        out[f"{tag}/x0_imgs"] = np.stack([np.asarray(p) for p in r.pred_original_samples])
        out[f"{tag}/eps"] = np_(torch.cat(r.model_outputs))
        assert r[0] is r.imgs

    f = ref_attr.SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=100.0, t1=0, t2=T)
    run("color_eta0", eta=0, attr_func=f)
    f = ref_attr.SingleColorAttrFunc(target=0.8, color_idx=0, loss_scale=100.0, t1=3, t2=15)
    run("color_eta1_window", eta=1.0, zs=zs, attr_func=f)
    f = ref_attr.SingleColorAttrFunc(target=-0.5, color_idx=1, loss_scale=60.0, use_mask=True,
                                     mask_attr_grad=True)
    run("color_maskgrad", eta=0, attr_func=f, mask=mask)
    # unguided sampling through Diffusion.generate_image
    for eta, tag in ((0, "gen_eta0"), (0.8, "gen_eta08")):
        m.unet.rec.clear()
        img, eps_list, x0_imgs, xts_imgs = wrapper.generate_image(
            xt, eta=eta, zs=zs if eta > 0 else None, num_inference_steps=T, return_xts=True)
        out[f"{tag}/img"] = np.asarray(img)
        out[f"{tag}/x0_imgs"] = np.stack([np.asarray(p) for p in x0_imgs])
        out[f"{tag}/xts_imgs"] = np.stack([np.asarray(p) for p in xts_imgs])
        out[f"{tag}/eps"] = np_(torch.cat(eps_list))
    # input validation behaviour (src/SegDiffEditPipeline.py:65-76)
    errs = []
    for kw in (dict(eta=1.0, zs=None, attr_func=f), dict(eta=0, zs=zs, attr_func=f),
               dict(eta=0, attr_func=None, mask=None)):
        try:
            pipe.edit_image(xt=xt, **kw)
            errs.append("none")
        except ValueError as ex:
            errs.append(str(ex))
    out["check_inputs_errors"] = np.array(errs)
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)
    print("pipeline.npz", len(out))


def gen_mask():
    from PIL import Image
    out = {}
    maps = {
        "testimg": np.array(Image.open(os.path.join(REF, "src/Segmentation/res/test_res/test-img.png"))),
        "hair": np.array(Image.open(os.path.join(REF, "src/Segmentation/hair.png"))),
    }
    keys = []
    for name, arr in maps.items():
        out[f"seg/{name}"] = arr.astype(np.uint8)
        seg = torch.from_numpy(arr.astype(np.int64))
        for classes in ([17], [17, 1], [1, 2, 3, 4, 5, 10, 12, 13], [0], [14, 16, 17, 18]):
            for d in (256, 64, 32):
                for dil in (False, True):
                    mc = MaskCreator(dilate_mask=dil, resize_size=(d, d))
                    mk = mc.create_mask(seg, classes)
                    assert mk.shape == (1, 3, d, d)
                    a = np_(mk)
                    assert np.array_equal(a[0, 0], a[0, 1]) and np.array_equal(a[0, 0], a[0, 2])
                    assert set(np.unique(a)) <= {0.0, 1.0}
                    k = f"{name}/c{'_'.join(map(str, classes))}/d{d}/dil{int(dil)}"
                    out["mask/" + k] = np.packbits(a[0, 0].astype(np.uint8))
                    keys.append(k)
    out["keys"] = np.array(keys)
    # raw antialiased resize on a random fp32 image (pins the accumulation order)
    img = randn(51, 1, 1, 96, 80)
    for (oh, ow) in ((48, 40), (12, 10), (37, 23)):
        r = torch.nn.functional.interpolate(img, size=(oh, ow), mode="bilinear", antialias=True)
        out[f"resize/{oh}x{ow}"] = np_(r[0, 0])
    out["resize_in"] = np_(img[0, 0])
    # Morphology with non-trivial structuring elements
    x = randn(52, 1, 1, 40, 40)
    for k in (3, 5, 7):
        d = Dilation2d(1, 1, k, soft_max=False)
        e = Erosion2d(1, 1, k, soft_max=False)
        w = randn(53 + k, 1, 1, k, k) * 0.3
        with torch.no_grad():
            d.weight.copy_(w)
            e.weight.copy_(w)
            out[f"morph/dil{k}"] = np_(d(x)[0, 0])
            out[f"morph/ero{k}"] = np_(e(x)[0, 0])
            ds = Dilation2d(1, 1, k, soft_max=True, beta=20)
            ds.weight.copy_(w)
            out[f"morph/dil_soft{k}"] = np_(ds(x)[0, 0])
        out[f"morph/w{k}"] = np_(w[0, 0])
    out["morph_in"] = np_(x[0, 0])
    np.savez_compressed(os.path.join(HERE, "mask.npz"), **out)
    print("mask.npz", len(out))


def gen_bisenet():
    """The reference's BiSeNet (src/Segmentation/model.py) on seeded weights; its ResNet-18 hub download
    (src/Segmentation/resnet.py:83) is stubbed - the weights are overwritten by the seeded state anyway."""
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}
    from Segmentation.model import BiSeNet
    from oracle.bisenet import seeded_weights
    net = seeded_weights(BiSeNet(19).eval(), 1234)
    x = randn(77, 1, 3, 128, 128)
    with torch.no_grad():
        out = net(x)[0]
    res = {"out_sub4": np_(out[:, :, ::4, ::4]), "argmax": out[0].argmax(0).numpy().astype(np.int16),
           "n_params": np.array(sum(p.numel() for p in net.parameters()), dtype=np.int64)}
    np.savez_compressed(os.path.join(HERE, "bisenet.npz"), **res)
    print("bisenet.npz", len(res))


if __name__ == "__main__":
    torch.set_grad_enabled(True)
    gen_bisenet()
    gen_step_math()
    gen_guidance()
    gen_inversion()
    gen_pipeline()
    gen_mask()

"""CPU tests of the host-side logic of the drop-in modules (no kernel launches): scheduler tables
and timesteps vs the oracle restatement, loop windows, registry, input validation, constants."""
import numpy as np
import pytest
import torch

from oracle.ddim_scheduler import DDIMScheduler as OracleScheduler


@pytest.mark.parametrize("preset", ["ddpm", "ldm", "sd"])
def test_native_scheduler_tables_match_oracle(preset):
    from b200edit.scheduler import DDIMScheduler
    a, b = DDIMScheduler.from_preset(preset), OracleScheduler.from_preset(preset)
    assert torch.equal(a.alphas_cumprod, b.alphas_cumprod)
    assert float(a.final_alpha_cumprod) == float(b.final_alpha_cumprod)
    for T in (50, 20, 7):
        a.set_timesteps(T)
        b.set_timesteps(T)
        assert torch.equal(a.timesteps, b.timesteps) and a.num_inference_steps == T
        assert a.timesteps.device.type == "cpu"          # the loop never syncs on a timestep
        t = int(a.timesteps[T // 2])
        assert a.previous_timestep(t) == t - 1000 // T
        assert float(a._get_variance(t, t - 1000 // T)) == float(b._get_variance(t, t - 1000 // T))
    assert a.config.clip_sample == b.config.clip_sample


def test_step_coefficients_match_reference_scalar_math():
    from b200edit.scheduler import DDIMScheduler
    from oracle import step_math as sm
    s, o = DDIMScheduler.from_preset("sd"), OracleScheduler.from_preset("sd")
    s.set_timesteps(50)
    o.set_timesteps(50)
    for t in [int(x) for x in s.timesteps]:
        oc = sm.step_coeffs(o, t)
        for eta in (0.0, 0.7, 1.0):
            c = s.coeffs(t, eta, "ddpm")
            assert c.sqrt_a_t == float(oc.sqrt_a_t) and c.sqrt_b_t == float(oc.sqrt_b_t)
            assert c.sqrt_a_prev == float(oc.sqrt_a_prev) and c.a_t_sq == float(oc.a_t_sq)
            assert c.dir_coef == float((1 - oc.a_prev - eta * oc.variance) ** 0.5)
            assert c.sigma == float(eta * oc.variance ** 0.5)
            d = s.coeffs(t, eta, "ddim")
            assert d.dir_coef == float((1 - oc.a_prev - (eta * oc.variance ** 0.5) ** 2) ** 0.5)
    assert s.coeffs(981, 0.0, "ddim") is s.coeffs(981, 0.0, "ddim")   # cached


def test_diffusion_loop_windows():
    from types import SimpleNamespace

    from b200edit.scheduler import DDIMScheduler
    from diffusion_utils import diffusion_loop, get_previous_timestep, get_variance_noise
    s = DDIMScheduler.from_preset("ddpm")
    s.set_timesteps(50)
    model = SimpleNamespace(scheduler=s)
    full = list(diffusion_loop(model, None, prog_bar=False))
    assert [i for i, _ in full] == list(range(50)) and int(full[0][1]) == 980 and int(full[-1][1]) == 0
    zs = torch.zeros(14, 3, 4, 4)                      # Tskip = 36 window
    win = list(diffusion_loop(model, zs, prog_bar=False))
    assert [i for i, _ in win] == list(range(14)) and int(win[0][1]) == 13 * 20
    assert get_previous_timestep(model, 980) == 960
    assert get_variance_noise(zs, 3, 0) is None and get_variance_noise(None, 3, 1.0) is None
    assert get_variance_noise(zs, 3, 1.0).shape == (3, 4, 4)


def test_registry_and_attr_func_contract():
    from attr_functions import AnyGANAttrFunc, AttrFunc, ClassifierAttrFunc, MultiColorAttrFunc, SingleColorAttrFunc
    from attr_functions_registry import AttrFuncRegistry, create_attr_func_registry
    reg = create_attr_func_registry()
    assert reg.get_attribute_functions() == ["SingleColorAttrFunc", "MultiColorAttrFunc", "NetAttrFunc", "AnyGANAttrFunc"]
    f = reg.get("SingleColorAttrFunc", dict(target=0.5, color_idx=1, loss_scale=3.0, t1=2, t2=9, use_mask=True))
    assert isinstance(f, SingleColorAttrFunc) and f.name == "SingleColorAttrFunc"
    assert f.kwargs == {"use_mask": True} and f.in_window(2) and not f.in_window(9) and not f.in_window(1)
    assert f.colour_spec() == ([None, 0.5, None], None)
    m = MultiColorAttrFunc(0.1, 0.2, 0.3)
    assert m.colour_spec() == ([0.1, 0.2, 0.3], [0.1, 0.2, 0.3])
    assert issubclass(AnyGANAttrFunc, ClassifierAttrFunc)
    with pytest.raises(ValueError, match="No strategy registered"):
        reg.get("nope")
    inst = SingleColorAttrFunc(target=0.1, color_idx=0)
    r2 = AttrFuncRegistry()
    r2.register(inst)
    assert r2.get("SingleColorAttrFunc") is inst
    with pytest.raises(NotImplementedError):
        SingleColorAttrFunc(target=0.1, color_idx=0, use_lpips=True)
    with pytest.raises(TypeError):
        AttrFunc()                                       # abstract

    class Mine(AttrFunc):
        def loss(self, img, **kw):
            return img.mean()
    assert Mine().colour_spec() is None                  # generic (autograd) path
    # CPU tensors take the differentiable torch path of the loss helpers
    from attr_functions import color_loss, l2_norm, single_color_loss
    x = torch.randn(2, 3, 4, 4)
    assert torch.allclose(single_color_loss(x, 1, 0.3), (x[:, 1] - 0.3).abs().mean())
    assert torch.allclose(color_loss(x, 0.1, 0.2, 0.3),
                          sum((x[:, i] - t).abs().mean() * t for i, t in enumerate((0.1, 0.2, 0.3))))
    assert torch.allclose(l2_norm(x, x * 0), x.pow(2).sum().sqrt())


def test_pipeline_validation_without_gpu():
    from types import SimpleNamespace

    from b200edit.scheduler import DDIMScheduler
    from diffusion_classes import DDPM
    from SegDiffEditPipeline import EditorOutput, SegDiffEditPipeline
    s = DDIMScheduler.from_preset("ddpm")
    s.set_timesteps(10)
    unet = SimpleNamespace(config=SimpleNamespace(in_channels=3, sample_size=8))
    w = DDPM(SimpleNamespace(unet=unet, scheduler=s, device=torch.device("cpu")))
    assert w.decode_is_identity and w.data_dimensionality == 8
    x = torch.zeros(1, 3, 8, 8)
    assert w.encode(x) is x and w.decode(x) is x
    pipe = SegDiffEditPipeline(w, None)
    with pytest.raises(ValueError, match="eta > 0 and zs is empty"):
        pipe.edit_image(xt=x, eta=1.0)
    with pytest.raises(ValueError, match="eta == 0 and zs is not empty"):
        pipe.edit_image(xt=x, eta=0, zs=torch.zeros(10, 3, 8, 8))
    with pytest.raises(ValueError, match="implies no edit"):
        pipe.edit_image(xt=x, eta=0)
    with pytest.raises(ValueError, match="not possible"):
        pipe.prepare_real_image_edit(x, eta=1.0, inversion_method="ddim")
    with pytest.raises(AssertionError):
        pipe.check_classes([19])
    pipe.check_classes(None)
    out = EditorOutput("img", ["a"], ["b"])
    assert out[0] == "img" and out.to_tuple() == ("img", ["a"], ["b"]) and out["model_outputs"] == ["b"]


def test_transforms_host_side():
    from PIL import Image

    from transforms import pil_to_tensor
    a = (np.arange(4 * 5 * 3) % 256).astype(np.uint8).reshape(4, 5, 3)
    t = pil_to_tensor(Image.fromarray(a))
    assert t.shape == (1, 3, 4, 5) and torch.allclose(t, torch.from_numpy(a).permute(2, 0, 1)[None] / 255 * 2 - 1)
    assert pil_to_tensor([Image.fromarray(a)] * 2).shape == (2, 3, 4, 5)
    with pytest.raises(Exception):
        pil_to_tensor(3)

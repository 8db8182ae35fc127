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


def test_batchnorm_folding_of_the_loss_networks_cpu():
    """Host-side parameter preparation of the native classifier / face parser: eval-mode BatchNorm folded into the
    convolution weights and biases reproduces conv -> BN exactly (fp64 folding, fp32 comparison), and every parameter the
    reference modules own is mapped to an engine parameter name."""
    import importlib.util
    import os
    import torch
    import torchvision
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):   # import the wrapper modules without loading libb200edit.so
        import sys
        import types
        pkg = types.ModuleType("b200edit_stub")
        src = open(os.path.join(here, "diffusion-image-editing_b200", "b200edit", name + ".py")).read()
        start = src.index("def _fold") if "def _fold" in src else None
        return src, pkg, start

    # ---- ResNet.fold_batchnorm (static, pure torch): compare conv+bn against the folded conv on a random input
    src, _, _ = load("resnet")
    ns = {}
    body = src[src.index("    @staticmethod\n    def fold_batchnorm"):src.index("    def load_torchvision_state_dict")]
    exec("import torch\nclass R:\n" + body, ns)
    net = torchvision.models.resnet18().eval()
    g = torch.Generator().manual_seed(0)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    sd = net.state_dict()
    folded = ns["R"].fold_batchnorm(sd)
    convs = [k[:-7] for k, v in sd.items() if k.endswith(".weight") and v.dim() == 4]
    assert sorted(k[:-7] for k in folded if k.endswith(".weight") and folded[k].dim() == 4) == sorted(convs)
    x = torch.randn(2, 64, 8, 8, generator=g)
    blk = net.layer1[0]
    with torch.no_grad():
        want = blk.bn1(blk.conv1(x))
        got = torch.nn.functional.conv2d(x, folded["layer1.0.conv1.weight"], folded["layer1.0.conv1.bias"], padding=1)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
        ds = net.layer2[0].downsample
        want = ds(x)
        got = torch.nn.functional.conv2d(x, folded["layer2.0.downsample.0.weight"], folded["layer2.0.downsample.0.bias"], stride=2)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)

    # ---- BiSeNet.fold_reference_state_dict: every used parameter of the reference module gets an engine name
    from oracle.bisenet import BiSeNet as OracleBiSeNet, seeded_weights
    src, _, _ = load("bisenet")
    ns2 = {}
    fold_src = src[src.index("def _fold"):src.index("class BiSeNet")]
    meth = src[src.index("    @staticmethod\n    def fold_reference_state_dict"):src.index("    def load_reference_state_dict")]
    exec("import torch\n" + fold_src + "\nclass B:\n" + meth, ns2)
    o = seeded_weights(OracleBiSeNet(19).eval(), 3)
    f2 = ns2["B"].fold_reference_state_dict(o.state_dict())
    for name in ("cp.resnet.conv1", "cp.resnet.layer2.0.downsample.0", "cp.arm16.conv", "cp.arm32.conv_atten", "cp.conv_head32",
                 "cp.conv_avg", "ffm.convblk", "conv_out.conv", "conv_out.conv_out"):
        assert name + ".weight" in f2 and name + ".bias" in f2, name
    assert "ffm.conv1.weight" in f2 and f2["ffm.conv1.weight"].shape == (64, 256) and f2["cp.conv_avg.weight"].shape == (128, 512)
    assert not any(k.startswith(("conv_out16", "conv_out32")) for k in f2)
    xa = torch.randn(1, 256, 8, 8, generator=g)
    with torch.no_grad():
        want = o.cp.arm16.conv(xa)
        got = torch.relu(torch.nn.functional.conv2d(xa, f2["cp.arm16.conv.weight"], f2["cp.arm16.conv.bias"], padding=1))
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_precision_selection_of_the_engines_cpu():
    """Precision names are validated on the host before any device work: the 16-bit type the library was not built for and
    unknown names raise instead of silently computing in another precision (b200edit/_C.py::resolve_precision); the face
    parser wrapper validates its own argument and defaults to f16 operands for mask creation / fp32-accurate forward for
    differentiated calls (b200edit/bisenet.py::MultiResBiSeNet)."""
    import pytest
    from b200edit import _C
    from b200edit.bisenet import MultiResBiSeNet
    fast = _C.fast_precision()
    assert fast in ("fp16", "bf16")
    assert _C.resolve_precision(None, "t") == (fast, 0) and _C.resolve_precision(fast, "t") == (fast, 0)
    assert _C.resolve_precision("fp32", "t") == ("fp32", 1)
    other = "bf16" if fast == "fp16" else "fp16"
    with pytest.raises(ValueError, match="built for"):
        _C.resolve_precision(other, "t")
    with pytest.raises(ValueError, match="precision must be"):
        _C.resolve_precision("fp8", "t")
    with pytest.raises(ValueError, match="precision"):
        MultiResBiSeNet(precision="fp64")
    net = MultiResBiSeNet(19, max_batch=1, device="cpu", precision=None)     # engines are built lazily: no device work here
    assert net.precision is None and net._engines == {}
    with pytest.raises(ValueError, match="multiple of 32"):
        net.engine(100)

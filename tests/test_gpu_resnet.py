"""GPU numerics of the native classifier network (ClassifierAttrFunc's predictor: torchvision resnet50 with an 80-way fc,
src/models.py:69-77) against torchvision itself - the reference's own dependency - in fp32 (torch eager on the GPU as the
checker, TF32 off) with the same weights, eval-mode BatchNorm with non-trivial running statistics.

Two modes.  precision="fp32" (what models.get_pretrained_anyGAN builds): fp32-accurate forward (split f16 operands), so
the backward pass routes gradients through the fp32 network's ReLU / max-pool masks - logits relative RMS <= 5e-5, input
gradient relative RMS <= 2e-2 with cosine >= 0.999 against fp32 autograd (measured 1e-3 .. 8e-3 / 0.99997).
precision="fp16" (f16 operands throughout, fp32 accumulation): logits relative RMS <= 2e-3 (measured 4e-4).  There the
INPUT GRADIENT of the reference's loss (one logit of sample 0) is ill-conditioned: the 4e-4 forward rounding flips ReLU /
max-pool masks of near-zero activations and every flip re-routes gradient paths - torch autograd through the same
network in bf16 is 0.22 (ResNet-18, 64x64) to 0.49 (ResNet-50, 512x512) relative RMS away from the fp32 gradient; the
fp16 bars are 0.2 / cosine 0.98 (measured 0.04 .. 0.15) and no worse than torch's bf16 autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu
tv = pytest.importorskip("torchvision")


def make_reference(kind, num_classes, seed):
    torch.manual_seed(seed)
    net = (tv.models.resnet50 if kind == "resnet50" else tv.models.resnet18)()
    net.fc = torch.nn.Linear(net.fc.in_features, num_classes)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):   # a trained network's statistics are not (0, 1)
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    return net.eval()


def run_pair(kind, S, B, num_classes, seed, pick=(0, 31, 0), precision=None):
    from b200edit.resnet import ResNet
    ref_net = make_reference(kind, num_classes, seed)
    layers = (3, 4, 6, 3) if kind == "resnet50" else (2, 2, 2, 2)
    native = ResNet("bottleneck" if kind == "resnet50" else "basic", layers, num_classes, S, max_batch=B, precision=precision)
    native.load_torchvision_state_dict(ref_net.state_dict())
    x = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(seed + 2)).mul(2).sub(1)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def loss_of(logits):   # ClassifierAttrFunc.loss: attr.view(-1, 40, 2)[0][idx_for_class][idx_of_interest]
        return logits.view(-1, num_classes // 2, 2)[pick[0]][pick[1]][pick[2]]

    xn = x.cuda().requires_grad_(True)
    ln = native(xn)
    gn, = torch.autograd.grad(loss_of(ln), xn)
    ref_net = ref_net.cuda()
    xr = x.cuda().requires_grad_(True)
    lr = ref_net(xr)
    gr, = torch.autograd.grad(loss_of(lr), xr)
    net16 = ref_net.bfloat16()
    x16 = x.cuda().bfloat16().requires_grad_(True)
    l16 = net16(x16)
    g16, = torch.autograd.grad(loss_of(l16.float()), x16)
    return ln.detach(), lr.detach(), l16.detach().float(), gn, gr, g16.float()


def rel(a, b):
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()


def check(ln, lr, l16, gn, gr, g16, tag):
    print(f"{tag}: logits rel-rms native {rel(ln, lr):.3e} torch-bf16 {rel(l16, lr):.3e} | input-gradient rel-rms native "
          f"{rel(gn, gr):.3e} cos {cos(gn, gr):.5f} torch-bf16 {rel(g16, gr):.3e} cos {cos(g16, gr):.5f}")
    assert ln.shape == lr.shape and gn.shape == gr.shape and torch.isfinite(ln).all() and torch.isfinite(gn).all()
    assert rel(ln, lr) <= 2e-3 and rel(ln, lr) <= 1.25 * rel(l16, lr) + 1e-3      # logits: measured 3.5e-4 .. 5e-4
    # input gradient with fp16 operands: measured 3.7e-2 / 0.9993 (ResNet-18, 64x64), 0.14 / 0.990 (ResNet-50, 256x256),
    # 0.147 / 0.989 (512x512) - 3x closer to fp32 than the bf16 build of round 1 (0.41 / 0.917) and than torch-bf16
    # autograd; what remains are ReLU / max-pool masks flipped by the 4e-4 forward rounding (see the module docstring)
    assert rel(gn, gr) <= 0.2 and cos(gn, gr) >= 0.98
    assert rel(gn, gr) <= rel(g16, gr) + 5e-3 and cos(gn, gr) >= cos(g16, gr) - 1e-3
    assert gn.shape[0] == 1 or gn[1:].abs().max().item() == 0.0      # the reference's loss only sees batch element 0


def check_accurate(ln, lr, l16, gn, gr, g16, tag):
    """precision="fp32": fp32-accurate forward (split operands) -> the masks of the backward pass are the fp32 network's."""
    print(f"{tag} [fp32-accurate forward]: logits rel-rms native {rel(ln, lr):.3e} | input-gradient rel-rms native "
          f"{rel(gn, gr):.3e} cos {cos(gn, gr):.5f}")
    assert torch.isfinite(ln).all() and torch.isfinite(gn).all()
    assert rel(ln, lr) <= 5e-5
    assert rel(gn, gr) <= 2e-2 and cos(gn, gr) >= 0.999


@pytest.mark.parametrize("kind,S,B,K,seed,pick", [("resnet18", 64, 2, 16, 1, (0, 3, 1)), ("resnet50", 256, 2, 80, 2, (0, 31, 0)),
                                                  ("resnet50", 512, 1, 80, 3, (0, 31, 0))])
def test_resnet_fp32_accurate_forward_gives_fp32_grade_input_gradient(kind, S, B, K, seed, pick):
    check_accurate(*run_pair(kind, S, B, K, seed=seed, pick=pick, precision="fp32"), f"{kind} {S}x{S}")


def test_resnet18_small_matches_torchvision():
    check(*run_pair("resnet18", 64, 2, 16, seed=1, pick=(0, 3, 1)), "resnet18 64x64")


def test_resnet50_256_matches_torchvision():
    check(*run_pair("resnet50", 256, 2, 80, seed=2), "resnet50 256x256")


def test_resnet50_512_matches_torchvision():
    """The size config 4 runs it at (SD decodes to 512x512)."""
    check(*run_pair("resnet50", 512, 1, 80, seed=3), "resnet50 512x512")


@pytest.mark.parametrize("precision,rel_bar,cos_bar", [("fp16", 0.2, 0.98), ("fp32", 2e-2, 0.999)])
def test_classifier_attr_func_with_native_predictor(precision, rel_bar, cos_bar):
    """ClassifierAttrFunc.apply with the native predictor (autograd node backed by the native dgrad) against the same
    strategy with torchvision's module: same update direction on x_t.  "fp32" (fp32-accurate forward) is what
    models.get_pretrained_anyGAN builds."""
    from attr_functions import ClassifierAttrFunc
    from models import create_diffusion_model
    from b200edit.resnet import ResNet
    ref_net = make_reference("resnet50", 80, 5)
    native = ResNet("bottleneck", (3, 4, 6, 3), 80, 64, max_batch=1, precision=precision)
    native.load_torchvision_state_dict(ref_net.state_dict())
    ref_net = ref_net.cuda()
    cfg = dict(sample_size=64, in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
               down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))
    w = create_diffusion_model("ddpm", sample_clipping=False, max_batch=1, seed=1, unet_config=cfg)
    w.scheduler.set_timesteps(10)
    g = torch.Generator().manual_seed(9)
    xt = torch.randn(1, 3, 64, 64, generator=g).cuda()
    eps = torch.randn(1, 3, 64, 64, generator=g).cuda()
    t = int(w.scheduler.timesteps[5])
    net16 = make_reference("resnet50", 80, 5).cuda().bfloat16()
    outs = []
    for pred in (native, ref_net, lambda x: net16(x.bfloat16()).float()):
        f = ClassifierAttrFunc(pred, idx_for_class=31, idx_of_interest=0, loss_scale=50.0)
        f.kwargs["mask"] = None
        x2, _ = f.apply(xt=xt.clone(), zt=None, model_output=eps, timestep=torch.tensor(t), step_idx=0, model=w, **f.kwargs)
        outs.append((x2 - xt).detach())
    dn, dr, d16 = outs
    print(f"ClassifierAttrFunc update [{precision}]: native rel-rms {rel(dn, dr):.3e} cos {cos(dn, dr):.5f} | torch-bf16 predictor "
          f"{rel(d16, dr):.3e} cos {cos(d16, dr):.5f}")
    assert dr.abs().max() > 0 and rel(dn, dr) <= rel_bar and cos(dn, dr) >= cos_bar      # measured 0.146 / 0.989 (fp16)
    assert rel(dn, dr) <= rel(d16, dr) + 2e-2


def test_metrics_harness_on_the_engine():
    """src/metrics.py drop-in (evaluation harness): generation, classifier-guided editing and the attribute predictor all
    on the engine; contract of the two metrics (keys, shapes, ranges, determinism under a fixed seed)."""
    import metrics
    from attr_functions import AnyGANAttrFunc
    from b200edit.resnet import resnet50_predictor
    from models import create_diffusion_model
    from SegDiffEditPipeline import SegDiffEditPipeline
    cfg = dict(sample_size=64, in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
               down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))
    w = create_diffusion_model("ddpm", sample_clipping=True, max_batch=1, seed=1, unet_config=cfg)
    predictor = resnet50_predictor(80, 64, max_batch=1, seed=2)
    func = AnyGANAttrFunc(predictor=predictor, idx_for_class=31, loss_scale=500.0, t1=0, t2=8)
    editor = SegDiffEditPipeline(w, None)

    def run():
        g = torch.Generator().manual_seed(5)
        acc = metrics.attribute_consistency(editor, w, func, 2, g, num_inference_steps=8, predictor=predictor)
        d0, d1 = metrics.avg_increase_decrease_per_attribute(editor, w, func, 2, g, num_inference_steps=8, predictor=predictor)
        return acc, d0, d1

    acc, d0, d1 = run()
    assert acc.shape == (40,) and float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0
    assert len(d0) == 40 and len(d1) == 40 and "31 Smiling" in d0 and "39 Young" in d1
    assert all(torch.isfinite(torch.tensor(list(d0.values())))) and any(abs(v) > 0 for v in d0.values())
    acc2, d0b, _ = run()
    assert torch.equal(acc, acc2) and d0 == d0b
    with pytest.raises(NotImplementedError):
        metrics.lpips(None, None)

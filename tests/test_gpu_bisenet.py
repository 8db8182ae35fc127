"""GPU numerics of the native face parser (BiSeNet behind SegmentationModel, src/models.py:80-118) against the oracle
restatement (pinned bit-exactly to the unmodified reference module, tests/test_oracle_golden.py) run in fp32 with the
same seeded weights and non-trivial eval-mode BatchNorm statistics.  Tolerance (IEEE f16 operands, fp32 accumulation):
logits relative RMS <= 3e-3 (measured 8e-4 .. 9e-4) and no worse than 1.25x the oracle itself run in bf16 by torch; the
parsing map (argmax) may differ from the fp32 one only where the top two logits are within rounding noise: mismatch
rate <= 1.25x torch-bf16's (measured 3e-4 .. 5e-4 vs 6e-3).  The differentiated parser runs the fp32-accurate forward
(precision="fp32"): logits <= 1e-4, input gradient relative RMS <= 2e-2 / cosine >= 0.999 vs fp32 autograd (measured
1.2e-3 .. 1.4e-3 / 1.00000); with f16 operands the gradient bar is 0.12 / 0.99 (ReLU / max-pool masks flip)."""
import pytest
import torch

from oracle.bisenet import BiSeNet as OracleBiSeNet, seeded_weights

pytestmark = pytest.mark.gpu


def run_pair(S, B, seed):
    from b200edit.bisenet import BiSeNet
    oracle = seeded_weights(OracleBiSeNet(19).eval(), seed)
    native = BiSeNet(19, S, max_batch=B)
    native.load_reference_state_dict(oracle.state_dict())
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(seed + 1))
    got = native(x.cuda())[0]
    torch.cuda.synchronize()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        oc = oracle.cuda()
        ref = oc(x.cuda())[0]
        ref16 = oc.bfloat16()(x.cuda().bfloat16())[0].float()
    return got, ref, ref16


def check(got, ref, ref16, tag):
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    rel16 = ((ref16 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mis = (got.argmax(1) != ref.argmax(1)).float().mean().item()
    mis16 = (ref16.argmax(1) != ref.argmax(1)).float().mean().item()
    print(f"{tag}: logits rel-rms native {rel:.3e} torch-bf16 {rel16:.3e} | parsing-map mismatch native {mis:.3e} torch-bf16 {mis16:.3e}")
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel <= 3e-3 and rel <= 1.25 * rel16 + 1e-3      # measured 8e-4 .. 9e-4 (bf16 build of round 1: 8e-3)
    assert mis <= 1.25 * mis16 + 2e-3


def test_bisenet_128_matches_oracle():
    check(*run_pair(128, 2, seed=3), "bisenet 128x128 B=2")


def test_bisenet_512_matches_oracle():
    """The size SegmentationModel runs it at (src/models.py:84: image_size (512, 512))."""
    check(*run_pair(512, 1, seed=4), "bisenet 512x512")


def test_segmentation_model_with_native_parser_feeds_the_mask_path():
    """SegmentationModel(net=native BiSeNet)(image) -> int parsing map on the device -> MaskCreator (the hot mask path)."""
    from b200edit.bisenet import BiSeNet
    from mask_creator import MaskCreator
    from models import SegmentationModel
    oracle = seeded_weights(OracleBiSeNet(19).eval(), 5)
    native = BiSeNet(19, 512, max_batch=1)
    native.load_reference_state_dict(oracle.state_dict())
    seg_model = SegmentationModel(net=native)
    img = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(6)).mul(2).sub(1).cuda()
    seg = seg_model(img)
    assert seg.shape == (512, 512) and seg.dtype == torch.int64 and seg.is_cuda and int(seg.max()) < 19
    cls = int(torch.bincount(seg.flatten()).argmax())
    mask = MaskCreator(dilate_mask=True, resize_size=(64, 64)).create_mask(seg, classes=[cls])
    assert mask.shape == (1, 3, 64, 64) and float(mask.max()) == 1.0
    with pytest.raises(Exception):
        native(torch.zeros(1, 3, 512, 512, device="cuda", requires_grad=True))   # enable_grad() not called


def grad_pair(S, seed, classes=(1, 10, 13), precision=None):
    """Gradient of NetAttrFunc.loss (softmax area of the selected classes, src/attr_functions.py:213-219) w.r.t. the image."""
    from b200edit.bisenet import BiSeNet
    oracle = seeded_weights(OracleBiSeNet(19).eval(), seed)
    oracle.conv_out.conv_out.weight.data.mul_(0.02)   # logits of O(1): an unsaturated softmax, as a trained parser has
    native = BiSeNet(19, S, max_batch=1, precision=precision)
    native.load_reference_state_dict(oracle.state_dict())
    native.enable_grad()
    x = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(seed + 1))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def loss_of(out):
        p = out.squeeze(0).softmax(dim=0)
        return (p.sum(dim=(1, 2)) / (256 * 256))[list(classes)].sum()

    xn = x.cuda().requires_grad_(True)
    gn, = torch.autograd.grad(loss_of(native(xn)[0]), xn)
    oc = oracle.cuda()
    xr = x.cuda().requires_grad_(True)
    gr, = torch.autograd.grad(loss_of(oc(xr)[0]), xr)
    o16 = oc.bfloat16()
    x16 = x.cuda().bfloat16().requires_grad_(True)
    g16, = torch.autograd.grad(loss_of(o16(x16)[0].float()), x16)
    return gn, gr, g16.float()


@pytest.mark.parametrize("S", [128, 256])
def test_bisenet_input_gradient_matches_autograd(S):
    """Input gradient through the native parser (dgrad twins, attention / pooling / bilinear / max-pool backward kernels)
    against torch autograd through the fp32 oracle; yardstick = torch autograd through the same network in bf16 (ReLU /
    max-pool masks flip under rounding, see tests/test_gpu_resnet.py): no further from fp32 than that, cosine >= 0.9."""
    gn, gr, g16 = grad_pair(S, seed=7)
    rel = lambda a, b: ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()   # noqa: E731
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()   # noqa: E731
    print(f"bisenet {S}x{S} input gradient: native rel-rms {rel(gn, gr):.3e} cos {cos(gn, gr):.5f} | torch-bf16 {rel(g16, gr):.3e} cos {cos(g16, gr):.5f}")
    assert torch.isfinite(gn).all() and gr.abs().max() > 0
    assert rel(gn, gr) <= 0.12 and cos(gn, gr) >= 0.99      # measured 7.7e-2 .. 8.8e-2 / 0.996 .. 0.997 (round 1, bf16: 0.23 / 0.975)
    assert rel(gn, gr) <= rel(g16, gr) + 1e-2 and cos(gn, gr) >= cos(g16, gr) - 2e-3


@pytest.mark.parametrize("S", [128, 256])
def test_bisenet_fp32_accurate_forward_gives_fp32_grade_input_gradient(S):
    """precision="fp32" (what MultiResBiSeNet / SegmentationModel run a differentiated parser in): the forward is
    fp32-accurate, so the backward pass routes gradients through the fp32 network's ReLU / max-pool masks.  Bar: the
    fp32 grade the review asked for - relative RMS <= 2e-2 and cosine >= 0.999 against fp32 autograd."""
    from b200edit.bisenet import BiSeNet
    gn, gr, _ = grad_pair(S, seed=7, precision="fp32")
    rel = ((gn - gr).pow(2).mean().sqrt() / gr.pow(2).mean().sqrt()).item()
    cos = torch.nn.functional.cosine_similarity(gn.flatten(), gr.flatten(), dim=0).item()
    oracle = seeded_weights(OracleBiSeNet(19).eval(), 3)
    native = BiSeNet(19, S, max_batch=2, precision="fp32")
    native.load_reference_state_dict(oracle.state_dict())
    x = torch.randn(2, 3, S, S, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():
        lrel = ((native(x)[0] - oracle.cuda()(x)[0]).pow(2).mean().sqrt() / oracle(x)[0].pow(2).mean().sqrt()).item()
    print(f"bisenet {S}x{S} [fp32-accurate forward]: logits rel-rms {lrel:.3e} | input gradient rel-rms {rel:.3e} cos {cos:.5f}")
    assert torch.isfinite(gn).all() and gr.abs().max() > 0
    assert lrel <= 1e-4
    assert rel <= 2e-2 and cos >= 0.999


@pytest.mark.parametrize("precision,rel_bar,cos_bar", [("fp16", 0.12, 0.99), ("fp32", 2e-2, 0.999)])
def test_net_attr_func_through_the_native_parser(precision, rel_bar, cos_bar):
    """NetAttrFunc.apply with SegmentationModel(native parser in gradient mode): update direction on x_t vs the torch module.
    "fp32" is the mode SegmentationModel() runs a differentiated parser in (MultiResBiSeNet)."""
    from attr_functions import NetAttrFunc
    from b200edit.bisenet import BiSeNet
    from models import SegmentationModel, create_diffusion_model
    oracle = seeded_weights(OracleBiSeNet(19).eval(), 8)
    oracle.conv_out.conv_out.weight.data.mul_(0.02)   # logits of O(1): an unsaturated softmax, as a trained parser has
    native = BiSeNet(19, 256, max_batch=1, precision=precision)
    native.load_reference_state_dict(oracle.state_dict())
    native.enable_grad()
    cfg = dict(sample_size=256, in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
               down_block_types=("DownBlock2D", "DownBlock2D"), up_block_types=("UpBlock2D", "UpBlock2D"))
    w = create_diffusion_model("ddpm", sample_clipping=False, max_batch=1, seed=1, unet_config=cfg)
    w.scheduler.set_timesteps(10)
    g = torch.Generator().manual_seed(9)
    xt = torch.randn(1, 3, 256, 256, generator=g).cuda()
    eps = torch.randn(1, 3, 256, 256, generator=g).cuda()
    t = int(w.scheduler.timesteps[-1])     # late step (alpha_bar ~ 1) and a large scale: the area loss is normalised by 256^2
    outs = []
    for net in (native, oracle.cuda()):
        f = NetAttrFunc(SegmentationModel(net=net, image_size=(256, 256)), idx_for_class=[1, 10], loss_scale=1e4)
        f.kwargs["mask"] = None
        x2, _ = f.apply(xt=xt.clone(), zt=None, model_output=eps, timestep=torch.tensor(t), step_idx=0, model=w, **f.kwargs)
        outs.append((x2 - xt).detach())
    dn, dr = outs
    rel = ((dn - dr).pow(2).mean().sqrt() / dr.pow(2).mean().sqrt()).item()
    cos = torch.nn.functional.cosine_similarity(dn.flatten(), dr.flatten(), dim=0).item()
    print(f"NetAttrFunc update through the native parser [{precision}]: rel-rms {rel:.3e} cos {cos:.5f}")
    assert dr.abs().max() > 0 and rel <= rel_bar and cos >= cos_bar      # measured 6.7e-2 / 0.998 (fp16)


def test_one_default_segmentation_model_serves_mask_creation_and_net_attr_func():
    """The reference workflow (src/models.py:80-118, src/attr_functions.py:202-219): ONE SegmentationModel() - the
    reference's positional signature (ckpt, n_classes, image_size) - produces the 512x512 parsing map for the mask AND is
    the loss network of NetAttrFunc, which feeds it the decoded 256x256 image (the reference's BiSeNet is fully
    convolutional).  The native parser keeps one engine per resolution behind the same object."""
    from attr_functions import NetAttrFunc
    from mask_creator import MaskCreator
    from models import SegmentationModel, create_diffusion_model
    seg_model = SegmentationModel("no/such/checkpoint.pth", 19, (512, 512))      # positional, as the reference is called
    with pytest.raises(TypeError):
        SegmentationModel(torch.nn.Identity())                                     # a network is NOT the first argument
    img = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(6)).mul(2).sub(1).cuda()
    seg = seg_model(img)
    assert seg.shape == (512, 512) and seg.dtype == torch.int64
    mask = MaskCreator(dilate_mask=True, resize_size=(256, 256)).create_mask(seg, classes=[int(torch.bincount(seg.flatten()).argmax())])
    assert mask.shape == (1, 3, 256, 256)
    cfg = dict(sample_size=256, in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
               down_block_types=("DownBlock2D", "DownBlock2D"), up_block_types=("UpBlock2D", "UpBlock2D"))
    w = create_diffusion_model("ddpm", sample_clipping=False, max_batch=1, seed=1, unet_config=cfg)
    w.scheduler.set_timesteps(10)
    g = torch.Generator().manual_seed(9)
    xt = torch.randn(1, 3, 256, 256, generator=g).cuda()
    eps = torch.randn(1, 3, 256, 256, generator=g).cuda()
    f = NetAttrFunc(seg_model, idx_for_class=[1, 10], loss_scale=1e4)
    f.kwargs["mask"] = None
    x2, _ = f.apply(xt=xt.clone(), zt=None, model_output=eps, timestep=torch.tensor(int(w.scheduler.timesteps[-1])), step_idx=0,
                    model=w, **f.kwargs)
    assert torch.isfinite(x2).all() and (x2 - xt).abs().max() > 0
    # same parser, same weights at both resolutions: the 512-engine still answers after the 256-engine was used
    assert torch.equal(seg_model(img), seg)

"""GPU numerics of the native face parser (BiSeNet behind SegmentationModel, src/models.py:80-118) against the oracle
restatement (pinned bit-exactly to the unmodified reference module, tests/test_oracle_golden.py) run in fp32 with the
same seeded weights and non-trivial eval-mode BatchNorm statistics.  Tolerance (bf16 operands, fp32 accumulation):
logits relative RMS <= 2.5e-2 and no worse than 1.25x the oracle itself run in bf16 by torch; the parsing map (argmax)
may differ from the fp32 one only where the top two logits are within bf16 noise: mismatch rate <= 1.25x torch-bf16's."""
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
    assert rel <= 2.5e-2 and rel <= 1.25 * rel16 + 1e-3
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
        native(img.requires_grad_(True))

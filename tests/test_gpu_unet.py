"""GPU numerics of the native UNet (fp16 tensor-core operands, fp32 accumulation) against the oracle
UNet2DModel run in fp32 with the same weights.

Tolerance (stated): the north star's bar for the 16-bit mode, LITERALLY - max-abs error of the predicted noise
<= 1e-2 (|eps| <= 1.9 on these inputs; measured 1.4e-3 .. 3.2e-3) - and relative RMS <= 3e-3 (measured 0.9e-3 ..
1.6e-3; the bf16 build of round 1 sat at 8e-3 / 1.45e-2).  As a yardstick the same test evaluates the ORACLE itself in
fp16 with torch (what the reference's diffusers + PyTorch path gives in half precision) and requires the native engine
to be no further from the fp32 result than 1.5x that."""
import pytest
import torch

from oracle.unet2d import DDPM256_CONFIG, LDM_CELEBAHQ_CONFIG, UNet2DModel as OracleUNet

pytestmark = pytest.mark.gpu

SMALL = dict(sample_size=32, in_channels=3, out_channels=3, block_out_channels=(64, 128, 128),
             layers_per_block=1, down_block_types=("DownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
             up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D"), norm_num_groups=32, norm_eps=1e-6)


# LDM-style layout in small: channel counts that are multiples of 32 but not of 64 (zero-padded internally),
# multi-head attention with 32-channel heads at two resolutions, symmetric stride-2 padding, flipped embedding
LDM_SMALL = dict(sample_size=32, in_channels=3, out_channels=3, block_out_channels=(32, 96, 160),
                 layers_per_block=1, down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"), norm_num_groups=32, norm_eps=1e-6,
                 attention_head_dim=32, flip_sin_to_cos=True, freq_shift=0, downsample_padding=1)


def run_pair(cfg, B, t, seed, precision=None):
    from b200edit.unet import UNet2DModel
    torch.manual_seed(seed)
    oracle = OracleUNet(**cfg).eval()
    native = UNet2DModel(**cfg, max_batch=B, precision=precision)
    native.load_state_dict(oracle.state_dict())
    x = torch.randn(B, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"],
                    generator=torch.Generator().manual_seed(seed + 1))
    got = native(x.cuda(), t)["sample"]
    torch.cuda.synchronize()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref = oracle.cuda()(x.cuda(), torch.tensor(t))["sample"]   # torch fp32 eager as the checker
        ref16 = oracle.half()(x.cuda().half(), torch.tensor(t))["sample"].float()
    return got, ref, ref16


def check(got, ref, ref16, tag):
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    err16 = (ref16 - ref).abs().max().item()
    rel16 = ((ref16 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"{tag}: native max-abs {err:.3e} rel-rms {rel:.3e} | torch-fp16 max-abs {err16:.3e} rel-rms {rel16:.3e}"
          f" | max|eps| {scale:.3f}")
    assert torch.isfinite(got).all()
    assert err <= 1e-2 and rel <= 3e-3          # the literal 1e-2 max-abs bar of the 16-bit mode
    assert rel <= 1.5 * rel16 + 2e-4


def check_fp32(got, ref, ref16, tag):
    """fp32-accurate mode (split-fp16 operands, three products per GEMM): the north star's fp32 bar,
    max-abs <= 1e-4 (x max|eps| when that exceeds 1), and relative RMS <= 5e-5."""
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"{tag}: fp32-accurate native max-abs {err:.3e} rel-rms {rel:.3e} | max|eps| {scale:.3f}")
    assert torch.isfinite(got).all()
    assert err <= 1e-4 * max(1.0, scale) and rel <= 5e-5


@pytest.mark.parametrize("B,t", [(1, 980), (3, 500)])
def test_small_unet_fp32_mode_matches_oracle(B, t):
    check_fp32(*run_pair(SMALL, B, t, seed=B, precision="fp32"), f"small unet fp32 B={B} t={t}")


def test_ldm_style_unet_fp32_mode_matches_oracle():
    check_fp32(*run_pair(LDM_SMALL, 2, 400, seed=12, precision="fp32"), "ldm-style small unet fp32")


def test_ldm_celebahq_unet_fp32_mode_matches_oracle():
    """Full LDM-CelebAHQ layout in the fp32-accurate mode: padded channel pitches (224 -> 256, 672 -> 704) carry three
    planes, 14-28 heads of 32 channels go through the fp32 attention core."""
    check_fp32(*run_pair(LDM_CELEBAHQ_CONFIG, 1, 500, seed=4, precision="fp32"), "ldm-celebahq unet fp32")


def test_ddpm256_unet_fp32_mode_matches_oracle():
    check_fp32(*run_pair(DDPM256_CONFIG, 2, 500, seed=0, precision="fp32"), "ddpm-256 unet fp32")


@pytest.mark.parametrize("B,t", [(1, 980), (3, 500), (2, 0)])
def test_small_unet_matches_oracle(B, t):
    check(*run_pair(SMALL, B, t, seed=B), f"small unet B={B} t={t}")


@pytest.mark.parametrize("B,t", [(1, 981), (3, 400)])
def test_ldm_style_unet_matches_oracle(B, t):
    check(*run_pair(LDM_SMALL, B, t, seed=10 + B), f"ldm-style small unet B={B} t={t}")


def test_ldm_celebahq_unet_matches_oracle():
    """The full LDM-CelebAHQ UNet2DModel layout (224/448/672/896 channels, 14-28 heads of 32 channels)."""
    check(*run_pair(LDM_CELEBAHQ_CONFIG, 2, 500, seed=4), "ldm-celebahq unet")


def test_ddpm256_unet_matches_oracle():
    check(*run_pair(DDPM256_CONFIG, 2, 500, seed=0), "ddpm-256 unet")


def test_batch_rebuild_and_per_sample_timesteps():
    from b200edit.unet import UNet2DModel
    torch.manual_seed(3)
    oracle = OracleUNet(**SMALL).eval()
    native = UNet2DModel(**SMALL, max_batch=4)
    native.load_state_dict(oracle.state_dict())
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(9)).cuda()
    full = native(x, 300)["sample"].clone()
    part = native(x[:2], 300)["sample"].clone()   # smaller batch: plan is rebuilt
    again = native(x, 300)["sample"]
    torch.cuda.synchronize()
    assert torch.equal(full[:2], part) and torch.equal(full, again)
    ts = torch.tensor([300, 300, 10, 900])
    mixed = native(x, ts)["sample"]
    assert torch.equal(mixed[:2], full[:2]) and not torch.equal(mixed[2:], full[2:])

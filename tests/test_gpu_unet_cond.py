"""GPU numerics of the native UNet2DConditionModel (SD 1.x layout: Transformer2DModel blocks with self-attention,
cross-attention over the text tokens and GEGLU) against the oracle restatement in fp32 with the same weights.
Tolerance as for the other networks: relative RMS <= 2e-2, max-abs <= 3e-2 * max|eps| and no worse than 1.25x the oracle
itself run in bf16 by torch."""
import pytest
import torch

from oracle.unet2d_condition import SD15_CONFIG, UNet2DConditionModel as OracleCond

pytestmark = pytest.mark.gpu

SMALL = dict(sample_size=16, in_channels=4, out_channels=4, block_out_channels=(64, 128), layers_per_block=1,
             cross_attention_dim=64, attention_head_dim=2, norm_num_groups=32, norm_eps=1e-5)
# head_dim 40 (not a multiple of 64, like SD's 320 / 8), three levels: 32 / 16 / 8 -> 1024 / 256 / 64 tokens
MID = dict(sample_size=32, in_channels=4, out_channels=4, block_out_channels=(320, 640, 640), layers_per_block=1,
           cross_attention_dim=128, attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5)


def run_pair(cfg, B, t, seed, L=77):
    from b200edit.unet_cond import UNet2DConditionModel
    torch.manual_seed(seed)
    oracle = OracleCond(**cfg).eval()
    native = UNet2DConditionModel(**cfg, max_batch=B)
    native.load_state_dict(oracle.state_dict())
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"], generator=g)
    ctx = torch.randn(B, L, cfg["cross_attention_dim"], generator=g)
    got = native(x.cuda(), t, encoder_hidden_states=ctx.cuda())["sample"]
    torch.cuda.synchronize()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        oc = oracle.cuda()
        ref = oc(x.cuda(), torch.tensor(t), encoder_hidden_states=ctx.cuda())["sample"]
        ref16 = oc.bfloat16()(x.cuda().bfloat16(), torch.tensor(t), encoder_hidden_states=ctx.cuda().bfloat16())["sample"].float()
    return got, ref, ref16


def check(got, ref, ref16, tag):
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    err16 = (ref16 - ref).abs().max().item()
    rel16 = ((ref16 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"{tag}: native max-abs {err:.3e} rel-rms {rel:.3e} | torch-bf16 max-abs {err16:.3e} rel-rms {rel16:.3e}"
          f" | max|eps| {scale:.3f}")
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel <= 2e-2 and err <= 3e-2 * max(1.0, scale)
    assert rel <= 1.25 * rel16 + 1e-3


@pytest.mark.parametrize("B,t,L", [(1, 981, 77), (2, 401, 77), (2, 1, 5)])
def test_small_cond_unet_matches_oracle(B, t, L):
    check(*run_pair(SMALL, B, t, seed=B + L, L=L), f"small cond unet B={B} t={t} L={L}")


def test_mid_cond_unet_matches_oracle():
    check(*run_pair(MID, 2, 500, seed=9), "cond unet 320/640/640, head_dim 40/80")


def test_sd15_cond_unet_matches_oracle():
    """The full Stable Diffusion 1.x UNet layout (320/640/1280/1280, 859.5 M parameters, 4096-token self-attention)."""
    check(*run_pair(SD15_CONFIG, 2, 500, seed=2), "sd-1.x cond unet")


def test_cfg_call_of_the_reference():
    """get_noise_pred's CFG branch (src/diffusion_utils.py:61-70) on the native conditional UNet."""
    from types import SimpleNamespace
    from b200edit.unet_cond import UNet2DConditionModel
    from diffusion_utils import get_noise_pred
    torch.manual_seed(3)
    oracle = OracleCond(**SMALL).eval()
    native = UNet2DConditionModel(**SMALL, max_batch=2)
    native.load_state_dict(oracle.state_dict())
    model = SimpleNamespace(unet=native, device=torch.device("cuda"))
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(1, 4, 16, 16, generator=g).cuda()
    text = torch.randn(2, 77, 64, generator=g).cuda()
    eps = get_noise_pred(model, lat, torch.tensor(500), text, 3.5)
    with torch.no_grad():
        oc = oracle.cuda()
        both = oc(torch.cat([lat] * 2), torch.tensor(500), encoder_hidden_states=text)["sample"]
        u, cnd = both.chunk(2)
        ref = u + 3.5 * (cnd - u)
        b16 = oc.bfloat16()(torch.cat([lat] * 2).bfloat16(), torch.tensor(500), encoder_hidden_states=text.bfloat16())["sample"].float()
        u16, c16 = b16.chunk(2)
        ref16 = u16 + 3.5 * (c16 - u16)
    # the combination u + s (c - u) amplifies the per-branch bf16 error by up to (1 + 2 s): yardstick = torch in bf16
    rel = ((eps - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    rel16 = ((ref16 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"cfg 3.5: native rel-rms {rel:.3e} | torch-bf16 {rel16:.3e}")
    assert eps.shape == lat.shape and rel <= 8e-2 and rel <= 1.25 * rel16 + 1e-3, (rel, rel16)

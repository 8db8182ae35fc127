"""GPU numerics of the tcgen05 implicit-GEMM convolution against a plain PyTorch fp32 reference
(cuDNN is used here only as the checker).  Inputs and weights are pre-rounded to the engine's 16-bit type (fp16; bf16
in -DB2E_ACT_BF16 builds) so the only differences are fp32 accumulation order and the final 16-bit rounding of the
output: tolerance = 2^-10 (fp16) / 2^-7 (bf16) relative to the output scale."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # N, H, W, Cin, Cout, k, stride
    (1, 16, 16, 64, 128, 3, 1),
    (2, 32, 32, 128, 128, 3, 1),
    (1, 256, 256, 128, 128, 3, 1),     # one output row = two 128-pixel tiles
    (2, 64, 64, 256, 256, 3, 1),
    (3, 8, 8, 512, 512, 3, 1),         # tiles span two images; odd batch -> OOB image
    (1, 8, 8, 1024, 512, 3, 1),
    (2, 16, 16, 512, 1536, 1, 1),      # attention qkv projection
    (2, 16, 16, 384, 256, 1, 1),       # 1x1 shortcut, 6 K chunks
    (2, 32, 32, 128, 128, 3, 2),       # Downsample2D: stride 2, pad (0,1,0,1)
    (1, 256, 256, 128, 128, 3, 2),
    (2, 16, 16, 64, 64, 3, 1),         # BLOCK_N = 64
    (1, 4, 4, 64, 64, 3, 1),           # tiny map: 8 images per tile
    # cluster split-K (64-wide tiles, the K-splits of a tile = one thread-block cluster meeting in the leader's smem):
    (8, 16, 16, 512, 256, 3, 1),       # 64 tiles  -> 2 splits
    (3, 16, 16, 512, 512, 3, 1),       # 48 tiles  -> 3 splits
    (8, 8, 8, 512, 512, 3, 1),         # 32 tiles  -> 4 splits (the batch-8 layers of the 8x8 level)
    (14, 8, 8, 1024, 256, 3, 1),       # 28 tiles, 144 k-blocks -> 5 splits
    (3, 16, 16, 1024, 256, 3, 1),      # 24 tiles, 144 k-blocks -> 6 splits
    (2, 8, 8, 256, 256, 3, 1),         # 4 tiles, 36 k-blocks -> 3 splits
    (1, 16, 16, 1024, 512, 3, 1),      # 16 tiles  -> 7 splits, 144 k-blocks (uneven split boundaries)
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_vs_torch(case):
    from b200edit import ops
    dt = ops.act_dtype()
    rel = 2 ** -10 if dt == torch.float16 else 2 ** -7
    N, H, W, Cin, Cout, k, stride = case
    g = torch.Generator(device="cpu").manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dt)
    b = torch.randn(Cout, generator=g)
    xd = x.cuda().permute(0, 2, 3, 1).contiguous()
    out = ops.conv2d_nhwc_f16(xd, w.float().cuda(), b.cuda(), stride=stride)
    torch.cuda.synchronize()
    xf = x.float().cuda()
    if stride == 2:
        ref = F.conv2d(F.pad(xf, (0, 1, 0, 1)), w.float().cuda(), b.cuda(), stride=2)
    else:
        ref = F.conv2d(xf, w.float().cuda(), b.cuda(), padding=k // 2)
    got = out.float().permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rel * scale + 1e-3 * rel * 128, f"max abs err {err} (scale {scale})"
    # fused residual segment (identity weights appended to K): out = conv(x) + r
    r = torch.randn(ref.shape, generator=g).to(dt)
    out_r = ops.conv2d_nhwc_f16(xd, w.float().cuda(), b.cuda(), stride=stride,
                                 residual=r.cuda().permute(0, 2, 3, 1).contiguous()) if stride == 1 else None
    if out_r is not None:
        ref_r = ref + r.float().cuda()
        err = (out_r.float().permute(0, 3, 1, 2) - ref_r).abs().max().item()
        scale = ref_r.abs().max().item()
        assert err <= rel * scale + 1e-3 * rel * 128, f"residual: max abs err {err} (scale {scale})"


def test_split_k_is_deterministic():
    """The cluster split-K sums the partial tiles in split order: two runs of the same launch are bit-identical."""
    from b200edit import ops
    dt = ops.act_dtype()
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(1, 8, 8, 512, generator=g).to(dt).cuda()
    w = (torch.randn(512, 512, 3, 3, generator=g) / 68.0).cuda()
    b = torch.randn(512, generator=g).cuda()
    outs = [ops.conv2d_nhwc_f16(x, w, b) for _ in range(4)]
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])


UP2_CASES = [
    # N, H, W, Cin, Cout  (low-resolution input)
    (1, 8, 8, 64, 64),        # tiny: tiles span the whole image, every border tap is out of bounds somewhere
    (2, 16, 16, 128, 128),
    (3, 8, 16, 64, 128),      # non-square, odd batch
    (1, 64, 64, 128, 128),    # the 64 -> 128 level in small
    (2, 128, 128, 128, 128),  # DDPM-256's last upsampler (128 -> 256)
    (1, 32, 32, 256, 192),    # Cout not a multiple of 128 (64-wide N tiles)
]


@pytest.mark.parametrize("case", UP2_CASES, ids=[str(c) for c in UP2_CASES])
def test_upsample_conv_as_subpixel_phases_vs_torch(case):
    """Upsample2D (nearest x2 + conv3x3) computed as four 2x2 phase convolutions on the low-resolution input against
    torch's interpolate + conv2d in fp32 on the same 16-bit-rounded operands.  Tolerance as above; the pre-summed phase
    weights are rounded once to 16 bit, so the weight-rounding part of the error is 2^-11 of |w_a + w_b| per tap."""
    from b200edit import ops
    dt = ops.act_dtype()
    rel = 2 ** -10 if dt == torch.float16 else 2 ** -7
    N, H, W, Cin, Cout = case
    g = torch.Generator(device="cpu").manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5)
    b = torch.randn(Cout, generator=g)
    out = ops.upsample_conv3x3_nhwc_f16(x.cuda().permute(0, 2, 3, 1).contiguous(), w.cuda(), b.cuda())
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(x.float().cuda(), scale_factor=2.0, mode="nearest"), w.cuda(), b.cuda(), padding=1)
    got = out.float().permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2 * rel * scale, f"max abs err {err} (scale {scale})"

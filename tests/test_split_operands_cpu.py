"""The fp32-accurate mode's operand split: x = hi + lo with hi = f16(x), lo = f16(x - hi) carries ~22
mantissa bits, and x_hi W_hi + x_lo W_hi + x_hi W_lo (fp32 accumulation; weights pre-scaled by 2^10 so their lo halves stay
out of the fp16 subnormals, accumulator scaled back by 2^-10) reproduces the fp32 product to ~2^-21 relative - pinned here
on the CPU in float64, independent of the CUDA code."""
import torch

WSCALE = 2.0 ** 10      # csrc/unet.cu: kSplitWScale


def split(v: torch.Tensor):
    hi = v.to(torch.float16)
    lo = (v - hi.to(v.dtype)).to(torch.float16)
    return hi, lo


def test_split_representation_has_22_bits():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1 << 16, generator=g, dtype=torch.float32) * 3.0
    hi, lo = split(x)
    err = (hi.double() + lo.double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)
    # hi keeps 11 bits, lo another 11 of the remainder (unless it falls into the fp16 subnormals: |x| < 2^-3 here is rare)
    big = x.abs() > 2.0 ** -2
    assert err[big].max().item() <= 2.0 ** -21


def test_three_product_gemm_matches_fp32_product():
    g = torch.Generator().manual_seed(2)
    M, K, N = 64, 1152, 32
    x = torch.randn(M, K, generator=g, dtype=torch.float32)
    w = torch.randn(N, K, generator=g, dtype=torch.float32) / K ** 0.5
    xh, xl = split(x)
    wh, wl = split(w * WSCALE)
    # fp16 x fp16 products are exact in fp32; model the fp32 accumulation with float64 (the tensor core's accumulation error
    # is a separate, measured term: DESIGN.md section 2)
    acc = xh.double() @ wh.double().T + xl.double() @ wh.double().T + xh.double() @ wl.double().T
    got = acc / WSCALE
    ref = x.double() @ w.double().T
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    worst = ((got - ref).abs().max() / ref.abs().max()).item()
    # dropped term x_lo W_lo ~ 2^-22 relative per product
    assert rel <= 2.0 ** -20 and worst <= 2.0 ** -18, (rel, worst)
    # without the 2^10 weight scale the lo halves of small weights underflow: the error is visibly larger
    wh0, wl0 = split(w)
    got0 = xh.double() @ wh0.double().T + xl.double() @ wh0.double().T + xh.double() @ wl0.double().T
    rel0 = ((got0 - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel0 >= rel

"""The sub-pixel phase decomposition the engine uses for Upsample2D (oracle/upsample_phases.py) equals
nearest-upsample x2 + conv3x3 - pinned on the CPU in float64 (exact up to summation order) for odd shapes too."""
import pytest
import torch

from oracle import upsample_phases as up


@pytest.mark.parametrize("shape", [(1, 3, 1, 1, 2), (2, 5, 4, 6, 3), (1, 8, 7, 5, 4), (3, 2, 16, 16, 2)])
def test_phase_decomposition_equals_upsample_then_conv(shape):
    N, Cin, H, W, Cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(Cout, generator=g, dtype=torch.float64)
    ref = up.upsample_conv_reference(x, w, b)
    got = up.upsample_conv_by_phases(x, w, b)
    assert got.shape == ref.shape == (N, Cout, 2 * H, 2 * W)
    assert (got - ref).abs().max().item() <= 1e-12 * max(1.0, ref.abs().max().item())


def test_phase_weights_fold_every_tap_exactly_once():
    w = torch.arange(9, dtype=torch.float64).reshape(1, 1, 3, 3) + 1.0
    total = sum(up.phase_weights(w, a, b).sum().item() for a in range(2) for b in range(2))
    # every 3x3 tap contributes to each of the four phases exactly once
    assert total == 4 * w.sum().item()
    assert up.phase_weights(w, 0, 0)[0, 0].tolist() == [[1.0, 2.0 + 3.0], [4.0 + 7.0, 5.0 + 6.0 + 8.0 + 9.0]]
    assert up.phase_weights(w, 1, 1)[0, 0].tolist() == [[1.0 + 2.0 + 4.0 + 5.0, 3.0 + 6.0], [7.0 + 8.0, 9.0]]

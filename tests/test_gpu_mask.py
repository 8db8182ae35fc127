"""GPU parity of the mask path (bit-exact): CUDA vs the reference goldens on its fixture parsing
maps and vs the numpy oracle on seeded random maps."""
import numpy as np
import pytest
import torch

from oracle import mask as omask

pytestmark = pytest.mark.gpu


def test_mask_creator_fixtures(golden):
    from b200edit import ops
    g = golden("mask")
    n = 0
    for k in g["keys"]:
        k = str(k)
        name, cls, d, dil = k.split("/")
        classes = [int(c) for c in cls[1:].split("_")]
        d = int(d[1:])
        seg = torch.from_numpy(g[f"seg/{name}"].astype(np.int64)).cuda()
        m = ops.mask_from_seg(seg, classes, dil == "dil1", (d, d)).cpu().numpy()
        ref = np.unpackbits(g["mask/" + k])[: d * d].reshape(d, d).astype(np.float32)
        assert m.shape == (1, 3, d, d)
        for c in range(3):
            assert np.array_equal(m[0, c], ref), k
        n += 1
    assert n == 60


def test_resize_matches_aten_order(golden):
    from b200edit import ops
    g = golden("mask")
    x = torch.from_numpy(g["resize_in"]).cuda()
    for k in [f for f in g.files if f.startswith("resize/")]:
        oh, ow = map(int, k.split("/")[1].split("x"))
        got = ops.resize_bilinear_aa(x, (oh, ow)).cpu().numpy()
        assert np.array_equal(got, g[k]), k


@pytest.mark.parametrize("hw,out", [((96, 80), (48, 40)), ((100, 60), (37, 23)), ((64, 64), (64, 64)),
                                     ((40, 56), (80, 112)), ((512, 512), (64, 64))])
def test_random_maps_vs_oracle(hw, out):
    """Ragged sizes, up- and down-scaling, repeated and absent classes, empty class list."""
    from b200edit import ops
    rng = np.random.RandomState(hw[0] * 7 + out[0])
    # blobby map: low-res random labels upsampled, so masks have interiors and borders
    small = rng.randint(0, 19, size=(hw[0] // 8 + 1, hw[1] // 8 + 1))
    seg = np.kron(small, np.ones((8, 8), dtype=np.int64))[: hw[0], : hw[1]]
    segd = torch.from_numpy(seg).cuda()
    for classes in ([3], [1, 2, 17], [5, 5, 6], [30], []):
        for dil in (False, True):
            got = ops.mask_from_seg(segd, classes, dil, out, channels=4).cpu().numpy()
            ref = omask.create_mask(seg, classes, dil, out, channels=4)
            assert np.array_equal(got, ref), (classes, dil)


def test_morphology_vs_golden(golden):
    from b200edit import ops
    g = golden("mask")
    x = torch.from_numpy(g["morph_in"])[None, None].cuda()
    for k in (3, 5, 7):
        w = torch.from_numpy(g[f"morph/w{k}"])[None, None].cuda()
        assert np.array_equal(ops.morphology2d(x, w, "dilation2d")[0, 0].cpu().numpy(), g[f"morph/dil{k}"])
        assert np.array_equal(ops.morphology2d(x, w, "erosion2d")[0, 0].cpu().numpy(), g[f"morph/ero{k}"])
        soft = ops.morphology2d(x, w, "dilation2d", soft_max=True, beta=20.0)[0, 0].cpu().numpy()
        # logsumexp in fp32 with different exp/log implementations: 1e-5 absolute
        assert np.allclose(soft, g[f"morph/dil_soft{k}"], atol=1e-5, rtol=1e-5)

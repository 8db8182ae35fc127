"""GPU parity of the guidance loss heads (NetAttrFunc / ClassifierAttrFunc) against
the goldens recorded from the unmodified reference heads and against the oracle on larger random inputs.

Tolerances: the softmax head is fp32 with a different summation order over classes / pixels than ATen's
vectorised softmax + sum -> rtol 1e-5 on the gradient (relative to its max), 1e-6 on the loss;
the classifier head is index arithmetic -> bit-exact (the squared regulariser: 1 ulp)."""
import numpy as np
import pytest
import torch

from oracle import step_math as sm

pytestmark = pytest.mark.gpu


def test_seg_area_head_golden(golden):
    from b200edit import ops
    g = golden("guidance")
    logits = torch.from_numpy(g["seg_logits"]).cuda()
    classes = [int(c) for c in g["seg_classes"]]
    loss, grad = ops.seg_area_head(logits, classes)
    assert np.allclose(loss.item(), g["seg_loss"], rtol=1e-6)
    ref = g["seg_dlogits"]
    assert grad.shape == logits.shape
    assert np.allclose(grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())


@pytest.mark.parametrize("shape,classes", [((19, 256, 256), [17, 1]), ((19, 64, 48), [0]), ((5, 7, 9), [4, 2, 0]),
                                           ((32, 16, 16), list(range(0, 32, 3)))])
def test_seg_area_head_vs_oracle(shape, classes):
    from b200edit import ops
    g = torch.Generator().manual_seed(sum(shape))
    logits = (3.0 * torch.randn(1, *shape, generator=g)).requires_grad_(True)
    loss_ref = sm.segmentation_area_loss(logits, classes)
    loss_ref.backward()
    loss, grad = ops.seg_area_head(logits.detach().cuda(), classes)
    assert np.allclose(loss.item(), loss_ref.item(), rtol=2e-6, atol=1e-9)
    ref = logits.grad.numpy()
    assert np.allclose(grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-6 * max(np.abs(ref).max(), 1e-12))


def test_seg_area_head_all_classes():
    """Selecting every class: the loss is HW / 65536 and the gradient is rounding noise around 0."""
    from b200edit import ops
    logits = 3.0 * torch.randn(1, 32, 16, 16, generator=torch.Generator().manual_seed(1))
    loss, grad = ops.seg_area_head(logits.cuda(), list(range(32)))
    assert np.allclose(loss.item(), 256 / 65536.0, rtol=1e-6)
    assert grad.abs().max().item() < 1e-11


def test_seg_area_autograd_through_network():
    """The fused head seeds the backward pass of a user-supplied torch parser (the reference's contract)."""
    from types import SimpleNamespace
    from attr_functions import NetAttrFunc
    torch.manual_seed(3)
    net = torch.nn.Conv2d(3, 19, 3, padding=1).cuda()
    seg = SimpleNamespace(net=lambda x: (net(x),))
    f = NetAttrFunc(seg, idx_for_class=[17, 1])
    img = torch.randn(1, 3, 32, 32, device="cuda", requires_grad=True)
    loss = f.loss(img)
    (gx,) = torch.autograd.grad(loss, img)
    img2 = img.detach().clone().requires_grad_(True)
    ref = sm.segmentation_area_loss(net(img2), [17, 1])
    (gr,) = torch.autograd.grad(ref, img2)
    assert np.allclose(loss.item(), ref.item(), rtol=1e-5)
    assert torch.allclose(gx, gr, rtol=1e-4, atol=1e-6 * gr.abs().max().item())


def test_classifier_head_golden(golden):
    from b200edit import ops
    g = golden("guidance")
    lg = torch.from_numpy(g["cls_logits"]).cuda()
    loss, grad = ops.classifier_head(lg, 31, 1)
    assert loss.item() == g["cls_loss"]
    assert np.array_equal(grad.cpu().numpy(), g["cls_dlogits"])
    loss, grad = ops.classifier_head(lg, 31, 0, (15, 1, torch.tensor([0.3, -0.6])))
    assert np.allclose(loss.item(), g["cls_reg_loss"], rtol=1e-6)
    assert np.allclose(grad.cpu().numpy(), g["cls_reg_dlogits"], rtol=1e-6)


def test_classifier_autograd():
    from attr_functions import ClassifierAttrFunc
    torch.manual_seed(5)
    pred = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 80)).cuda()
    f = ClassifierAttrFunc(pred, idx_for_class=7, idx_of_interest=1,
                           regularize_idx_idx_score=(3, 0, torch.tensor([0.25, -0.5])))
    x = torch.randn(2, 3, 8, 8, device="cuda", requires_grad=True)
    (gx,) = torch.autograd.grad(f.loss(x), x)
    x2 = x.detach().clone().requires_grad_(True)
    ref = sm.classifier_logit_loss(pred(x2), 7, 1, (3, 0, torch.tensor([0.25, -0.5])))
    (gr,) = torch.autograd.grad(ref, x2)
    assert torch.allclose(gx, gr, rtol=1e-5, atol=1e-7)
    assert gx[1].abs().max().item() == 0.0      # only batch element 0 is used (src/attr_functions.py:239)


def test_head_errors():
    from b200edit import ops
    from b200edit._C import B2EError
    lg = torch.randn(1, 19, 8, 8, device="cuda")
    with pytest.raises(B2EError):
        ops.seg_area_head(lg, [19])                 # class id out of range
    with pytest.raises(ValueError):
        ops.seg_area_head(torch.randn(2, 19, 8, 8, device="cuda"), [1])   # the reference squeezes batch 1
    with pytest.raises(B2EError):
        ops.classifier_head(torch.randn(1, 79, device="cuda"), 3, 0)

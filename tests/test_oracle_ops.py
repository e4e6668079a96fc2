"""Cross-checks inside the oracle: plain-numpy stencil restatements vs torch's CPU kernels, Dice closed forms,
sliding-window invariants.  (The Dice / sliding-window restatements follow MONAI's public definitions; MONAI is
absent here, so these are pinned by hand-computed cases only -- "parity unpinned" beyond that, see DESIGN.md.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_direct, group_norm, trilinear_upsample2x, dice_loss, dice_loss_grad, sliding_window_inference


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("dims", [(6, 8, 10), (5, 7, 9)])
def test_conv3d_direct_vs_torch(stride, dims):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3) + dims)
    w = rng.standard_normal((4, 3, 3, 3, 3))
    ref = F.conv3d(torch.from_numpy(x), torch.from_numpy(w), stride=stride, padding=1).numpy()
    np.testing.assert_allclose(conv3d_direct(x, w, stride, 1), ref, atol=1e-10)


def test_conv3d_1x1_vs_torch():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 5, 4, 4, 6))
    w = rng.standard_normal((3, 5, 1, 1, 1))
    ref = F.conv3d(torch.from_numpy(x), torch.from_numpy(w)).numpy()
    np.testing.assert_allclose(conv3d_direct(x, w, 1, 0), ref, atol=1e-10)


@pytest.mark.parametrize("groups,c", [(8, 16), (4, 4), (1, 1)])
def test_group_norm_vs_torch(groups, c):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, c, 4, 6, 8)) * 3 + 1
    g, b = rng.standard_normal(c), rng.standard_normal(c)
    ref = F.group_norm(torch.from_numpy(x), groups, torch.from_numpy(g), torch.from_numpy(b), 1e-5).numpy()
    np.testing.assert_allclose(group_norm(x, groups, g, b), ref, atol=1e-10)


@pytest.mark.parametrize("dims", [(4, 6, 8), (3, 5, 7), (1, 2, 3)])
def test_trilinear_vs_torch(dims):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3) + dims)
    ref = F.interpolate(torch.from_numpy(x), scale_factor=2, mode="trilinear", align_corners=False).numpy()
    np.testing.assert_allclose(trilinear_upsample2x(x), ref, atol=1e-12)


def test_dice_hand_computed():
    # one sample, one channel, 4 voxels: p = sigmoid(0) = .5 everywhere; t = [1,1,0,0]
    logits = torch.zeros(1, 1, 1, 2, 2, dtype=torch.float64)
    t = torch.tensor([1, 1, 0, 0], dtype=torch.uint8).view(1, 1, 1, 2, 2)
    I, P, T = 1.0, 2.0, 2.0
    expect = 1 - (2 * I + 1e-5) / (P + T + 1e-5)
    assert abs(float(dice_loss(logits, t)) - expect) < 1e-15
    # include_background=False drops channel 0 when C > 1; batch=True pools over n
    logits2 = torch.zeros(2, 2, 1, 2, 2, dtype=torch.float64)
    t2 = torch.zeros(2, 2, 1, 2, 2, dtype=torch.uint8)
    t2[:, 1, 0, 0, :] = 1
    f = 1 - (2 * 1.0 + 1e-5) / (2.0 + 2.0 + 1e-5)
    assert abs(float(dice_loss(logits2, t2, include_background=False)) - f) < 1e-15
    fb = 1 - (2 * 2.0 + 1e-5) / (4.0 + 4.0 + 1e-5)
    assert abs(float(dice_loss(logits2, t2, include_background=False, batch=True)) - fb) < 1e-15
    # jaccard: 1 - (2I+e)/(2(P+T-I)+e)
    fj = 1 - (2 * I + 1e-5) / (2 * (P + T - I) + 1e-5)
    assert abs(float(dice_loss(logits, t, jaccard=True)) - fj) < 1e-15


def test_dice_grad_closed_form_vs_autograd():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 4, 5, 6, dtype=torch.float64, generator=g, requires_grad=True)
    t = (torch.rand(2, 3, 4, 5, 6, generator=g) > 0.6).to(torch.uint8)
    dice_loss(x, t).backward()
    assert float((x.grad - dice_loss_grad(x.detach(), t)).abs().max()) < 1e-14


def test_sliding_window_identity_and_coverage():
    x = torch.randn(1, 2, 20, 24, 28, generator=torch.Generator().manual_seed(1))
    for mode in ("constant", "gaussian"):
        y = sliding_window_inference(x, (16, 16, 16), lambda p: p * 2.0, overlap=0.25, mode=mode)
        assert float((y - 2 * x).abs().max()) < 1e-5
    calls = []
    sliding_window_inference(torch.zeros(1, 1, 32, 32, 32), (16, 16, 16), lambda p: (calls.append(1), p)[1], overlap=0.25)
    assert len(calls) == 27   # roi/volume = 1/2 with overlap .25 -> 3 starts per axis

"""Pins the Dice and sliding-window restatements (oracle/unet3d_oracle.py) to known-answer cases of MONAI's own unit
tests.  MONAI is not installed here and its repository is not reachable from this sandbox, so the cases below are
quoted from memory of ``tests/test_dice_loss.py`` (TEST_CASES: inputs, kwargs and the 6-digit expected values) and of
``tests/test_sliding_window_inference.py`` (``compute = lambda data: data + 1`` must give ``inputs + 1`` for every
tiling, overlap and blending mode).  Each Dice value was re-derived by hand before it was written down (see the
comments); if the quoted values ever disagree with MONAI's file, MONAI's file wins.  The CUDA kernels are held to the
same cases in tests/test_gpu_prepost.py."""
import numpy as np
import pytest
import torch

from oracle import dice_loss, sliding_window_inference

# sigmoid([1, -1, -1, 1]) = [.731059, .268941, .268941, .731059]
X1 = torch.tensor([[[[1.0, -1.0], [-1.0, 1.0]]]])
T1 = torch.tensor([[[[1.0, 0.0], [1.0, 1.0]]]])
X2 = torch.tensor([[[[1.0, -1.0], [-1.0, 1.0]]], [[[1.0, -1.0], [-1.0, 1.0]]]])
T2 = torch.tensor([[[[1.0, 1.0], [1.0, 1.0]]], [[[1.0, 0.0], [1.0, 0.0]]]])

MONAI_DICE_CASES = [
    # I = 1.731059, P = 2, T = 3: 1 - (2I + 1e-6) / (5 + 1e-6)
    (dict(include_background=True, sigmoid=True, smooth_nr=1e-6, smooth_dr=1e-6), X1, T1, 0.307576),
    # sample 0: I = 2, P = 2, T = 4 -> 0.333328; sample 1: I = 1, P = 2, T = 2 -> 0.499988; mean
    (dict(include_background=True, sigmoid=True, smooth_nr=1e-4, smooth_dr=1e-4), X2, T2, 0.416657),
    # jaccard: 1 - (2I + 1e-5) / (2 (P + T - I) + 1e-5)
    (dict(include_background=True, sigmoid=True, jaccard=True, smooth_nr=1e-5, smooth_dr=1e-5), X1, T1, 0.470451),
    # squared_pred: P = sum p^2 = 1.213552, T = sum t^2 = 3
    (dict(include_background=True, sigmoid=True, squared_pred=True, smooth_nr=1e-5, smooth_dr=1e-5), X1, T1, 0.178337),
]


@pytest.mark.parametrize("kw,x,t,expected", MONAI_DICE_CASES)
def test_dice_restatement_reproduces_monai_published_values(kw, x, t, expected):
    got = float(dice_loss(x.double(), t.double(), **kw))
    assert abs(got - expected) < 2e-6


# (image shape, roi, sw_batch_size, overlap, mode) in the style of MONAI's TEST_CASES (3-D rows)
MONAI_SW_CASES = [
    ((1, 3, 16, 15, 7), (4, 10, 7), 3, 0.25, "constant"),
    ((2, 3, 16, 15, 7), (4, 10, 7), 3, 0.25, "gaussian"),
    ((1, 3, 16, 15, 7), (20, 22, 23), 10, 0.25, "constant"),      # roi larger than the image
    ((1, 3, 16, 15, 7), (4, 4, 4), 1, 0.5, "gaussian"),
    ((1, 1, 33, 17, 40), (16, 16, 24), 4, 0.6, "constant"),
]


@pytest.mark.parametrize("shape,roi,swb,overlap,mode", MONAI_SW_CASES)
def test_sliding_window_restatement_identity_property(shape, roi, swb, overlap, mode):
    x = torch.randn(shape, generator=torch.Generator().manual_seed(0))
    out = sliding_window_inference(x, roi, lambda data: data + 1, overlap=overlap, mode=mode, sw_batch_size=swb)
    np.testing.assert_allclose(out.numpy(), (x + 1).numpy(), rtol=1e-4, atol=1e-5)

"""Pins oracle/prepost_oracle.py (label map <-> one-hot restatements) to fixtures produced by the UNMODIFIED reference
``unet3d/utils/one_hot.py`` (tests/golden/make_golden_prepost.py), and checks the NormalizeIntensity restatement on
hand-computed cases (MONAI absent: parity unpinned, named so)."""
import os
import sys

import numpy as np

from oracle.prepost_oracle import one_hot_encode, label_map_from_one_hot, normalize_intensity

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from recipe_prepost import ONE_HOT_CASES, LABEL_MAP_CASES, label_map_input, prediction_input  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "prepost.npz"))


def test_one_hot_restatement_matches_reference_fixtures():
    for name, (shape, values, n_labels, labels, seed) in ONE_HOT_CASES.items():
        got = one_hot_encode(label_map_input(shape, values, seed).numpy(), n_labels, labels)
        shp = tuple(int(v) for v in GOLD["one_hot_shape::" + name])
        ref = np.unpackbits(GOLD["one_hot::" + name])[: int(np.prod(shp))].reshape(shp)
        assert got.shape == shp and (got == ref).all(), name
        assert got.sum() > 0, name


def test_label_map_restatement_matches_reference_fixtures():
    for name, (shape, labels, kw, seed) in LABEL_MAP_CASES.items():
        got = label_map_from_one_hot(prediction_input(shape, seed).numpy(), labels, **kw)
        ref = GOLD["label_map::" + name]
        assert got.shape == ref.shape and (got == ref).all(), name
        assert (ref != 0).any(), name


def test_normalize_intensity_hand_computed_unpinned():
    x = np.array([[[0.0, 2.0], [4.0, 6.0]]], dtype=np.float32)              # one channel: mean 3, population std sqrt(5)
    y = normalize_intensity(x)
    assert np.allclose(y, (x - 3.0) / np.sqrt(5.0), atol=1e-6)
    ynz = normalize_intensity(x, nonzero=True)                                # non-zero voxels: 2,4,6 -> mean 4, std sqrt(8/3)
    assert ynz[0, 0, 0] == 0.0 and np.allclose(ynz[0].ravel()[1:], (np.array([2, 4, 6]) - 4.0) / np.sqrt(8.0 / 3.0), atol=1e-6)
    two = np.stack([x[0], 10.0 * x[0] + 1.0])
    yc = normalize_intensity(two, channel_wise=True)
    assert np.allclose(yc[0], yc[1], atol=1e-5)                               # per-channel affine invariance
    const = np.full((1, 2, 2), 7.0, dtype=np.float32)
    assert np.allclose(normalize_intensity(const), 0.0)                       # std == 0 -> divide by 1

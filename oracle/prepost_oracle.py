"""CPU restatement of the reference's label-map <-> one-hot conversions and of MONAI's NormalizeIntensity.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Parity status: ``one_hot_encode`` and ``label_map_from_one_hot`` are PINNED against the reference's own
``unet3d/utils/one_hot.py`` (imported in the build container by ``tests/golden/make_golden_prepost.py``; fixtures in
``tests/golden/prepost.npz``).  ``normalize_intensity`` restates monai.transforms.NormalizeIntensity from its documented
behaviour -- MONAI is absent from this image: parity unpinned.
"""
from __future__ import annotations

import numpy as np


def one_hot_encode(data: np.ndarray, n_labels: int, labels=None, do_round: bool = True) -> np.ndarray:
    """unet3d/utils/one_hot.py:7-37: data (n, 1, ...) -> uint8 (n, n_labels, ...); np.rint == torch.round (half to even)."""
    x = np.asarray(data, dtype=np.float32)
    while x.ndim < 5:
        x = x[None]
    assert x.shape[1] == 1
    if do_round:
        x = np.rint(x)
    y = np.zeros((x.shape[0], n_labels) + x.shape[2:], dtype=np.uint8)
    for i in range(n_labels):
        group = [i + 1] if labels is None else (labels[i] if isinstance(labels[i], list) else [labels[i]])
        for lab in group:
            lab = np.float32(lab)
            y[:, i][np.abs(x[:, 0] - lab) <= 1e-8 + 1e-5 * np.abs(lab)] = 1        # one_hot.py:40-43 (torch.isclose)
    return y


def label_map_from_one_hot(p: np.ndarray, labels, threshold: float = 0.5, sum_then_threshold: bool = False,
                           label_hierarchy: bool = False) -> np.ndarray:
    """unet3d/utils/one_hot.py:46-118 on a channel-first array (L, ...) -> int16 label map."""
    p = np.asarray(p, dtype=np.float32)
    if label_hierarchy:                                                            # one_hot.py:92-110
        roi = np.ones(p.shape[1:], dtype=bool)
        out = np.zeros(p.shape[1:], dtype=np.int16)
        for i, lab in enumerate(labels):
            roi = np.logical_and(p[i] > threshold, roi)
            out[roi] = lab
        return out
    if all(isinstance(g, list) for g in labels):                                   # one_hot.py:53-62
        maps, i = [], 0
        for g in labels:
            maps.append(label_map_from_one_hot(p[i:i + len(g)], g, threshold, sum_then_threshold))
            i += len(g)
        return np.stack(maps, axis=0)
    n = len(labels)                                                                # one_hot.py:70-89
    mask = (p[:n].sum(axis=0) > threshold) if sum_then_threshold else (p[:n] > threshold).any(axis=0)
    arg = np.argmax(p[:n], axis=0) + 1
    arg = np.where(mask, arg, 0)
    out = np.zeros(p.shape[1:], dtype=np.int16)
    for i, lab in enumerate(labels):
        out[arg == i + 1] = lab
    return out


def normalize_intensity(img: np.ndarray, nonzero: bool = False, channel_wise: bool = False) -> np.ndarray:
    """monai.transforms.NormalizeIntensity (selected by unet3d/datasets/segmentation.py:77-87) -- parity unpinned."""
    x = np.asarray(img, dtype=np.float64)
    out = x.copy()

    def norm(a):
        sel = (a != 0) if nonzero else np.ones(a.shape, dtype=bool)
        if not sel.any():
            return a
        m = a[sel].mean()
        s = a[sel].std()
        if s == 0:
            s = 1.0
        r = a.copy()
        r[sel] = (a[sel] - m) / s
        return r
    if channel_wise:
        for c in range(x.shape[0]):
            out[c] = norm(x[c])
    else:
        out = norm(x)
    return out.astype(np.float32)

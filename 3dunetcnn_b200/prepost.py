"""The steps right before and right after the U-Net path, on the device.

Mirrors (paths relative to /root/reference):
  * ``unet3d/utils/one_hot.py:7-37``   ``compile_one_hot_encoding``  label map -> one-hot uint8 target
  * ``unet3d/utils/one_hot.py:46-118`` ``convert_one_hot_to_label_map`` (+ hierarchy)  prediction -> label map
  * ``unet3d/datasets/segmentation.py:77-87``  ``normalization="zero_mean"`` -> ``monai.transforms.NormalizeIntensity``

Same names, argument meaning and error behaviour; tensors live on the GPU and every voxel is touched by one
hand-written kernel of libb200unet instead of a chain of boolean-mask torch ops on the host.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import lib as _lib


def _plain(t: torch.Tensor) -> torch.Tensor:
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s runs only on CUDA tensors (no CPU fallback)" % what)


def compile_one_hot_encoding(data, n_labels, labels=None, dtype=torch.uint8, return_4d=True, round=True):
    """one_hot.py:7-37.  ``data``: label map (n_samples, 1, ...) -- lower-rank inputs gain leading axes; ``labels``:
    label values, a nested list groups several values into one channel; default ``1..n_labels``."""
    _need_cuda(data, "compile_one_hot_encoding")
    x = _plain(data)
    while x.dim() < 5:
        x = x[None]
    assert x.shape[1] == 1
    if dtype != torch.uint8:
        raise NotImplementedError("compile_one_hot_encoding: only dtype=torch.uint8 (the reference's default) is implemented")
    values, begin = [], [0]
    for i in range(n_labels):
        if labels is not None:
            group = labels[i] if type(labels[i]) == list else [labels[i]]
        else:
            group = [i + 1]
        values.extend(float(v) for v in group)
        begin.append(len(values))
    x = x.contiguous().float()
    n = x.shape[0]
    spatial = x[0, 0].numel()
    y = torch.empty((n, n_labels) + tuple(x.shape[2:]), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load_library().b200unet_one_hot(x.data_ptr(), n, spatial, (C.c_float * len(values))(*values),
                                                        (C.c_int32 * len(begin))(*begin), n_labels, int(bool(round)),
                                                        y.data_ptr(), _lib.stream_ptr()), "one_hot")
    if return_4d:
        assert y.shape[0] == 1
        y = y[0]
    return y


def normalize_intensity(img: torch.Tensor, nonzero: bool = False, channel_wise: bool = False,
                        subtrahend=None, divisor=None) -> torch.Tensor:
    """``monai.transforms.NormalizeIntensity`` on one channel-first image (C, ...): ``(img - mean) / std`` (population
    std; ``std == 0`` -> 1), per channel when ``channel_wise``; ``nonzero``: statistics over, and changes to, the
    non-zero voxels only.  Parity unpinned (MONAI absent): restated from its documented behaviour."""
    _need_cuda(img, "normalize_intensity")
    if subtrahend is not None or divisor is not None:
        raise NotImplementedError("normalize_intensity: explicit subtrahend/divisor are not implemented")
    x = _plain(img).contiguous().float()
    groups = x.shape[0] if channel_wise else 1
    spatial = x.numel() // groups
    y = torch.empty_like(x)
    stats = torch.empty((groups, 3), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load_library().b200unet_zscore(x.data_ptr(), groups, spatial, int(bool(nonzero)), stats.data_ptr(),
                                                       y.data_ptr(), _lib.stream_ptr()), "zscore")
    return y


def _label_map(p: torch.Tensor, labels: Sequence[int], act: int, threshold: float, hierarchy: bool, sum_then_threshold: bool):
    x = p.contiguous().float()
    L = len(labels)
    if x.shape[0] < L:
        raise ValueError("one-hot encoding has %d channels but %d labels were given" % (x.shape[0], L))
    spatial = x[0].numel()
    out = torch.empty(tuple(x.shape[1:]), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load_library().b200unet_label_map(x.data_ptr(), L, spatial, (C.c_int32 * L)(*[int(v) for v in labels]), act,
                                                          float(threshold), int(bool(hierarchy)), int(bool(sum_then_threshold)),
                                                          out.data_ptr(), _lib.stream_ptr()), "label_map")
    return out


_ACT = {None: 0, "sigmoid": 1, "softmax": 2}


def convert_one_hot_to_label_map(one_hot_encoding, labels, axis=0, threshold=0.5, sum_then_threshold=False,
                                 dtype=torch.int16, label_hierarchy=False, activation: Optional[str] = None):
    """one_hot.py:46-67 on a channel-first prediction (L, ...).  ``activation`` (extension): apply sigmoid / softmax
    to logits inside the same kernel (volumetric.py:151-156 runs it as a separate pass)."""
    _need_cuda(one_hot_encoding, "convert_one_hot_to_label_map")
    if axis != 0:
        raise NotImplementedError("convert_one_hot_to_label_map: only axis=0 (channel first) is implemented")
    if dtype != torch.int16:
        raise NotImplementedError("convert_one_hot_to_label_map: only dtype=torch.int16 (the reference's default)")
    if activation not in _ACT:
        raise ValueError("activation must be None, 'sigmoid' or 'softmax'")
    x = _plain(one_hot_encoding)
    if label_hierarchy:
        return _label_map(x, labels, _ACT[activation], threshold, True, False)
    if all(type(_labels) == list for _labels in labels):
        # several label-map volumes, one per group of channels (one_hot.py:53-62); the activation spans all channels, so
        # it is applied per group only when it is channel-independent
        if activation == "softmax":
            raise NotImplementedError("softmax across grouped label maps: apply it before the call")
        maps, i = [], 0
        for _labels in labels:
            maps.append(_label_map(x[i:i + len(_labels)], _labels, _ACT[activation], threshold, False, sum_then_threshold))
            i += len(_labels)
        return torch.stack(maps, dim=axis)
    return _label_map(x, labels, _ACT[activation], threshold, False, sum_then_threshold)


def convert_one_hot_to_label_map_using_hierarchy(one_hot_encoding, labels, threshold=0.5, axis=0, dtype=torch.int16):
    """one_hot.py:92-110."""
    return convert_one_hot_to_label_map(one_hot_encoding, labels, axis=axis, threshold=threshold, dtype=dtype,
                                        label_hierarchy=True)

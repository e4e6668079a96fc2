"""Inference mirror: sliding-window inferer + ``volumetric_predictions`` calling contract.

Reference call sites: /root/reference/unet3d/predict/volumetric.py:131-177 (no_grad loop, ``inferer(x, model)`` or
``model(x)``, activation), unet3d/scripts/script_utils.py:290-293 (``monai.inferers.<name>(**kw)``).  MONAI and
nibabel are absent from this image, so the inferer is restated here (tiles gathered / accumulated on the device)
and file output goes through a caller-supplied ``writer`` instead of ``monai.data.NibabelWriter``.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch


def _scan_starts(size: int, roi: int, overlap: float) -> List[int]:
    interval = max(int(roi * (1 - overlap)), 1)
    if size <= roi:
        return [0]
    n = int(math.ceil(float(size - roi) / interval)) + 1
    return [min(i * interval, size - roi) for i in range(n)]


def _gaussian_importance(roi: Sequence[int], device, sigma_scale: float = 0.125) -> torch.Tensor:
    w = None
    for r in roi:
        c = (r - 1) / 2.0
        s = sigma_scale * r
        g = torch.exp(-0.5 * ((torch.arange(r, dtype=torch.float32, device=device) - c) / s) ** 2)
        w = g if w is None else w[..., None] * g
    w = w / w.max()
    return torch.clamp(w, min=float(w[w > 0].min()))


class SlidingWindowInferer:
    """``monai.inferers.SlidingWindowInferer(roi_size, sw_batch_size, overlap, mode)`` restated (parity unpinned)."""

    def __init__(self, roi_size, sw_batch_size: int = 1, overlap: float = 0.25, mode: str = "constant", **unused):
        self.roi_size = tuple(int(r) for r in (roi_size if hasattr(roi_size, "__len__") else (roi_size,) * 3))
        self.sw_batch_size = int(sw_batch_size)
        self.overlap = float(overlap)
        if mode not in ("constant", "gaussian"):
            raise ValueError("mode must be 'constant' or 'gaussian'")
        self.mode = mode

    def __call__(self, inputs: torch.Tensor, network: Callable, *args, **kwargs) -> torch.Tensor:
        n, _, D, H, W = inputs.shape
        roi = [min(r, s) for r, s in zip(self.roi_size, (D, H, W))]
        dev = inputs.device
        w = torch.ones(roi, device=dev) if self.mode == "constant" else _gaussian_importance(roi, dev)
        starts = [(d, h, x) for d in _scan_starts(D, roi[0], self.overlap) for h in _scan_starts(H, roi[1], self.overlap)
                  for x in _scan_starts(W, roi[2], self.overlap)]
        out = None
        cnt = torch.zeros((1, 1, D, H, W), dtype=torch.float32, device=dev)
        # tiles of all batch items are grouped so the plan sees a constant batch of sw_batch_size
        jobs = [(b, s) for b in range(n) for s in starts]
        for j0 in range(0, len(jobs), self.sw_batch_size):
            chunk = jobs[j0:j0 + self.sw_batch_size]
            patch = torch.stack([inputs[b, :, d:d + roi[0], h:h + roi[1], x:x + roi[2]] for b, (d, h, x) in chunk])
            pad = self.sw_batch_size - len(chunk)
            if pad:
                patch = torch.cat([patch, patch[-1:].expand(pad, -1, -1, -1, -1)])
            pred = network(patch.contiguous(), *args, **kwargs)
            if out is None:
                out = torch.zeros((n, pred.shape[1], D, H, W), dtype=torch.float32, device=dev)
            for k, (b, (d, h, x)) in enumerate(chunk):
                out[b, :, d:d + roi[0], h:h + roi[1], x:x + roi[2]] += pred[k].float() * w
                if b == 0:
                    cnt[0, :, d:d + roi[0], h:h + roi[1], x:x + roi[2]] += w
        return out / cnt


_INFERERS = {"SlidingWindowInferer": SlidingWindowInferer}


def build_inferer_from_config(config):
    """script_utils.py:290-293."""
    kw = dict(config)
    name = kw.pop("name")
    if name not in _INFERERS:
        raise ValueError("inferer {} not supported".format(name))
    return _INFERERS[name](**kw)


def _filename_of(x, index: int):
    """volumetric.py:11-51 semantics: inputs must carry ``meta['filename_or_obj']``."""
    meta = getattr(x, "meta", None)
    if meta is None:
        raise TypeError("volumetric_predictions requires inputs with a 'meta' attribute (MetaTensor-like)")
    if "filename_or_obj" not in meta:
        raise KeyError("filename_or_obj")
    f = meta["filename_or_obj"]
    return f[index] if isinstance(f, (list, tuple)) else f


def volumetric_predictions(model, dataloader, prediction_dir, activation=None, resample=False, interpolation="trilinear",
                           inferer=None, writer: Optional[Callable] = None):
    """volumetric.py:131-177 without the NIfTI side: returns [(filename, prediction tensor)] and calls
    ``writer(filename, tensor, prediction_dir)`` per item when given."""
    results = []
    with torch.no_grad():
        for item in dataloader:
            x = item["image"]
            filenames = [_filename_of(x, i) for i in range(x.shape[0])]
            xd = x.to(next(model.parameters()).device)
            predictions = inferer(xd, model) if inferer is not None else model(xd)
            if activation == "sigmoid":
                predictions = torch.sigmoid(predictions)
            elif activation == "softmax":
                predictions = torch.softmax(predictions, dim=1)
            elif activation is not None:
                predictions = getattr(torch, activation)(predictions)
            for i, fn in enumerate(filenames):
                results.append((fn, predictions[i]))
                if writer is not None:
                    writer(fn, predictions[i], prediction_dir)
    return results

"""Import the UNMODIFIED reference ``UNet3D`` from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call this; it is used by ``tests/golden/make_golden.py`` (fixture generation), by the
``-m "not gpu"`` tests that pin the restatement when the reference is present, and by
``bench.py --impl reference`` when it happens to run in this container.

Two shims are required:
  1. ``monai`` is not installed and ``unet3d/models/pytorch/__init__.py:1`` star-imports
     ``monai.networks.nets`` -> empty stub modules are placed in ``sys.modules``.
  2. ``unet3d/models/pytorch/segmentation/unet.py:38`` calls ``F.pad`` without importing ``F``
     -> ``torch.nn.functional`` is injected as the module attribute ``F``.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200UNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "unet3d", "models"))


def load_reference_unet_module():
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch
    for name in ("monai", "monai.networks", "monai.networks.nets"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__all__ = []
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import unet3d.models.pytorch.segmentation.unet as ref_unet  # noqa: E402
    ref_unet.F = torch.nn.functional
    return ref_unet


def reference_unet3d(**kwargs):
    """Construct the reference ``UNet3D`` (its ctor prints widths; silenced)."""
    ref_unet = load_reference_unet_module()
    with contextlib.redirect_stdout(io.StringIO()):
        return ref_unet.UNet3D(**kwargs)

"""CPU oracle for the 3D U-Net hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``3dunetcnn_b200``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or
the timed CPU baseline -- never as the thing shipped.

Parity status: PINNED against the reference's own ``UNet3D`` class (imported
in the build container through ``oracle/ref_loader.py``) by
``tests/golden/make_golden.py``; the resulting fixtures are committed under
``tests/golden/``.  ``prepost_oracle`` (label map <-> one-hot) is PINNED the same
way against the reference's own ``unet3d/utils/one_hot.py``
(``tests/golden/make_golden_prepost.py`` -> ``tests/golden/prepost.npz``).
The Dice and sliding-window restatements follow MONAI's public definitions;
MONAI itself is absent from this image, so they are held to known-answer cases
of MONAI's own unit tests quoted from memory (``tests/test_oracle_monai_cases.py``:
four Dice values to 6 digits, the ``compute = data + 1`` identity of the
sliding-window inferer) -- pinned to published values, not to an importable
MONAI.  ``dynunet_oracle`` and ``prepost_oracle.normalize_intensity`` restate
MONAI code that no fixture here can reach: PARITY UNPINNED.
"""
from .unet3d_oracle import (  # noqa: F401
    UNetConfig,
    unet3d_forward,
    unet3d_state_dict_spec,
    make_state_dict,
    dice_loss,
    dice_loss_grad,
    sliding_window_inference,
    trilinear_upsample2x,
    group_norm,
    conv3d_direct,
)
from . import prepost_oracle, dynunet_oracle  # noqa: F401,E402

"""Generate tests/golden/prepost.npz by running the UNMODIFIED reference ``unet3d/utils/one_hot.py`` on CPU.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_prepost.py
``monai.data.MetaTensor`` (the only monai symbol that module needs) is stubbed by a torch.Tensor subclass carrying
``.meta``; scipy is installed.  Inputs are regenerated from seeds by ``prepost_inputs`` and are not stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from recipe_prepost import ONE_HOT_CASES, LABEL_MAP_CASES, label_map_input, prediction_input  # noqa: E402


def load_reference_one_hot():
    class MetaTensor(torch.Tensor):
        meta = None          # results of torch ops on a MetaTensor keep the subclass but not instance attributes

        @staticmethod
        def __new__(cls, x, meta=None, *a, **k):
            t = torch.as_tensor(x).as_subclass(cls)
            t.meta = meta
            return t

    for name in ("monai", "monai.data"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["monai.data"].MetaTensor = MetaTensor
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_one_hot", "/root/reference/unet3d/utils/one_hot.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, MetaTensor


def main():
    ref, MetaTensor = load_reference_one_hot()
    out = {}
    for name, (shape, values, n_labels, labels, seed) in ONE_HOT_CASES.items():
        data = MetaTensor(label_map_input(shape, values, seed), meta={})
        y = ref.compile_one_hot_encoding(data, n_labels=n_labels, labels=labels, return_4d=False)
        out["one_hot::" + name] = np.packbits(torch.as_tensor(y).numpy().astype(np.uint8))
        out["one_hot_shape::" + name] = np.array(y.shape)
    for name, (shape, labels, kw, seed) in LABEL_MAP_CASES.items():
        p = prediction_input(shape, seed)
        lm = ref.convert_one_hot_to_label_map(p, labels=labels, **kw)
        out["label_map::" + name] = torch.as_tensor(lm).numpy().astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "prepost.npz"), **out)
    print("wrote prepost.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()

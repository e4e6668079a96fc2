#!/usr/bin/env python
"""Measure, per parameter tensor, the gradient error of (a) this library in bf16 mode, (b) this library in split mode
and (c) the oracle graph under torch.autocast(bf16) on the same GPU, all against the oracle graph in fp64.
Output: one JSON line per configuration (gpurun_out/bf16_grad_study.jsonl).  The test
tests/test_gpu_bf16_vs_autocast.py asserts the bound this measurement supports."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import UNetConfig, make_state_dict, unet3d_forward, dice_loss  # noqa: E402

pkg = importlib.import_module("3dunetcnn_b200")
DEV = "cuda"


def study(kw, shape, seed):
    cfg = UNetConfig(**kw)
    sd = make_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    t = (torch.rand((shape[0], cfg.n_outputs) + tuple(shape[2:]), generator=g) > 0.7).to(torch.uint8)
    xd, td = x.to(DEV), t.to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def oracle_run(dtype, autocast):
        sdr = {k: v.to(DEV, dtype).requires_grad_(True) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = unet3d_forward(sdr, xd.to(dtype), cfg)
        loss = dice_loss(out.float() if autocast else out, td)
        loss.backward()
        return out.detach().double(), float(loss), {k: v.grad.double() for k, v in sdr.items()}

    ref_out, ref_loss, ref_g = oracle_run(torch.float64, False)
    ac_out, ac_loss, ac_g = oracle_run(torch.float32, True)
    f32_out, f32_loss, f32_g = oracle_run(torch.float32, False)

    def ours(precision):
        m = pkg.UNet3D(precision=precision, **kw).to(DEV)
        m.load_state_dict(sd, strict=True)
        m.train()
        m.set_dropout_scale(torch.ones(shape[0], cfg.base_width))
        out = m(xd)
        loss = pkg.DiceLoss(sigmoid=True)(out, td)
        loss.backward()
        torch.cuda.synchronize()
        return out.detach().double(), float(loss), {k: p.grad.double() for k, p in m.named_parameters()}

    res = {"config": kw, "shape": list(shape)}
    runs = {"autocast_bf16": (ac_out, ac_loss, ac_g), "cudnn_fp32": (f32_out, f32_loss, f32_g), "ours_bf16": ours("bf16"),
            "ours_split": ours("split")}
    for name, (out, loss, gr) in runs.items():
        per = {}
        num = den = 0.0
        for k in ref_g:
            d = (gr[k] - ref_g[k])
            per[k] = float(d.norm() / (ref_g[k].norm() + 1e-30))
            num += float((d * d).sum())
            den += float((ref_g[k] * ref_g[k]).sum())
        norm_ratio = {k: float(gr[k].norm() / (ref_g[k].norm() + 1e-30)) for k in ref_g}
        res[name] = {"logits_rel": float((out - ref_out).norm() / ref_out.norm()), "ddice": abs(loss - ref_loss),
                     "grad_rel_whole": (num / den) ** 0.5, "grad_rel_worst": max(per.values()),
                     "grad_rel_median": sorted(per.values())[len(per) // 2], "per_tensor": per,
                     "norm_ratio_min": min(norm_ratio.values()), "norm_ratio_max": max(norm_ratio.values())}
    ratio = {k: res["ours_bf16"]["per_tensor"][k] / (res["autocast_bf16"]["per_tensor"][k] + 1e-30) for k in ref_g}
    res["ratio_ours_over_autocast"] = {"max": max(ratio.values()), "median": sorted(ratio.values())[len(ratio) // 2],
                                       "argmax": max(ratio, key=ratio.get)}
    return res


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    cases = [(dict(n_features=4, n_outputs=3, base_width=8), (1, 4, 32, 32, 32), 1),
             (dict(n_features=4, n_outputs=3, base_width=16), (1, 4, 64, 64, 64), 2),
             (dict(n_features=4, n_outputs=3, base_width=16), (1, 4, 64, 64, 64), 3),
             (dict(n_features=4, n_outputs=3, base_width=32), (1, 4, 64, 64, 64), 4)]
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        cases.append((dict(n_features=4, n_outputs=3, base_width=32), (1, 4, 128, 128, 128), 5))
    with open(os.path.join(ROOT, "gpurun_out", "bf16_grad_study.jsonl"), "w") as f:
        for kw, shape, seed in cases:
            r = study(kw, shape, seed)
            slim = {k: ({kk: vv for kk, vv in v.items() if kk != "per_tensor"} if isinstance(v, dict) else v) for k, v in r.items()}
            print(json.dumps(slim), flush=True)
            f.write(json.dumps(r) + "\n")

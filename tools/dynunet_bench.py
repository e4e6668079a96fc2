"""Throughput of one training step (forward + sigmoid-Dice + backward) of the DynUNet the reference's example config trains
(examples/brats2020/brats2020_config.json:2-107: 4 -> 3 channels, filters 64/96/128/192/256/384, 128^3 patches), batch 2,
bf16 mode, eager launches and CUDA-graph replay.  Prints one JSON line.  Not a BASELINE.json config: recorded for DESIGN.md."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
L = 6
kw = dict(spatial_dims=3, in_channels=4, out_channels=3, kernel_size=[[3, 3, 3]] * L, strides=[[1, 1, 1]] + [[2, 2, 2]] * (L - 1),
          upsample_kernel_size=[[2, 2, 2]] * (L - 1), filters=[64, 96, 128, 192, 256, 384])
torch.manual_seed(0)
model = pkg.DynUNet(**kw).cuda().train()
crit = pkg.DiceLoss(sigmoid=True)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
x = torch.randn(2, 4, 128, 128, 128, device="cuda")
t = (torch.rand(2, 3, 128, 128, 128, device="cuda") > 0.7).to(torch.uint8)


def eager():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    opt.step()
    return loss


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        loss = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, float(loss)


ms_e, loss_e = timeit(eager)
plan = model._plan_for(x)
macs = plan.algorithmic_macs()
flops = 2.0 * sum(macs.values())
out = {"model": "DynUNet brats2020 config", "batch": 2, "eager_ms_per_step": ms_e, "eager_volumes_per_s": 2 / (ms_e / 1e3), "loss": loss_e,
       "algorithmic_tflop_per_step": flops / 1e12, "tflops_eager": flops / (ms_e / 1e3) / 1e12,
       "launches": model.launches_last_forward + model.launches_last_backward, "workspace_gb": plan.ws_bytes / 2 ** 30}
try:
    step = pkg.train.GraphedTrainStep(model, crit, opt, x.shape, t.shape)
    ms_g, loss_g = timeit(lambda: step(x, t))
    out.update(graph_ms_per_step=ms_g, graph_volumes_per_s=2 / (ms_g / 1e3), tflops_graph=flops / (ms_g / 1e3) / 1e12)
except Exception as e:  # noqa: BLE001
    out["graph_error"] = repr(e)[:200]
print(json.dumps(out))

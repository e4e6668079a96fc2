"""Per-launch timing of one warm C2 training step (CUDA events around every launch of the plan) -> CSV + summary."""
import collections
import csv
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "layer_times.csv")
torch.manual_seed(0)
model = pkg.UNet3D(n_features=4, n_outputs=3, base_width=32).cuda()
crit = pkg.DiceLoss(sigmoid=True)
x = torch.randn(2, 4, 128, 128, 128, device="cuda")
t = (torch.rand(2, 3, 128, 128, 128, device="cuda") > 0.7).to(torch.uint8)
model.train()


def step():
    model.zero_grad(set_to_none=True)
    crit(model(x), t).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
plan = model._plan_for(x)
plan.profile_begin(1000)
step()
torch.cuda.synchronize()
plan.profile_dump(out)
plan.profile_end()
rows = list(csv.DictReader(open(out)))
tot = sum(float(r["ms"]) for r in rows)
print("total %.3f ms over %d launches" % (tot, len(rows)))
for r in sorted(rows, key=lambda r: -float(r["ms"]))[:45]:
    print("%8.3f ms  %s" % (float(r["ms"]), r["label"]))
agg = collections.defaultdict(float)
for r in rows:
    agg[r["label"].split(" ")[0]] += float(r["ms"])
print({k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})

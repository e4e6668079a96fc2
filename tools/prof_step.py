"""Two training steps of the C2 workload (for ncu: first step warms up, second is the one to look at)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
size = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
model = pkg.UNet3D(n_features=4, n_outputs=3, base_width=bw).cuda()
crit = pkg.DiceLoss(sigmoid=True)
x = torch.randn(2, 4, size, size, size, device="cuda")
t = (torch.rand(2, 3, size, size, size, device="cuda") > 0.7).to(torch.uint8)
model.train()
for _ in range(steps):
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "launches", model.launches_last_forward, model.launches_last_backward)

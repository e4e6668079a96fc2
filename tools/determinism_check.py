"""Run the C2 backward twice on identical inputs and report per-parameter relative differences (race detector)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
torch.manual_seed(0)
r = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = pkg.UNet3D(n_features=4, n_outputs=3, base_width=32, precision=os.environ.get("PREC", "bf16")).cuda()
x = torch.randn(2, 4, r, r, r, device="cuda")
model.train()
model.set_dropout_scale(torch.ones(2, 32))
g = torch.randn(2, 3, r, r, r, device="cuda") * 1e-6
outs = []
for rep in range(3):
    model.zero_grad(set_to_none=True)
    y = model(x)
    y.backward(g * (2.0 if rep == 2 else 1.0))
    outs.append((y.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()}))
print("logits run0 vs run1 rel", float((outs[0][0] - outs[1][0]).norm() / outs[0][0].norm()))
rows = []
for k in outs[0][1]:
    a, b, c = outs[0][1][k], outs[1][1][k], outs[2][1][k]
    rows.append((float((a - b).norm() / (a.norm() + 1e-30)), float((2 * a - c).norm() / (c.norm() + 1e-30)), float(a.norm()), k))
rows.sort(reverse=True)
print("rel(run0,run1)  rel(2*run0,run2)  |g|  name")
for r_ in rows[:14]:
    print("%.3e  %.3e  %.3e  %s" % r_)
print("... smallest:")
for r_ in rows[-4:]:
    print("%.3e  %.3e  %.3e  %s" % r_)

"""clock64 timeline of CTA 0 of the halo-resident conv kernel (producer / MMA / epilogue roles), first 32 tiles."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ci, co, r = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 32, 128))]
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
dbg = torch.zeros(3 * 32 * 4 + 64, dtype=torch.int64, device="cuda")
os.environ["B200UNET_HALO_DBG"] = str(dbg.data_ptr())
pkg = importlib.import_module("3dunetcnn_b200")
L = pkg.lib
n = 2
x = L.Act.empty(n, r, r, r, ci); x.hi.normal_()
w = torch.randn(co, ci, 3, 3, 3, device="cuda") / (ci * 27) ** 0.5
whi, wlo, cop, cip, _ = L.pack_weights(w, 0)
y = L.Act.empty(n, r, r, r, co)
stats = torch.zeros(n, co, 2, dtype=torch.float64, device="cuda")
res = L.Act.empty(n, r, r, r, co); res.hi.normal_()
coef = torch.rand(n, co, 4, device="cuda"); bst = torch.zeros(n, co, 2, dtype=torch.float64, device="cuda")
kw = dict(stats=stats, stats_ld=co)
if mode == "res":
    kw["res"] = res
if mode == "mode1":
    kw = dict(mode=1, gn_x=res, coef=coef, coef_ld=co, bstats=bst)
for _ in range(3):
    L.conv3d(x, whi, wlo, 3, 1, y, cop, cip, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ.pop("B200UNET_HALO_DBG")          # time it without the stamps
e0.record()
for _ in range(5):
    L.conv3d(x, whi, wlo, 3, 1, y, cop, cip, **kw)
e1.record()
torch.cuda.synchronize()
print("kernel time without stamps: %.4f ms per launch" % (e0.elapsed_time(e1) / 5))
ep = dbg.cpu()[384:]
d = dbg.cpu()[:384].view(3, 32, 4)
t0 = int(d[0, 0, 0])
print("shape ci%d co%d r%d mode %s; all times in cycles relative to the producer's first stamp" % (ci, co, r, mode))
print("tile | prod: start, halo_empty_ok | mma: start, acc_empty_ok, halo_full_ok, committed | epi: start, acc_full_ok, drained, flushed")
for t in range(12):
    f = lambda v: "%7d" % (int(v) - t0) if int(v) else "      -"
    print("%4d | %s %s | %s %s %s %s | %s %s %s %s" % ((t,) + tuple(f(v) for v in list(d[0, t, :2]) + list(d[1, t]) + list(d[2, t]))))
print("MMA commit-to-commit deltas of tiles 12..31:", [int(d[1, t, 3]) - int(d[1, t - 1, 3]) for t in range(12, 32) if int(d[1, t, 3])])
per = (int(d[1, 11, 3]) - int(d[1, 3, 3])) / 8.0
print("steady-state cycles per tile (MMA commit to commit): %.0f ; epilogue drain time per tile: %.0f" % (per, float((d[2, 3:11, 2] - d[2, 3:11, 1]).double().mean())))

print("epilogue of tile 4, per plane: deltas [chunks staged, fence done, wait_read done, barrier done, store issued]")
for dpl in range(4):
    v = [int(ep[dpl * 8 + i]) for i in (0, 2, 3, 4, 5, 6)]
    if v[0]:
        print("  plane %d:" % dpl, [v[i + 1] - v[i] for i in range(5)], " total", v[5] - v[0])
top = [int(ep[40 + i]) for i in range(6)]
if top[0]:
    print("epilogue warp 2, tile 4: [decode+coef, load_side issue, prefetch issue, wait acc_full, drain] =", [top[i + 1] - top[i] for i in range(5)])

"""Race detector: run every conv kernel variant at the C2 layer shapes several times on identical inputs and compare
the outputs BITWISE (the activation outputs involve no atomics, so any difference is a synchronisation bug); the fp64
statistics and fp32 wgrad atomics are compared with a tolerance."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
L = pkg.lib
DEV = "cuda"
SHAPES = [(8, 32, 128), (32, 32, 128), (64, 32, 128), (32, 64, 128), (32, 64, 64), (64, 64, 64), (128, 128, 64), (64, 128, 32),
          (128, 128, 32), (256, 256, 32), (128, 256, 16), (256, 256, 16)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
split = len(sys.argv) > 2 and sys.argv[2] == "split"
torch.manual_seed(1)
bad = 0
for (ci, co, r) in SHAPES:
    n = 2
    x = L.Act.empty(n, r, r, r, ci, split=split) if split else L.Act.empty(n, r, r, r, ci)
    x.hi.normal_()
    if split: x.lo.normal_(); x.lo.mul_(1e-3)
    w = torch.randn(co, ci, 3, 3, 3, device=DEV) / (ci * 27) ** 0.5
    whi, wlo, cop, cip, _ = L.pack_weights(w, 0)
    res = L.Act.empty(n, r, r, r, co, split=split) if split else L.Act.empty(n, r, r, r, co)
    res.hi.normal_()
    coef = torch.rand(n, co, 4, device=DEV)
    for mode in ("plain", "res", "mode1"):
        outs = []
        for rep in range(reps):
            y = L.Act.empty(n, r, r, r, co, split=split) if split else L.Act.empty(n, r, r, r, co)
            y.hi.fill_(7.0)
            stats = torch.zeros(n, co, 2, dtype=torch.float64, device=DEV)
            if mode == "plain":
                L.conv3d(x, whi, wlo if split else None, 3, 1, y, cop, cip, stats=stats, stats_ld=co)
            elif mode == "res":
                L.conv3d(x, whi, wlo if split else None, 3, 1, y, cop, cip, res=res, stats=stats, stats_ld=co)
            else:
                L.conv3d(x, whi, wlo if split else None, 3, 1, y, cop, cip, mode=1, gn_x=res, coef=coef, coef_ld=co, bstats=stats)
            torch.cuda.synchronize()
            outs.append((y.hi.clone(), stats.clone()))
        nd = sum(int((outs[0][0] != o[0]).sum()) for o in outs[1:])
        sd = max(float((outs[0][1] - o[1]).abs().max() / (outs[0][1].abs().max() + 1e-30)) for o in outs[1:])
        flag = "RACE" if nd or sd > 1e-9 else "ok"
        bad += flag != "ok"
        print("[det] conv ci%-3d co%-3d r%-3d %-5s: differing output elements %d, stats rel diff %.2e  %s" % (ci, co, r, mode, nd, sd, flag), flush=True)
    dy = L.Act.empty(n, r, r, r, co); dy.hi.normal_()
    outs = []
    for rep in range(reps):
        dw = torch.zeros(27, cip, cop, device=DEV)
        L.conv3d_wgrad(x, dy, 3, 1, cip, cop, dw)
        torch.cuda.synchronize()
        outs.append(dw.clone())
    wd = max(float((outs[0] - o).norm() / outs[0].norm()) for o in outs[1:])
    flag = "RACE" if wd > 1e-5 else "ok"
    bad += flag != "ok"
    print("[det] wgrad ci%-3d co%-3d r%-3d: rel diff %.2e %s" % (ci, co, r, wd, flag), flush=True)
print("[det] bad", bad)

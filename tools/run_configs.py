"""Run BASELINE.json configs[2] (C3: 4ch 160x192x128, batch 2, fwd+bwd) and configs[4] (C5: 1ch 256^3 5-level width-48
sliding-window inference) once each on cuda:0 and print wall/GPU times + sanity numbers (finite loss, output norm)."""
import importlib, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
torch.manual_seed(0)
out = {}

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

# ---- C3
model = pkg.UNet3D(n_features=4, n_outputs=3, base_width=32).cuda().train()
crit = pkg.DiceLoss(sigmoid=True)
x = torch.randn(2, 4, 160, 192, 128, device="cuda")
t = (torch.rand(2, 3, 160, 192, 128, device="cuda") > 0.7).to(torch.uint8)
def step():
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    return loss
ms, loss = timed(step, 3)
gn = sum(float(p.grad.double().pow(2).sum()) for p in model.parameters()) ** 0.5
out["C3_4ch_160x192x128_b2_fwd_bwd"] = {"ms_per_step": round(ms, 2), "volumes_per_s": round(2 / (ms / 1e3), 1), "loss": float(loss),
                                        "grad_norm": gn, "finite": bool(torch.isfinite(loss)) and gn == gn}
print(json.dumps(out), flush=True)
del model, x, t
torch.cuda.empty_cache()
# ---- C5
model = pkg.UNet3D(n_features=1, n_outputs=1, base_width=48, encoder_blocks=[1, 2, 2, 4, 4]).cuda().eval()
vol = torch.randn(1, 1, 256, 256, 256, device="cuda")
inf = importlib.import_module("3dunetcnn_b200.predict").SlidingWindowInferer(roi_size=(128, 128, 128), overlap=0.25)
def infer():
    with torch.no_grad():
        return inf(vol, model)
ms, y = timed(infer, 1)
out["C5_1ch_256cube_5level_w48_sliding_window"] = {"ms_per_volume": round(ms, 1), "out_shape": list(y.shape),
                                                   "finite": bool(torch.isfinite(y).all()), "out_norm": float(y.double().norm())}
print(json.dumps(out))

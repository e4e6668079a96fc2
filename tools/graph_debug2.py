"""Python-level bisect of the GraphedTrainStep capture: replicas of its body with the capture status of the stream printed
after every statement (cudaStreamIsCapturing: 0 none, 1 active, 2 invalidated).  One variant per process."""
import ctypes
import importlib
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rt = None


def status(tag):
    global rt
    if rt is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
    st = ctypes.c_int(-1)
    e = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    print("   [%s] capture status %d (err %d)" % (tag, st.value, e), flush=True)
    return st.value


def variant(name):
    pkg = importlib.import_module("3dunetcnn_b200")
    torch.manual_seed(0)
    kw = dict(n_features=2, n_outputs=2, base_width=8, encoder_blocks=[1, 1, 1], decoder_blocks=[1, 1, 1])
    model = pkg.UNet3D(precision="bf16", dropout=0.0, **kw).cuda().train()
    crit = pkg.DiceLoss(sigmoid=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = torch.randn(2, 2, 16, 16, 16, device="cuda")
    t = (torch.rand(2, 2, 16, 16, 16, device="cuda") > 0.5).to(torch.uint8)
    images = torch.zeros_like(x)
    target = torch.zeros_like(t)
    images.copy_(x, non_blocking=True)
    target.copy_(t, non_blocking=True)
    model.use_flat_gradients(True)
    dev = next(model.parameters()).device
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            loss = crit(model(images), target)
            loss.backward()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    if name == "del_loss":
        del loss
    if name == "keep_graph_alive":
        keep = loss   # noqa: F841
    model._overwrite_grads = True
    g = torch.cuda.CUDAGraph()
    holder = {}
    with torch.cuda.graph(g):
        status("begin")
        out = model(images)
        status("after forward")
        holder["loss"] = crit(out, target)
        status("after criterion")
        if name == "assign_over_old":
            loss = holder["loss"]          # rebinding frees the warm-up loss (and its autograd graph) INSIDE the capture
            status("after rebinding the warm-up loss")
        holder["loss"].backward()
        status("after backward")
    g.replay()
    torch.cuda.synchronize()
    print("OK   %s loss %.5f" % (name, float(holder["loss"])), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2:
        variant(sys.argv[1])
    else:
        env = dict(os.environ, B200UNET_CAPTURE_DEBUG="1")
        for name in ("replica", "del_loss", "assign_over_old"):
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=600, env=env)
            print("== %s rc %d" % (name, r.returncode))
            print("\n".join(r.stdout.strip().splitlines()[-8:]))
            if r.returncode:
                print(" | ".join([ln for ln in r.stderr.strip().splitlines() if "Error" in ln][-2:])[:300])

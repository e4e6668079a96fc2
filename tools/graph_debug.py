"""Which part of the step survives CUDA-graph capture, and in which capture_error_mode?  Each variant runs in its own process
(a failed capture leaves torch's default generator registered to a dead graph)."""
import importlib
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def variant(name, mode):
    pkg = importlib.import_module("3dunetcnn_b200")
    L = pkg.lib
    torch.manual_seed(0)
    kw = dict(n_features=2, n_outputs=2, base_width=8, encoder_blocks=[1, 1, 1], decoder_blocks=[1, 1, 1])
    model = pkg.UNet3D(precision="bf16", dropout=0.0, **kw).cuda().train()
    model.use_flat_gradients(True)
    crit = pkg.DiceLoss(sigmoid=True)
    x = torch.randn(2, 2, 16, 16, 16, device="cuda")
    t = (torch.rand(2, 2, 16, 16, 16, device="cuda") > 0.5).to(torch.uint8)

    def body():
        if name == "memset":
            x.zero_()
            return
        if name == "conv_only":
            a = L.Act.empty(2, 16, 16, 16, 16)
            w = torch.zeros(27, 16, 16, dtype=torch.bfloat16, device="cuda")
            y = L.Act.empty(2, 16, 16, 16, 16)
            L.conv3d(a, w, None, 3, 1, y, 16, 16)
            return
        if name == "wgrad_only":
            a = L.Act.empty(2, 16, 16, 16, 16)
            dw = torch.zeros(27, 16, 16, device="cuda")
            L.conv3d_wgrad(a, a, 3, 1, 16, 16, dw)
            return
        if name == "forward_nograd":
            with torch.no_grad():
                model(x)
            return
        out = model(x)
        if name == "forward":
            return
        loss = crit(out, t)
        if name == "forward_dice":
            return
        loss.backward()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            model.zero_grad(set_to_none=True)
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    model._overwrite_grads = True
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode=mode):
        body()
    g.replay()
    torch.cuda.synchronize()
    print("OK   %-16s %s" % (name, mode), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        variant(sys.argv[1], sys.argv[2])
    else:
        for mode in ("global", "relaxed"):
            for name in ("memset", "conv_only", "wgrad_only", "forward_nograd", "forward", "forward_dice", "full"):
                r = subprocess.run([sys.executable, __file__, name, mode], capture_output=True, text=True, timeout=300)
                if r.returncode == 0:
                    print(r.stdout.strip().splitlines()[-1])
                else:
                    tail = [ln for ln in (r.stderr or r.stdout).strip().splitlines() if "Error" in ln or "error" in ln][-2:]
                    print("FAIL %-16s %s : %s" % (name, mode, " | ".join(tail)[:300]), flush=True)

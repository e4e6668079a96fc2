"""Which part of the step survives CUDA-graph capture?  Each variant runs in its own process (a failed capture leaves
torch's default generator registered to a dead graph).  B200UNET_CAPTURE_DEBUG=1 makes the library name the op after
which the capturing stream reports cudaStreamCaptureStatusInvalidated."""
import importlib
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def variant(name):
    pkg = importlib.import_module("3dunetcnn_b200")
    torch.manual_seed(0)
    opts = set(name.split("+"))
    kw = dict(n_features=2, n_outputs=2, base_width=8, encoder_blocks=[1, 1, 1], decoder_blocks=[1, 1, 1])
    size = 16
    if "big" in opts:
        kw = dict(n_features=4, n_outputs=3, base_width=32)
        size = 64
    model = pkg.UNet3D(precision="split" if "split" in opts else "bf16", dropout=0.2 if "drop" in opts else 0.0, **kw).cuda().train()
    crit = pkg.DiceLoss(sigmoid=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused="fused" in opts)
    x = torch.randn(2, kw["n_features"], size, size, size, device="cuda")
    t = (torch.rand(2, kw["n_outputs"], size, size, size, device="cuda") > 0.5).to(torch.uint8)
    if "gts" in opts:
        step = pkg.train.GraphedTrainStep(model, crit, opt, x.shape, t.shape)
        xi, ti = (x.cpu().pin_memory(), t.cpu().pin_memory()) if "pinned" in opts else (x, t)
        for _ in range(3):
            loss = step(xi, ti)
        torch.cuda.synchronize()
        print("OK   %-28s loss %.5f" % (name, float(loss)), flush=True)
        return
    model.use_flat_gradients(True)

    def body():
        loss = crit(model(x), t)
        loss.backward()
        return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            (opt if "optzero" in opts else model).zero_grad(set_to_none=True)
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    model._overwrite_grads = True
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = body()
    g.replay()
    torch.cuda.synchronize()
    print("OK   %-28s loss %.5f" % (name, float(loss)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2:
        variant(sys.argv[1])
    else:
        env = dict(os.environ, B200UNET_CAPTURE_DEBUG="1")
        for name in ("raw", "raw+split", "raw+drop", "raw+optzero", "raw+big", "gts", "gts+split", "gts+drop", "gts+pinned", "gts+big+drop+fused"):
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=600, env=env)
            if r.returncode == 0:
                print(r.stdout.strip().splitlines()[-1])
            else:
                tail = [ln for ln in (r.stderr or r.stdout).strip().splitlines() if "Error" in ln or "error" in ln or "invalidated" in ln][-3:]
                print("FAIL %-28s : %s" % (name, " | ".join(tail)[:500]), flush=True)

"""Run under torchrun with >= 2 ranks (one GPU each): the overlapped gradient exchange of train.GraphedTrainStep (two graphs, the
early bucket slice all-reduced on a side stream under the tail of backward) must leave every rank with the same parameters, and
with the parameters of the plain exchange (one graph, one all-reduce after backward) -- bit for bit at 2 ranks, where the
two-operand sum has one order.  Prints one line per rank-0 check and exits non-zero on a mismatch."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
rank, world, local = pkg.parallel.init_process_group_from_env("nccl")
assert world >= 2, "run with torchrun --nproc-per-node >= 2"
dev = torch.device("cuda", local)
KW = dict(n_features=4, n_outputs=3, base_width=16)
SHAPE, TSHAPE = (2, 4, 64, 64, 64), (2, 3, 64, 64, 64)


def run(overlap: bool):
    torch.manual_seed(0)
    model = pkg.UNet3D(precision="bf16", deterministic=True, dropout=0.0, **KW).to(dev)
    crit = pkg.DiceLoss(sigmoid=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    sync = pkg.parallel.GradAllReduce(model.parameters(), model=model)
    sync.broadcast_parameters(0)
    step = pkg.train.GraphedTrainStep(model, crit, opt, SHAPE, TSHAPE, grad_sync=sync, split_backward=overlap)
    losses = []
    for i in range(4):
        g = torch.Generator().manual_seed(1000 * rank + i)          # every rank its own shard
        x = torch.randn(SHAPE, generator=g).to(dev)
        t = (torch.rand(TSHAPE, generator=g) > 0.7).to(torch.uint8).to(dev)
        losses.append(float(step(x, t).item()))
    torch.cuda.synchronize()
    assert (step.graph_tail is not None) == overlap
    flat = torch.cat([p.detach().reshape(-1) for p in model.ordered_parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g_) for g_ in gathered[1:])
    return flat, same, losses


flat_o, same_o, loss_o = run(True)
flat_p, same_p, loss_p = run(False)
diff = float((flat_o.double() - flat_p.double()).norm() / flat_p.double().norm())
ok = same_o and same_p and (torch.equal(flat_o, flat_p) if world == 2 else diff < 1e-6)
if rank == 0:
    print("ddp_check world=%d: ranks agree (overlapped / plain) %s / %s; overlapped vs plain parameters rel-L2 %.3e, bit-equal %s; losses %s"
          % (world, same_o, same_p, diff, torch.equal(flat_o, flat_p), ["%.5f" % v for v in loss_o]), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

#!/usr/bin/env python
"""bench.py -- volumes/sec of the 3D U-Net forward + Dice + backward (+Adam) hot path.

  python bench.py --gpus N --steps K --warmup W            B200 arm (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation of the path

Workload (BASELINE.json configs[1], the config `metric` is quoted on): 4-channel 128^3 volumes, UNet3D
base_width=32 (reference defaults otherwise), bf16 tensor-core operands / fp32 accumulate, batch 2 per GPU,
synthetic data, random-init weights.  One "step" = zero_grad, forward, sigmoid-Dice, backward, [gradient
all-reduce,] fused Adam step.

Prints ONE JSON line on rank 0.  `value`: inputs resident in HBM, CUDA-event timed, max over ranks.  `e2e`: the
same step through the reference-facing API (train.batch_loss) from pinned HOST buffers with a loss.item() read-back
every step.  `roofline`: the halo-resident implicit-GEMM convolution kernel k_conv_halo (the forward and data-gradient
launches it serves; the dominant kernel of the step), algorithmic FLOPs / CUDA-event time summed over its launches,
against the measured bf16 peak; `traffic` = its DRAM bytes per launch from the committed ncu capture.  `cpu_baseline`: the oracle
port of the reference model (torch CPU ops, all host threads) on a bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_KW = dict(n_features=4, n_outputs=3, base_width=32)
VOLUME = (128, 128, 128)
BATCH_PER_GPU = 2
METRIC = "volumes/sec fwd+bwd 4ch 128^3 UNet"
UNIT = "volumes/s"


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# mean DRAM bytes per k_conv_halo launch in one C2 training step (ncu, profiles/r01_final_launches.txt); refresh with
# tools/gpu_trip_prof.sh + tools/summarize_ncu.py when the kernel or the dispatch changes
NCU_TRAFFIC_BYTES_PER_LAUNCH = 358.5e6


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))), "hbm_gbs": float(p.get("hbm_gbs", 6650.0)),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md): 1.59 PFLOP/s, 6.65 TB/s"}


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) >= 9:
                    rows.append(f)
            os.unlink(self.path)
        except Exception:
            pass
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                   reasons=sorted(reasons), samples=len(rows), power_w_max=max(float(r[3]) for r in rows if r[3].replace(".", "").isdigit()))
        return out


def synth_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, MODEL_KW["n_features"]) + VOLUME, generator=g, dtype=torch.float32)
    t = (torch.rand((batch, MODEL_KW["n_outputs"]) + VOLUME, generator=torch.Generator().manual_seed(seed + 1)) > 0.7).to(torch.uint8)
    return x, t


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def cpu_reference_throughput(steps, warmup, budget_s=150.0, batch=1):
    """fwd + Dice + bwd of the oracle port (torch CPU ops == the reference's arithmetic library) on all host threads.
    Each step is a bounded sample: a crop of one 128^3 volume sized so (steps+warmup) steps fit `budget_s`."""
    from oracle import UNetConfig, make_state_dict, unet3d_forward, dice_loss
    from oracle.ref_loader import reference_available
    cfg = UNetConfig(**MODEL_KW)
    kind = "port"
    model = None
    if reference_available():
        try:
            from oracle.ref_loader import reference_unet3d
            model = reference_unet3d(**MODEL_KW)
            model.train()
            model.encoder.layers[0].dropout.p = 0.0
            kind = "reference"
        except Exception:
            model = None
    sd = {k: v.requires_grad_(True) for k, v in make_state_dict(cfg, seed=0).items()}

    def step(shape):
        g = torch.Generator().manual_seed(1)
        x = torch.randn((batch, cfg.n_features) + shape, generator=g)
        t = (torch.rand((batch, cfg.n_outputs) + shape, generator=g) > 0.7).to(torch.uint8)
        t0 = time.perf_counter()
        if model is not None:
            model.zero_grad(set_to_none=True)
            loss = dice_loss(model(x), t)
        else:
            for p in sd.values():
                p.grad = None
            loss = dice_loss(unet3d_forward(sd, x, cfg), t)
        loss.backward()
        return time.perf_counter() - t0

    # pick the thread count that is actually fastest on this host (128-core boxes thrash with one thread per core
    # on these small 3-D convolutions); the chosen count is reported as `cores`
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = None
    for nt in sorted({min(ncpu, 8), 16, 32, 64, ncpu}):      # ascending; stop as soon as more threads get slower
        if nt > ncpu:
            continue
        torch.set_num_threads(nt)
        step((32, 32, 32))
        tt = step((32, 32, 32))
        if best is None or tt < best[0]:
            best = (tt, nt)
        elif tt > 1.2 * best[0]:
            break
    t32, nthreads = best
    torch.set_num_threads(nthreads)
    full = 128 ** 3
    crops = [(128, 128, 128), (128, 128, 64), (128, 64, 64), (64, 64, 64), (64, 64, 32), (64, 32, 32), (32, 32, 32)]
    chosen = crops[-1]
    for c in crops:
        est = t32 * (c[0] * c[1] * c[2]) / 32 ** 3 * 0.8
        if est * (steps + warmup) <= budget_s:
            chosen = c
            break
    frac = chosen[0] * chosen[1] * chosen[2] / full
    for _ in range(warmup):
        step(chosen)
    times = [step(chosen) for _ in range(steps)]
    total = sum(times)
    vps = batch * frac * steps / total
    return {"value": vps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%d step(s) of fwd+Dice+bwd on a %dx%dx%d crop (%.4g of a 128^3 volume), batch %d, fp32, %.1f s/step"
                      % (steps, chosen[0], chosen[1], chosen[2], frac, batch, total / steps)}, total / steps * 1e3


def run_reference_arm(args):
    rank, world, _ = env_rank()
    if rank != 0:
        return
    cb, ms = cpu_reference_throughput(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 4ch 128^3 UNet3D base_width=32 fwd+Dice+bwd on host CPU cores (bounded sample per step)",
                       "sample": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    import torch.distributed as dist
    pkg = importlib.import_module("3dunetcnn_b200")
    rank, world, local = env_rank()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        pkg.parallel.init_process_group_from_env("nccl")
    if args.gpus != world and rank == 0:
        print("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run); using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)

    torch.manual_seed(0)
    model = pkg.UNet3D(precision=args.precision, **MODEL_KW).to(dev)
    crit = pkg.DiceLoss(sigmoid=True, include_background=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    sync = pkg.parallel.GradAllReduce(model.parameters())
    sync.broadcast_parameters(0)
    model.train()

    xh, th = synth_batch(BATCH_PER_GPU, seed=100 + rank)
    x, t = xh.to(dev), th.to(dev)

    def step_resident():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), t)
        loss.backward()
        sync()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    launches_per_step = model.launches_last_forward + model.launches_last_backward + 3   # + Dice sums/finalize/bwd

    # ---- e2e: host buffers in pinned memory, H2D inside the timed region, loss.item() every step
    pinned = [(xh.clone().pin_memory(), th.clone().pin_memory()) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    state = {"i": 0, "next": None}

    def prefetch():
        xs, ts = pinned[state["i"] % 2]
        state["i"] += 1
        with torch.cuda.stream(copy_stream):
            xd = xs.to(dev, non_blocking=True)
            td = ts.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        state["next"] = (xd, td, ev)

    def step_e2e():
        if state["next"] is None:
            prefetch()
        xd, td, ev = state["next"]
        torch.cuda.current_stream().wait_event(ev)
        xd.record_stream(torch.cuda.current_stream())
        td.record_stream(torch.cuda.current_stream())
        prefetch()                                   # next step's H2D overlaps this step's compute
        opt.zero_grad(set_to_none=True)
        loss, _ = pkg.train.batch_loss(model, xd, td, crit, n_gpus=1)
        loss.backward()
        sync()
        opt.step()
        return float(loss.item())                    # D2H read of the step's result

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel accounting (CUDA events around every launch of the plan), same K steps repeated
    plan = model._plan_for(x)
    macs = plan.algorithmic_macs()
    plan.profile_begin(args.steps * (launches_per_step + 8))
    for _ in range(args.steps):
        step_resident()
    torch.cuda.synchronize()
    prof = plan.profile_end()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    vols = BATCH_PER_GPU * world * args.steps
    value = vols / (ms_total / 1e3)
    e2e_value = vols / (ms_e2e / 1e3)
    # dominant kernel: k_conv_halo (forward and data-gradient convolutions on the halo-resident kernel: 42% of the step
    # in the ncu launch list profiles/r01_final_launches.txt); algorithmic FLOPs of exactly those launches / their time
    conv_ms = prof["conv_halo"]["ms"]
    conv_launches = prof["conv_halo"]["launches"]
    conv_flops = 2.0 * macs["conv_halo"] * args.steps
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    all_conv_ms = sum(prof[k]["ms"] for k in ("conv_halo", "conv_fwd", "conv_dgrad", "conv_wgrad"))
    all_conv_flops = 2.0 * sum(macs[k] for k in ("conv_halo", "conv_fwd", "conv_dgrad", "conv_wgrad")) * args.steps
    kernels = {}
    for k, v in prof.items():
        if v["launches"]:
            kernels[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
            if macs.get(k):
                kernels[k]["tflops"] = 2.0 * macs[k] * args.steps / (v["ms"] / 1e3) / 1e12
    step_flops = 2.0 * sum(macs.values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "bf16x3-split", "data": "synthetic",
        "config": {"workload": "C2: 4ch 128^3 UNet3D base_width=32, fwd + sigmoid-Dice + bwd + fused Adam, batch 2 per GPU",
                   "global_batch": BATCH_PER_GPU * world, "volume": list(VOLUME), "parallelism": "dp%d" % world,
                   "l2": "no flush: each step streams ~8 GB of activations, inputs (80 MB) exceed nothing but every tensor is re-read from HBM",
                   "grad_sync": "one flat fp32 NCCL all-reduce after backward" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(xh.numel() * 4 + th.numel()), "d2h_bytes_per_step": 4,
                "api": "train.batch_loss(model, images, target, criterion) + backward + Adam; pinned host buffers, H2D on a copy stream"},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "k_conv_halo (halo-resident implicit-GEMM conv: forward + data-gradient launches)",
                     "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                     # DRAM bytes per launch (read + write) of the same kernel, ncu capture profiles/r01_final_launches.txt
                     "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": "profiles/r01_final_launches.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the kernel's launches of one step)",
                     "all_conv_kernels_tflops": all_conv_flops / (all_conv_ms / 1e3) / 1e12 if all_conv_ms > 0 else 0.0,
                     "launches_per_step": conv_launches / args.steps, "ms_per_step": conv_ms / args.steps,
                     "peak_source": peaks["source"],
                     "timing": "CUDA-event pair around every launch on the launching stream, K steps repeated after the timed region",
                     "whole_step_frac_of_peak": step_flops / (ms_total / args.steps / 1e3) / 1e12 / peaks["tflops"]},
        "kernels": kernels,
        "algorithmic_flop_per_volume": step_flops / BATCH_PER_GPU,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_reference_throughput(steps=1, warmup=0, budget_s=30.0)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "split"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()

"""tcgen05.mma throughput vs N / layout / descriptor shape (cycles per M=128,K=16 MMA, one issuing warp per SM)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dunetcnn_b200")
L = pkg.lib
lib = L.load_library()
SW = {"none": 0, "sw128": 2, "sw64": 4, "sw32": 6}
out = torch.zeros(296, dtype=torch.int64, device="cuda")
print("--- with concurrent bulk copies into shared memory (operand-write pressure), sw128, dense descriptors")
src = torch.zeros(148 * 32768, dtype=torch.uint8, device="cuda")
print("%-6s %8s | cycles/MMA  TFLOP/s | copy B/clk/SM | operand-read B/clk" % ("N", "copy_B"))
for N in (64, 128, 256):
    for cb in (0, 8192, 16384, 32768):
        inner, reps = 8, 2000
        out.zero_()
        L.check(lib.b200unet_umma_rate(N, SW["sw128"], 1024, 1024, 32, inner, reps, 148, out.data_ptr(), src.data_ptr(), cb, 0, L.stream_ptr()))
        torch.cuda.synchronize()
        tot = out[:148].double().mean().item()
        cyc = tot / (inner * reps)
        tf = 2 * 128 * N * 16 / cyc * 148 * 1.9e9 / 1e12
        print("%-6d %8d | %7.1f   %7.0f | %8.1f | %8.1f" % (N, cb, cyc, tf, out[148:].double().mean().item() / tot, (128 + N) * 32 / cyc), flush=True)

print("--- stage-structured issue (wait + fence + elect{MMAs + commit}), 1 vs 2 issuing warps")
print("%-6s %6s %6s | cycles/MMA  TFLOP/s" % ("N", "MMA/st", "warps"))
for N in (32, 96, 128, 256):
    for mps in (4, 12):
        for nw, bit in ((1, 8), (2, 16)):
            stages = 9600 // mps
            out.zero_()
            L.check(lib.b200unet_umma_rate(N, SW["sw128"], 1024, 1024, 32, mps, stages, 148, out.data_ptr(), None, 0, bit, L.stream_ptr()))
            torch.cuda.synchronize()
            cyc = out[:148].double().mean().item() / (mps * stages)
            print("%-6d %6d %6d | %7.1f   %7.0f" % (N, mps, nw, cyc, 2 * 128 * N * 16 / cyc * 148 * 1.9e9 / 1e12), flush=True)
print("--- a tcgen05.commit after every `inner` MMAs (pipeline-stage hand-back), sw128 dense")
print("%-6s %6s | cycles/MMA  TFLOP/s" % ("N", "inner"))
for mode in (1, 3, 7, 6):
  print("hand-back mode bits (1 commit, 2 wait, 4 fence):", mode)
  for N in (32, 96, 128):
    for inner in (4, 12, 24):
        reps = 9600 // inner
        out.zero_()
        L.check(lib.b200unet_umma_rate(N, SW["sw128"], 1024, 1024, 32, inner, reps, 148, out.data_ptr(), None, 0, mode, L.stream_ptr()))
        torch.cuda.synchronize()
        cyc = out[:148].double().mean().item() / (inner * reps)
        print("%-6d %6d | %7.1f   %7.0f" % (N, inner, cyc, 2 * 128 * N * 16 / cyc * 148 * 1.9e9 / 1e12), flush=True)
print("%-6s %-6s %6s %6s %6s | cycles/MMA  -> TFLOP/s (148 SMs @1.9GHz)" % ("N", "layout", "a_sbo", "b_sbo", "a_step"))
for layout, rb in (("sw64", 64), ("sw32", 32)):
    for N in (16, 32, 64, 96, 128, 192, 256):
        for (a_sbo, a_step) in ((8 * rb, 32), (10 * rb, 32), (10 * rb, rb * 180)):
            inner, reps = 8, 400
            L.check(lib.b200unet_umma_rate(N, SW[layout], a_sbo, 8 * rb, a_step % 32768, inner, reps, 148, out.data_ptr(), None, 0, 0, L.stream_ptr()))
            torch.cuda.synchronize()
            cyc = out[:148].double().mean().item() / (inner * reps)
            tf = 2 * 128 * N * 16 / cyc * 148 * 1.9e9 / 1e12
            print("%-6d %-6s %6d %6d %6d | %7.1f   %7.0f" % (N, layout, a_sbo, 8 * rb, a_step, cyc, tf), flush=True)


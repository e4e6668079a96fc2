"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/NAME.txt
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/NAME.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def launches(src, dst):
    """Launch list of one warm step: time share per kernel and, when the capture has them, DRAM bytes per launch."""
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    data = [row for row in r if len(row) == len(hdr)]
    # one record per launch ID: {metric: value}
    recs = collections.OrderedDict()
    for d in data:
        key = d[idx["ID"]]
        rec = recs.setdefault(key, {"name": d[idx["Kernel Name"]]})
        v = float(d[idx["Metric Value"]].replace(",", ""))
        unit = d[idx["Metric Unit"]]
        m = d[idx["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v * 1e3 if unit == "s" else v   # -> ms
        else:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        rec[m] = v
    recs = list(recs.values())
    starts = [i for i, r_ in enumerate(recs) if "k_input_pack" in r_["name"]]
    # the last COMPLETE step: from the second-to-last k_input_pack to the last one (the capture may end mid-step)
    seg = recs[starts[-2]:starts[-1]] if len(starts) >= 2 else recs
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r_ in seg:
        name = re.sub(r"\(.*", "", r_["name"]).replace("void ", "")
        name = re.sub(r"<.*", "", name)
        tot[name][0] += 1
        tot[name][1] += r_.get("gpu__time_duration.sum", 0.0)
        tot[name][2] += r_.get("dram__bytes_read.sum", 0.0) + r_.get("dram__bytes_write.sum", 0.0)
    s_ = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none :\n")
        f.write("# one warm training step of `python bench.py` (cold-cache, serialised per-launch times: compare SHARES)\n")
        f.write("%-40s %6s %10s %7s %14s %12s\n" % ("kernel", "n", "ms", "share", "DRAM MB/launch", "DRAM GB/s"))
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s %6d %10.3f %6.1f%% %14.1f %12.0f\n" % (k[:40], v[0], v[1], 100 * v[1] / s_, v[2] / v[0] / 1e6,
                                                               v[2] / (v[1] / 1e3) / 1e9 if v[1] else 0))
        f.write("%-40s %6d %10.3f\n" % ("TOTAL", sum(v[0] for v in tot.values()), s_))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src)
        for d in data:
            f.write("---- %s  grid %s  block %s\n" % (d[hdr.index("Kernel Name")][:90], d[hdr.index("Grid Size")], d[hdr.index("Block Size")]))
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("   %-68s %s %s\n" % (m, d[i], units[i]))
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Summarises ncu reports into small CSV/markdown files for profiles/ (run where ncu is installed).
    python tools/ncu_summary.py launches gpurun_out/launches_c3.csv > profiles/r01_launches_c3.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/r01_prof.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
]


def short(name):
    return name.split("(")[0].split("::")[-1].replace("unnamed>", "").strip()


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu launch list ({path}): gpu__time_duration.sum, --clock-control none; cold-cache serialised times: compare SHARES\n")
    print("| kernel | launches | avg us | share of all launches |\n|---|---|---|---|")
    for k, v in agg.items():
        print(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print(f"# ncu --set full summary ({path}); per launch\n")
    for r in rows[2:]:
        print(f"## {short(r[ki])}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in h:
                i = h.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

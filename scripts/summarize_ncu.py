"""Turn ncu CSV output into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep   > profiles/rNN_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("vfp::", "")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else v
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {len(rows) - 1} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {t:10.1f} {t / n:9.1f} {100 * t / tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", short(r[hdr.index("Kernel Name")]))
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"    {m:75s} {r[i]:>14s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

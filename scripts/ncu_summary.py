#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: a per-kernel share table from a launch-list CSV
(--metrics gpu__time_duration.sum) and the key metrics of a --set full .ncu-rep."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ki].split("(")[0]
        a = agg.setdefault(n, [0, 0.0, r[ui]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time from %s (cold-cache, serialised: compare SHARES)" % path)
    print("%-50s %6s %16s %7s" % ("kernel", "count", "time", "share"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-50s %6d %13.0f %s %6.1f%%" % (n[:50], a[0], a[1], a[2], 100 * a[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    print("# key metrics from %s" % path)
    for v in rows[2:]:
        print("## launch: %s grid %s block %s" % (v[h.index("Kernel Name")][:80], v[h.index("Grid Size")], v[h.index("Block Size")]))
        for k in KEYS:
            if k in h:
                print("%-85s %-14s %s" % (k, u[h.index(k)], v[h.index(k)]))
        try:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            tm = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
            r = float(v[h.index("dram__bytes_read.sum")].replace(",", "")) * mult[u[h.index("dram__bytes_read.sum")]]
            w = float(v[h.index("dram__bytes_write.sum")].replace(",", "")) * mult[u[h.index("dram__bytes_write.sum")]]
            t = float(v[h.index("gpu__time_duration.sum")].replace(",", "")) * tm[u[h.index("gpu__time_duration.sum")]]
            print("derived: dram read+write = %.4g bytes per launch over %.4g s -> %.1f GB/s (under the profiler)" % (r + w, t, (r + w) / t / 1e9))
        except Exception as e:
            print("derived: n/a (%s)" % e)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (full if p.endswith(".ncu-rep") else launches)(p)
        print()

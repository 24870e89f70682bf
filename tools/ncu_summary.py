#!/usr/bin/env python
"""Summarise ncu output into text files that can be committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv        > profiles/x_launches_summary.txt
    python tools/ncu_summary.py report   gpurun_out/x.ncu-rep             > profiles/x_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-72s %6s %12s %8s %10s" % ("kernel", "n", "total us", "share", "avg us"))
    for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %6d %12.1f %8.3f %10.2f" % (n, v[0], v[1] / 1e3, v[1] / tot, v[1] / v[0] / 1e3))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none, per launch (from %s)" % path)
    for r in rows[2:]:
        print("\n== %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for k in KEYS:
            if k in hdr:
                print("   %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
              for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled") and hdr[i].endswith("per_issue_active.ratio")
              and "not_issued" not in hdr[i]]
        st.sort(key=lambda x: -x[1])
        print("   top stalls (warps per issue): " + ", ".join("%s %.2f" % s for s in st[:5]))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

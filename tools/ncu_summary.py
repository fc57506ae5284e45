#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/<name>.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof_k1.ncu-rep         > profiles/<name>.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|"
                  r"occupancy_limit_\w+)|sm__warps_active\.avg\.pct_of_peak_sustained_active|l1tex__t_sector_hit_rate\.pct|"
                  r"lts__t_sector_hit_rate\.pct|lts__t_sectors_srcunit_tex_op_read\.sum|lts__t_bytes\.sum|smsp__inst_executed\.sum|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio|smsp__warp_issue_stalled_\w+_per_warp_active\.pct|"
                  r"sm__pipe_tensor_cycles_active.*|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum)$")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:48]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-48s %6s %12s %8s" % ("kernel", "n", "total_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-48s %6d %12.1f %7.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
    print("%-48s %6d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        d = dict(zip(hdr, vals))
        print("## %s  grid=%s block=%s" % (d.get("Kernel Name", "?"), d.get("launch__grid_size"), d.get("launch__block_size")))
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h) and v not in ("", "0"):
                print("%-78s %16s %s" % (h, v, u))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2])

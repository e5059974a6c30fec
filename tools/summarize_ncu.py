#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof_conv.ncu-rep profiles/r1_conv_tc_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard", "lts__t_sector_hit_rate.pct"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        t = float(row["Metric Value"].replace(",", ""))
        t = t / 1000 if row["Metric Unit"] == "ns" else (t * 1000 if row["Metric Unit"] == "ms" else t)
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += t
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none); per-launch times are cold-cache and\n"
                "# serialised: compare SHARES.  %d launches, %.1f us total.\n\n" % (n, tot))
        f.write("| kernel | launches | total us | share | example grid x block |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% | %s x %s |\n" % (k, v[0], v[1], 100 * v[1] / tot, v[2], v[3]))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"]).decode()
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none (%s)\n" % src)
        for row in r[2:]:
            d = dict(zip(hdr, row))
            f.write("\n## %s  grid %s block %s\n\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
            for h, u, v in zip(hdr, units, row):
                if h in KEYS:
                    f.write("- %s = %s %s\n" % (h, v, u))
    print(open(dst).read())


def traffic(src, dst, batch="9", precision="2", workload="C2"):
    """per-launch dram bytes / duration / tensor-pipe % of the conv launches of ONE step (ncu --metrics ... -k regex:conv_tc) ->
    profiles/conv_traffic.json, the file bench.py reads `roofline.traffic` from."""
    import json
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(int(r[0]), {"kernel": re.sub(r"\(.*", "", r[4]).replace("void ", "")})[r[12]] = float(r[14].replace(",", ""))
    n = len(per)
    rd = sum(v.get("dram__bytes_read.sum", 0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0) for v in per.values())
    t = sum(v.get("gpu__time_duration.sum", 0) for v in per.values())
    tp = sum(v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * v.get("gpu__time_duration.sum", 0) for v in per.values()) / max(t, 1)
    out = {"source": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... "
                     "-k regex:conv_tc, the %d conv launches of one forward of %s frames, PE_GRAPH=0 PE_TWO_LANES=0)" % (src, n, batch),
           "batch": int(batch), "precision": int(precision), "workload": workload, "launches": n, "dram_read_bytes_per_step": rd,
           "dram_write_bytes_per_step": wr, "traffic_bytes_per_launch": (rd + wr) / max(n, 1), "tensor_pipe_pct_time_weighted": round(tp, 1),
           "gpu_time_us_sum_under_ncu": round(t / 1e3, 2)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

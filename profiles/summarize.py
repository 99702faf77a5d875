#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py <round-tag> <name>=<path.ncu-rep> ... [launches=<launches.csv>]

For every report: profiles/<tag>_<name>.txt with the metrics the roofline numbers come from (duration, DRAM bytes,
pipe utilisation, issue activity, occupancy limiters, top stall reasons) and the hottest source lines; for the launch
list: profiles/<tag>_launches.txt with per-kernel launch counts, mean duration and share of the step."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize_report(tag, name, rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, zip(units, vals)))
    lines = ["# %s  (ncu --set full --clock-control none; source: %s)" % (d.get("Kernel Name", ("", name))[1], os.path.basename(rep)), ""]
    for k in KEYS:
        if k in d:
            lines.append("%-70s %12s %s" % (k, d[k][1], d[k][0]))
    stalls = sorted(((float(v[1]), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                     and v[1] not in ("", "n/a")), reverse=True)
    lines += ["", "warp stall reasons per issued instruction (top 8):"]
    for v, k in stalls[:8]:
        lines.append("  %-40s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    src = ncu_csv(rep, "source")
    if len(src) > 3:
        h = src[1]
        iS, iN = h.index("Source"), h.index("# Samples")
        data = [r for r in src[2:] if len(r) > iN]
        tot = sum(int(r[iN]) for r in data) or 1
        lines += ["", "hottest SASS lines (share of warp samples):"]
        for r in sorted(data, key=lambda r: -int(r[iN]))[:10]:
            lines.append("  %5.1f%%  %s" % (100.0 * int(r[iN]) / tot, r[iS].strip()[:100]))
    path = os.path.join(HERE, "%s_%s.txt" % (tag, name))
    open(path, "w").write("\n".join(lines) + "\n")
    dram = None
    try:
        def b(key):
            u, v = d[key]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v.replace(",", "")) * mult
        dram = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
    except Exception:
        pass
    return path, dram


def summarize_launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[ki]].append(float(r[vi].replace(",", "")))
    ours = {k: v for k, v in agg.items() if "fma_loop" not in k and "at::" not in k and "elementwise" not in k}
    tot = sum(sum(v) for v in ours.values()) or 1.0
    lines = ["# launch list: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = " %5.1f%% of our kernels" % (100 * sum(v) / tot) if k in ours else ""
        lines.append("%-70s launches %3d  mean %.4f ms%s" % (k[:70], len(v), sum(v) / len(v) / 1e6, share))
    out = os.path.join(HERE, "%s_launches.txt" % tag)
    open(out, "w").write("\n".join(lines) + "\n")
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    traffic = {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        if name == "launches":
            print(summarize_launches(tag, path))
        else:
            p, dram = summarize_report(tag, name, path)
            print(p, dram)
            traffic[name] = dram
    if traffic:
        json.dump(traffic, open(os.path.join(HERE, "%s_dram_bytes.json" % tag), "w"), indent=1)

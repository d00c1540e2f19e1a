"""Turns the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
Usage: python profiles/summarize.py <tag>   (reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep)"""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
out = []

lp = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {k: j for j, k in enumerate(rows[h])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
        agg[name][0] += 1
        agg[name][1] += float(r[idx["Metric Value"]].replace(",", "")) * f
    tot = sum(v[1] for v in agg.values())
    out.append(f"# launch list ({os.path.basename(lp)}): ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised")
    out.append(f"# total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    out.append(f"{'ms':>10} {'share':>7} {'n':>6} {'avg_us':>10}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]:10.3f} {100 * v[1] / tot:6.2f}% {v[0]:6d} {1e3 * v[1] / v[0]:10.1f}  {k}")

rp = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum",
            "sm__ops_path_tensor_src_fp64.sum.per_second", "sm__inst_issued.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
    out.append("")
    out.append(f"# ncu --set full --clock-control none ({os.path.basename(rp)}), one block per captured launch")
    seen = set()
    for r in data:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"## {name}")
        for w in want:
            if w in idx:
                out.append(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
open(os.path.join(root, "profiles", f"summary_{tag}.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

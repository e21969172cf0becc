#!/usr/bin/env python
"""Summarise gpurun_out/{launches_<tag>.csv, prof_<tag>.ncu-rep} into profiles/<tag>_ncu.md."""
import csv
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
out = ["# ncu summary `%s`" % tag, "",
       "Command: `python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline %s` on one B200 "
       "(ncu --clock-control none).  Per-launch times are cold-cache and serialised: compare SHARES." % " ".join(sys.argv[2:]),
       "", "## launch list (gpu__time_duration.sum, all launches of the process)", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % tag)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki].split("(")[0][:70]
    d[k][0] += 1
    d[k][1] += v
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:25]:
    out.append("| `%s` | %d | %.3f | %.3f |" % (k, v[0], v[1] / 1e6, v[1] / tot))
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
if len(rows) > 2:
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
    out += ["", "## `--set full` capture", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        out.append("### `%s`" % name)
        out.append("")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                out.append("- %s = %s %s" % (w, r[i], units[i]))
        out.append("")
open("profiles/%s_ncu.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out[-40:]))

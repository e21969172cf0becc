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
# kernel class of each captured launch: the hop launches its kernels in a fixed order (hop.cu); compact form first
ORDER_COMPACT = ["spmm_mean_fwd", "proj_fwd_compact_tcgen05", "proj_fwd_tcgen05", "gather_gz_compact", "wgrad_tn_tcgen05",
                 "wgrad_tn_compact_tcgen05", "dgrad_nt_tcgen05", "dgrad_nt_compact_tcgen05", "spmm_transpose_bwd"]
ORDER_DENSE = ["spmm_mean_fwd", "proj_fwd_tcgen05", "wgrad_tn_tcgen05", "dgrad_nt_tcgen05", "spmm_transpose_bwd"]
if len(rows) > 2:
    import json
    caps = rows[2:]
    order = ORDER_COMPACT if len(caps) >= len(ORDER_COMPACT) else ORDER_DENSE
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

    def to_bytes(v, unit):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)

    kernels = {}
    for cls, r in zip(order, caps):
        kernels[cls] = {"kernel": r[hdr.index("Kernel Name")].split("(")[0][:80],
                        "dram_bytes_read": to_bytes(r[ir], units[ir]), "dram_bytes_write": to_bytes(r[iw], units[iw]),
                        "dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]),
                        "gpu_time_duration": "%s %s" % (r[it], units[it])}
    json.dump({"source": "profiles/%s_ncu.md (ncu --set full --clock-control none, one hop of `bench.py %s`)" % (tag, " ".join(sys.argv[2:])),
               "kernels": kernels}, open("profiles/r02_traffic.json" if tag.startswith("r02") else "profiles/%s_traffic.json" % tag, "w"), indent=1)
    out += ["", "## DRAM traffic per launch and kernel class (feeds bench.py's roofline.traffic)", "",
            "| class | kernel | read GB | write GB |", "|---|---|---:|---:|"]
    for cls, v in kernels.items():
        out.append("| %s | `%s` | %.3f | %.3f |" % (cls, v["kernel"], v["dram_bytes_read"] / 1e9, v["dram_bytes_write"] / 1e9))
open("profiles/%s_ncu.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out[-40:]))

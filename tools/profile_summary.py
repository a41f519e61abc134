"""Turn an .ncu-rep (ncu --set full) into the committed text/JSON summaries under profiles/.
usage: python tools/profile_summary.py <report.ncu-rep> <out-stem> [json-key]"""
import csv
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_sleeping"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


rep, stem = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
lines, js = [f"# ncu --set full summary of {os.path.basename(rep)}", ""], {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append(f"## {name}")
    for w in WANT:
        if w in hdr:
            lines.append(f"    {w:78s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
    wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
    lines.append(f"    => dram traffic per launch = {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})")
    lines.append("")
    js[name] = {"dram_bytes_per_launch": rd + wr, "duration": r[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")]}
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
ops = {}
for r in csv.reader(sass.splitlines()):
    if len(r) > 6 and r[0].startswith("0x") or (len(r) > 6 and r[0][:1].isdigit()):
        t = r[1].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        try:
            ops[op] = ops.get(op, 0) + int(r[5])
        except (ValueError, IndexError):
            pass
if ops:
    tot = sum(ops.values())
    lines.append("## executed warp instructions by SASS opcode (all kernels in the report)")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
        lines.append(f"    {op:10s} {n:14d}  {100 * n / tot:5.1f}%")
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{stem}.txt", "w").write("\n".join(lines) + "\n")
if key:
    path = "profiles/ncu_summary.json"
    d = json.load(open(path)) if os.path.exists(path) else {}
    first = next(iter(js.values()))
    d[key] = dict(first, report=os.path.basename(rep), kernel=next(iter(js)))
    json.dump(d, open(path, "w"), indent=1)
print("\n".join(lines))

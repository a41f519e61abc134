"""Per-source-line hot spots of an ncu report captured with --import-source on.
usage: python tools/ncu_source_hot.py report.ncu-rep [top]
Prints, per CUDA source line, executed warp instructions and stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fpath = ""
lines = {}
cur = None
hdr = None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_smp, i_inst = r.index("# Samples"), r.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= i_inst:
        continue
    if r[0]:          # a CUDA source line
        cur = (fpath, int(r[0]))
        lines.setdefault(cur, [r[1].strip(), 0, 0])
    elif cur:         # a SASS instruction attributed to it
        num = lambda s: int(s) if s.strip().isdigit() else 0
        lines[cur][1] += num(r[i_inst])
        lines[cur][2] += num(r[i_smp])
tot_i = sum(v[1] for v in lines.values()) or 1
tot_s = sum(v[2] for v in lines.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("by samples:")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"  {100*v[2]/tot_s:5.1f}% smp {100*v[1]/tot_i:5.1f}% inst  {f}:{ln}  {v[0][:90]}")

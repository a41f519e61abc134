"""SASS evidence for the Blackwell-native paths: per kernel of libsignalops_cuda.so, how many TMA (UTMALDG / UTMASTG /
UBLKCP), FP64 tensor-core (DMMA), mbarrier (SYNCS) and register-rebalancing (USETMAXREG) instructions the
binary holds.  usage: python tools/sass_evidence.py > profiles/r2_sass_opcodes.txt   (needs only cuobjdump)"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "signaloperators.jl_b200", "lib", "libsignalops_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "DMMA", "SYNCS", "USETMAXREG", "LDGSTS", "DFMA", "DMUL", "DADD", "LDS", "STS",
         "UTC", "LDTM", "HMMA"]
kern, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
        counts[kern]["_all"] += 1
demangled = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (static, sm_100a)")
print("# TMA = UTMALDG/UTMASTG (tensor-map copies), UBLKCP (bulk copies); FP64 tensor cores = DMMA (mma.sync.m8n8k4.f64; tcgen05 has no")
print("# FP64 kind, so there is no UTC*MMA / LDTM by design); SYNCS = mbarrier; USETMAXREG = setmaxnreg")
tot = collections.Counter()
rows = []
for (k, c), name in zip(counts.items(), demangled or counts):
    name = name[:name.rfind(">") + 1] if ">" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void sigops::", "").replace("sigops::", "").replace("(int)", "").replace("(bool)", "")
    if not any(c[w] for w in ("UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "USETMAXREG")):
        for w in WATCH:
            tot[w] += c[w]
        continue
    rows.append((name, c))
    for w in WATCH:
        tot[w] += c[w]
cols = ["UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "SYNCS", "USETMAXREG", "DFMA", "LDS", "STS"]
print(f"{'kernel':70s} " + " ".join(f"{w:>10s}" for w in cols) + f" {'all':>8s}")
for name, c in rows:
    print(f"{name[:70]:70s} " + " ".join(f"{c[w]:10d}" for w in cols) + f" {c['_all']:8d}")
print(f"{'whole library (' + str(len(counts)) + ' kernels)':70s} " + " ".join(f"{tot[w]:10d}" for w in cols))
print(f"# UTC*MMA / LDTM / HMMA anywhere: {tot['UTC']} / {tot['LDTM']} / {tot['HMMA']}")

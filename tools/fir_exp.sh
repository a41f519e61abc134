#!/bin/bash
# Ablations of the tensor-map FIR kernel on config 3 (64 signals); SIGOPS_FIR_EXP variants give wrong results by design
# (bit mask: 1 no staging stores, 2 no tensor stores, 4 no tap-band building, 8 no ring loads, 16 no barrier waits).
#   gpurun -- 'bash tools/fir_exp.sh'
timeout -k 10 120 python tools/profile_step.py cfg3 10 2>&1 | tail -1
echo "helper-built bands:"; SIGOPS_NO_FIR_BANDS=1 timeout -k 10 120 python tools/profile_step.py cfg3 10 2>&1 | tail -1
for e in 4 8 7 31; do echo "SIGOPS_NO_FIR_BANDS=1 SIGOPS_FIR_EXP=$e"; SIGOPS_NO_FIR_BANDS=1 SIGOPS_FIR_EXP=$e timeout -k 10 120 python tools/profile_step.py cfg3 10 2>&1 | tail -1; done
SIGOPS_NO_FIR_BANDS=1 SIGOPS_FIR_DBG=1 timeout -k 10 120 python tools/profile_step.py cfg3 2 2>&1 | grep "cycles/tile" | tail -1
timeout -k 10 200 python tools/profile_step.py cfg3 10 1024 2>&1 | tail -1

for e in 0 1 3 4 8 12 7; do echo "EXP=$e"; SIGOPS_FIR_EXP=$e SIGOPS_FIR_DBG=1 timeout -k 10 120 python tools/profile_step.py cfg3 3 2>&1 | tail -2; done
for ns in 7 8; do echo "NSLOT=$ns"; SIGOPS_FIR_NSLOT=$ns SIGOPS_FIR_DBG=1 timeout -k 10 120 python tools/profile_step.py cfg3 3 2>&1 | tail -2; done

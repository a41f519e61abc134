for e in 0 4 1 3 7; do echo "EXP=$e"; SIGOPS_FIR_EXP=$e timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -1; done
for ns in 7 8; do echo "NSLOT=$ns"; SIGOPS_FIR_NSLOT=$ns timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -1; done
for tps in 305 203 1218; do echo "TPS=$tps"; SIGOPS_FIR_TPS=$tps timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -1; done

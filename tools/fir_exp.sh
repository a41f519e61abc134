for g in 1 2; do echo "GROUPS=$g"; SIGOPS_FIR_GROUPS=$g timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -1; done
SIGOPS_FIR_DBG=1 timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -2
timeout -k 10 200 python tools/profile_step.py cfg3 5 1024 2>&1 | tail -1

SIGOPS_DEBUG=1 timeout -k 10 200 python tools/profile_step.py cfg4 5 2>&1 | tail -40

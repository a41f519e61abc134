timeout -k 10 120 python tools/profile_step.py cfg3 5 2>&1 | tail -1
timeout -k 10 120 python tools/profile_step.py cfg3a 5 2>&1 | tail -1
timeout -k 10 200 python tools/profile_step.py cfg3 5 1024 2>&1 | tail -1
timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3

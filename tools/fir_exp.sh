timeout -k 10 900 python -m pytest tests/test_gpu_iir_tmap.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
SIGOPS_DEBUG=1 timeout -k 10 200 python tools/profile_step.py cfg5full 5 2>&1 | grep -E "IIR stage|cfg5full" | tail -2
timeout -k 10 200 python tools/profile_step.py cfg5full 5 4 2>&1 | tail -1
timeout -k 10 200 python tools/profile_step.py cfg5 5 2>&1 | tail -1

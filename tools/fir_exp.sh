timeout -k 10 600 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_fir_mma.py tests/test_gpu_parity.py tests/test_gpu_resample_full.py -x -q -m gpu 2>&1 | tail -2
timeout -k 10 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_hostpath.py -x -q -m gpu -k "every_output and scalar" 2>&1 | tail -8

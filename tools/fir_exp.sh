timeout -k 10 600 python -m pytest tests/test_gpu_randn.py tests/test_gpu_iir_tmap.py -x -q -m gpu 2>&1 | tail -15

#!/bin/bash
# One gpurun call of round 2 (arguments pick the parts to run).
mkdir -p gpurun_out
for part in "$@"; do
case $part in
info)
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
  nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ;;
link) python tools/host_link_probe.py 2>&1 | tee gpurun_out/host_link.txt ;;
dmma) ./tools/dmma_peak 2>&1 | tee gpurun_out/dmma_peak.txt ;;
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ;;
hosttests) python -m pytest tests/test_gpu_hostpath.py -m gpu -x -q 2>&1 | tail -25 ;;
firtests) python -m pytest tests/test_gpu_fir_mma.py tests/test_gpu_resample_full.py tests/test_resample_analytic.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 ;;
bench) python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err ;;
benchref) python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; tail -c 1500 gpurun_out/bench_ref.json ;;
steps)
  for c in cfg1 cfg2 cfg3 cfg3a cfg4 cfg5; do python tools/profile_step.py $c 5 2>&1 | tail -1; done ;;
cfg3big) python tools/profile_step.py cfg3 5 1024 2>&1 | tail -1 ;;
*) echo "running: $part"; bash -c "$part" ;;
esac
done

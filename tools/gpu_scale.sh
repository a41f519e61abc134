#!/bin/bash
# One scaling point of bench.py (both arms + the single-process mode) on N GPUs of one box:  gpurun --gpus N -- bash tools/gpu_scale.sh N
N=$1
mkdir -p gpurun_out
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err
echo "refrc=$?"
timeout -k 10 600 python bench.py --single-process --gpus $N --steps 3 > gpurun_out/bench_sp_n$N.json 2>> gpurun_out/bench_n$N.err
echo "sprc=$?"

timeout -k 10 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
SIGOPS_DEBUG=1 timeout -k 10 120 python tools/profile_step.py cfg3 10 2>&1 | grep -E "epochs|cfg3:" | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; echo "ref rc=$?"

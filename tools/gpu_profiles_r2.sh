#!/bin/bash
# Round-2 profile captures (one gpurun call).  Reports land in gpurun_out/; summaries are made on the CPU box.
mkdir -p gpurun_out
# launch list of the bench command (device times are cold-cache and serialised: shares only)
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --ninst 256 --e2e-ninst 32 > gpurun_out/r2_bench_under_ncu.log 2>&1
# full captures of the dominant kernels
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_fir_tmap -s 2 -c 1 -f -o gpurun_out/r2_fir_tmap python tools/profile_step.py cfg3 3 > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_iir_tmap -s 2 -c 1 -f -o gpurun_out/r2_iir_tmap_lv python tools/profile_step.py cfg5full 3 4 > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_iir_tmap -s 2 -c 1 -f -o gpurun_out/r2_iir_tmap python tools/profile_step.py cfg2 3 > /dev/null 2>&1
ls -la gpurun_out/r2_*.ncu-rep

# same-box A/B: the tree before this session's FIR changes (build_old/) against the current one, config 3 at 1024 signals
for i in 1 2; do
echo "old:"; (cd build_old && timeout -k 10 300 python tools/profile_step.py cfg3 40 1024 2>&1 | tail -1)
echo "new:"; timeout -k 10 300 python tools/profile_step.py cfg3 40 1024 2>&1 | tail -1
done
echo "old 64:"; (cd build_old && timeout -k 10 300 python tools/profile_step.py cfg3 20 2>&1 | tail -1)
echo "new 64:"; timeout -k 10 300 python tools/profile_step.py cfg3 20 2>&1 | tail -1
nvidia-smi --query-gpu=power.limit,clocks.max.sm,clocks.sm,temperature.gpu --format=csv

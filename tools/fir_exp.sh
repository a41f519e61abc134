timeout -k 10 600 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_wav.py tests/test_gpu_randn.py -x -q -m gpu 2>&1 | tail -3
for t in 8 16; do
  SIGOPS_COPY_THREADS=$t python bench.py --steps 3 --warmup 3 --ninst 128 --configs '' 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('threads $t: pinned', round(e['value']), 'pageable', round(e['pageable']['value']), 'api', round(e['public_api']['value']))"
done

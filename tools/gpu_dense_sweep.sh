timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_comm.py tests/test_gpu_interval_sharding.py tests/test_golden_fixtures.py -q -m gpu -k "dense or config4 or comm or fused" 2>&1 | tail -8
for w in 16 12 8; do
  CB200_DENSE20_WARPS=$w timeout 300 python bench.py --only 4 --steps 3 --warmup 3 --no-cpu-baseline --candidates 8192 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('warps $w', 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
"
done

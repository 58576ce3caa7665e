set -x
for cfg in 3 0 4 5; do
  PCS_NE_CFG=$cfg python bench.py --steps 100 --warmup 5 --no-cpu --no-lm 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])
"
done
PCS_NE_CFG=4 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -3
PCS_NE_CFG=5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -3

# A/B of K_ne variants (PCS_NE_CFG values as arguments): kernel time from the library's CUDA events
for cfg in ${@:-3 0}; do
  PCS_NE_CFG=$cfg python bench.py --steps 200 --warmup 10 --no-cpu --no-lm --no-callbacks 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))
"
done

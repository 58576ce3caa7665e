# A/B of K_ne variants: "cfg[:waves]" arguments (PCS_NE_CFG, PCS_NE_WAVES); kernel time from the library's CUDA events
for spec in ${@:-3 0}; do
  cfg=${spec%%:*}; waves=${spec##*:}; [ "$waves" = "$spec" ] && waves=1
  PCS_NE_CFG=$cfg PCS_NE_WAVES=$waves python bench.py --steps 200 --warmup 10 --no-cpu --no-lm --no-callbacks 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg waves $waves', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))
"
done

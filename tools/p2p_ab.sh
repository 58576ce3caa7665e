#!/bin/bash
# single-CTA vs multi-CTA peer-memory all-reduce at N = $1 (weak scaling bench, short)
N=${1:-8}
for m in 0 1 0 1; do
  PCS_P2P_MULTI=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+m)) bench.py --gpus $N --steps 100 --warmup 10 --no-lm --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('multi=$m', round(d['value']), round(d['ms_per_step'],4), d['config']['exchange'], d['config']['exchange_check'])"
done

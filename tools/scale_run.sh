#!/bin/bash
# bench.py at N = 2 .. $1 GPUs of this box (weak scaling, config 4 per GPU) + config 5 (dome128, strong scaling) at N = $1
MAXN=${1:-4}; MINN=${2:-2}
for N in 2 4 8; do
  [ $N -le $MAXN ] && [ $N -ge $MINN ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510+N)) bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "N=$N rc=$?"; tail -1 gpurun_out/scale_n$N.json | cut -c1-400
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXN --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $MAXN --workload dome128 --steps 20 --warmup 3 --lm-iters 5 > gpurun_out/dome128_n$MAXN.json 2> gpurun_out/dome128_n$MAXN.err
echo "dome128 N=$MAXN rc=$?"; tail -1 gpurun_out/dome128_n$MAXN.json | cut -c1-600; tail -3 gpurun_out/dome128_n$MAXN.err

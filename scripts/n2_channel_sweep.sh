#!/bin/bash
# N=2 (or N=$1): NCCL channels = SMs reserved under the overlapped exchange, after the round-2 changes (reserve only in the
# exchange window, high-priority communication stream, bucket-pipelined early Adam)
N=${1:-2}
mkdir -p gpurun_out
for ch in 12 16 24 32; do
  TACORL_NCCL_CHANNELS=$ch timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-tacorl 2>gpurun_out/n${N}_ch_$ch.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=$N channels=$ch', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done 2>&1 | tee gpurun_out/n${N}_channel_sweep.txt

#!/bin/bash
# N=2: NCCL protocol for the 64 MB gradient buckets (the tuner picks RING_LL on this box: 245 us per bucket)
mkdir -p gpurun_out
for proto in default Simple LL128; do
  if [ "$proto" = default ]; then unset NCCL_PROTO; else export NCCL_PROTO=$proto; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-tacorl --timeline gpurun_out/timeline_n2_$proto.json 2>gpurun_out/n2_proto_$proto.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('proto=$proto', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done 2>&1 | tee gpurun_out/n2_proto_sweep.txt

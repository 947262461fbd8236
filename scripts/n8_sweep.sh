#!/bin/bash
# N-GPU sweep of the overlapped gradient all-reduce settings (NCCL channels = SMs left free by the persistent conv kernels,
# wire dtype, algorithm).  Usage: scripts/n8_sweep.sh [N]   -> one line per configuration (ms/step resident, e2e).
N=${1:-8}
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-tacorl --no-cpu-baseline --no-fp32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
}
run NCCL_MAX_NCHANNELS=16 NCCL_MIN_NCHANNELS=16 TACORL_SM_RESERVE=16 TACORL_WIRE=fp32
run NCCL_MAX_NCHANNELS=8 NCCL_MIN_NCHANNELS=8 TACORL_SM_RESERVE=8 TACORL_WIRE=fp32
run NCCL_MAX_NCHANNELS=24 NCCL_MIN_NCHANNELS=24 TACORL_SM_RESERVE=24 TACORL_WIRE=fp32
run NCCL_MAX_NCHANNELS=32 NCCL_MIN_NCHANNELS=32 TACORL_SM_RESERVE=32 TACORL_WIRE=fp32
run NCCL_MAX_NCHANNELS=32 NCCL_MIN_NCHANNELS=32 TACORL_SM_RESERVE=32 TACORL_WIRE=bf16
run NCCL_MAX_NCHANNELS=16 NCCL_MIN_NCHANNELS=16 TACORL_SM_RESERVE=16 TACORL_WIRE=fp32 NCCL_ALGO=Tree
run NCCL_MAX_NCHANNELS=16 NCCL_MIN_NCHANNELS=16 TACORL_SM_RESERVE=0 TACORL_WIRE=fp32

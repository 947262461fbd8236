# N=2 sweep of the SM reserve / NCCL channel count (every launch under its own timeout)
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 30 --warmup 5 --no-tacorl 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"; }
TACORL_SM_RESERVE=16 run "reserve=16,nch=default"
NCCL_MAX_NCHANNELS=32 NCCL_MIN_NCHANNELS=32 TACORL_SM_RESERVE=32 run "reserve=32,nch=32"
NCCL_MAX_NCHANNELS=24 NCCL_MIN_NCHANNELS=24 TACORL_SM_RESERVE=24 run "reserve=24,nch=24"

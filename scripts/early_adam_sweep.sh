#!/bin/bash
# N=1: where to put the Adam update of the sub-networks whose gradient is final early (decoder, plan recogniser):
# TACORL_EARLY_SUBNET=1 starts each sub-network's slice as its own BPTT finishes (under the other sub-network's
# recurrence chains), TACORL_EARLY_BG=1 uses the background-sized grid (one CTA per SM).
mkdir -p gpurun_out
for sub in 0 1; do for bg in 0 1; do
  TACORL_EARLY_SUBNET=$sub TACORL_EARLY_BG=$bg python bench.py --steps 30 --warmup 5 --no-tacorl --no-fp32 --no-cpu-baseline \
    2>gpurun_out/early_${sub}_${bg}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('subnet=$sub bg=$bg', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done; done 2>&1 | tee gpurun_out/early_adam_sweep.txt

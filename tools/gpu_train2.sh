#!/bin/bash
# usage: gpu_train2.sh  -> train workload at N = 2 under several NCCL channel limits (gpurun --gpus 2)
set -u
O=gpurun_out; mkdir -p $O
for ch in default 2 4 8; do
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload train --steps 20 --warmup 3 > $O/train2_$ch.json 2> $O/train2_$ch.err
  python - <<PY
import json
d=json.loads(open("$O/train2_$ch.json").read().strip().splitlines()[-1])
print("$ch", round(d["ms_per_step"],3), d.get("regions_ms_per_step"), {k:round(v,3) for k,v in d.items() if "allreduce" in k and isinstance(v,float)})
PY
done

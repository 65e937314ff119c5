#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-c}
timeout 200 python tools/bench_conv.py --res 0 --timeline 1 --shapes layer1,layer2,layer3,layer4 2>&1 | tee $O/s5_bench_conv_$TAG.txt
timeout 200 python tools/bench_conv.py --res 1 --shapes layer1,layer2,layer3,layer4 2>&1 | tee -a $O/s5_bench_conv_$TAG.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "conv or stem or neck or network or stride or backbone or bf16" > $O/s5_pytest_$TAG.log 2>&1; echo "pytest rc $?"; tail -3 $O/s5_pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/s5_bench_$TAG.json 2> $O/s5_bench_$TAG.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("$O/s5_bench_$TAG.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["launches_per_step"], d["stages_ms"], d["parity_spot"]["status"])
PY

#!/bin/bash
# conv_flat producer split: micro-benchmark per layer shape for NA = default / 2 / 3, kernel tests, bench
set -u
O=gpurun_out; mkdir -p $O
for na in 0 2 3; do
  for res in 0 1; do
    echo "== YAD_FLAT_NA=$na res=$res"
    YAD_FLAT_NA=$na timeout 300 python tools/bench_conv.py --res $res --shapes layer1,layer2,layer3,layer4
  done
done > $O/s5_bench_conv_a.txt 2>&1
cat $O/s5_bench_conv_a.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "conv or stem or neck or network or stride or backbone or bf16" > $O/s5_pytest_a.log 2>&1; echo "pytest rc $?"; tail -5 $O/s5_pytest_a.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/s5_bench_a.json 2> $O/s5_bench_a.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("$O/s5_bench_a.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["launches_per_step"], d["stages_ms"], d["parity_spot"]["status"])
PY

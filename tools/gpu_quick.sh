#!/bin/bash
# usage: gpu_quick.sh TAG [pytest -k expression]   -> gpurun_out/r2_pytest_TAG.log, r2_bench_TAG.json
set -u
O=gpurun_out; TAG=$1; K=${2:-}
mkdir -p $O
if [ -n "$K" ]; then python -m pytest tests -x -q -m gpu -k "$K" > $O/r2_pytest_$TAG.log 2>&1; else python -m pytest tests -x -q -m gpu > $O/r2_pytest_$TAG.log 2>&1; fi
echo "pytest rc $?"; tail -3 $O/r2_pytest_$TAG.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_bench_$TAG.json 2> $O/r2_bench_$TAG.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("$O/r2_bench_$TAG.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["launches_per_step"], d["stages_ms"], d["parity_spot"]["status"])
PY

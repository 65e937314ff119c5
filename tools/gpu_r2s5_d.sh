#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-p}
echo "== single-CTA path"
YAD_FLAT_PAIR=0 timeout 300 python -m pytest tests -x -q -m gpu -k "conv_flat" 2>&1 | tail -3
echo "== pair path"
YAD_FLAT_PAIR=1 timeout 300 python -m pytest tests -x -q -m gpu -k "conv_flat" > $O/s5_pytest_pair_$TAG.log 2>&1; echo "pytest rc $?"; tail -15 $O/s5_pytest_pair_$TAG.log
for pm in 0 1; do
  echo "== bench_conv PAIR=$pm"
  YAD_FLAT_PAIR=$pm timeout 200 python tools/bench_conv.py --res 0 --timeline 1 --shapes layer1,layer2,layer3,layer4 2>&1 | grep -v "MMA warp"
  YAD_FLAT_PAIR=$pm timeout 200 python tools/bench_conv.py --res 1 --shapes layer1,layer2,layer3,layer4 2>&1
done | tee $O/s5_bench_conv_pair_$TAG.txt

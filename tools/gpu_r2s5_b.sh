#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for f in 0 1 2 3 4 7; do
  echo "== YAD_FLAT_DBG=$f"
  YAD_FLAT_DBG=$f timeout 300 python tools/bench_conv.py --res 0 --shapes layer1,layer2,layer3,layer4
done > $O/s5_bench_conv_b.txt 2>&1
cat $O/s5_bench_conv_b.txt

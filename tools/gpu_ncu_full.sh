#!/bin/bash
# usage: gpu_ncu_full.sh TAG KERNEL_REGEX SKIP [BATCH]  -> gpurun_out/r2_full_TAG.ncu-rep (one --set full capture with source)
set -u
O=gpurun_out; TAG=$1; K=$2; SKIP=$3; BATCH=${4:-512}
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $O/r2_full_$TAG -f python tools/profile_step.py --batch $BATCH --warm 1 --steps 1 > $O/ncu_full_$TAG.log 2>&1; echo "ncu rc $?"
tail -3 $O/ncu_full_$TAG.log

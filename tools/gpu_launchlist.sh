#!/bin/bash
# usage: gpu_launchlist.sh TAG  -> gpurun_out/r2_launches_TAG.csv (ncu per-launch durations + tensor pipe of one 512-clip step)
set -u
O=gpurun_out; TAG=$1
mkdir -p $O
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:'frontend|conv_|neck_|decode|nms|stem_|hmean|resize|sppf|compact' -c 120 --csv --log-file $O/r2_launches_$TAG.csv \
  python tools/profile_step.py --batch 512 --warm 1 --steps 1 > $O/ncu_launch_$TAG.log 2>&1; echo "ncu list rc $?"
python tools/launch_table.py $O/r2_launches_$TAG.csv

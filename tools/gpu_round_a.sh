#!/bin/bash
# One GPU-box visit: gpu tests, default bench, launch list, one --set full capture of the neck, sanitizer passes on kernel-level tests.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r2_pytest_h.log 2>&1; echo "pytest rc $?"
python bench.py --steps 20 --warmup 3 > $O/r2_bench_h.json 2> $O/r2_bench_h.err; echo "bench rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_h.csv python tools/profile_step.py --batch 512 --warm 2 --steps 1 > $O/ncu_launch_h.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:neck_fused -s 2 -c 1 -o $O/r2_neck_v2 -f python tools/profile_step.py --batch 512 --warm 2 --steps 1 > $O/ncu_neck.log 2>&1; echo "ncu neck rc $?"
SEL='test_conv_flat or test_stem_conv_tensor_core or test_fused_stem_vs_two_convs or test_frontend_short or test_decode_vs_oracle or test_nms_edge or (test_fused_neck and 2.0) or (test_conv_kernels and tc_bf16)'
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $O/r2_sanitizer_$tool.txt 2>&1; echo "sanitizer $tool rc $?"
done

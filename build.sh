#!/bin/bash
# Builds the C-ABI library in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$ROOT/yolo-inspired-audio-activity-detection_b200/csrc"
OUT="$SRC/libyad_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -I"$ROOT/include")
mkdir -p "$SRC/build"
pids=()
for f in api frontend conv_simt conv_tc conv_flat conv_stem_tc conv_stem_fused neck_fused conv_tf32 glue decode_nms train_ops loss train_net anchors; do
  [ -f "$SRC/$f.cu" ] || continue
  if [ ! -f "$SRC/build/$f.o" ] || [ "$SRC/$f.cu" -nt "$SRC/build/$f.o" ] || [ "$SRC/common.cuh" -nt "$SRC/build/$f.o" ] || [ "$SRC/tc_ptx.cuh" -nt "$SRC/build/$f.o" ] || [ "$SRC/fft500.cuh" -nt "$SRC/build/$f.o" ] \
     || [ "$ROOT/include/yad_b200.h" -nt "$SRC/build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} -c "$SRC/$f.cu" -o "$SRC/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "$SRC"/build/*.o -gencode arch=compute_100a,code=sm_100a
echo "built $OUT"

#!/bin/bash
# Build libva_sm100.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libva_sm100.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
       --expt-relaxed-constexpr -Xptxas -v)
OBJS=()
mkdir -p "$HERE/build"
for f in va_logits va_upsample va_tail va_fused_tc va_api; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ "$HERE/va_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/va_up_common.cuh" -nt "$HERE/build/$f.o" ] \
     || [ "$HERE/../../include/vision_assist_b200.h" -nt "$HERE/build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" 2> "$HERE/build/$f.ptxas.log" || { cat "$HERE/build/$f.ptxas.log"; exit 1; }
  fi
  OBJS+=("$HERE/build/$f.o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${OBJS[@]}" -cudart static
echo "built $OUT"

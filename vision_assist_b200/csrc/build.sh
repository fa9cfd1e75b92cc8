#!/bin/bash
# Build libva_sm100.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${VA_OUT:-$HERE/../libva_sm100.so}"
BUILD="${VA_BUILD_DIR:-$HERE/build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
       --expt-relaxed-constexpr -Xptxas -v ${VA_EXTRA_FLAGS:-})
OBJS=()
mkdir -p "$BUILD"
for f in va_logits va_upsample va_tail va_fused_tc va_nms va_api; do
  if [ ! -f "$BUILD/$f.o" ] || [ "$HERE/$f.cu" -nt "$BUILD/$f.o" ] || [ "$HERE/va_common.cuh" -nt "$BUILD/$f.o" ] || [ "$HERE/va_up_common.cuh" -nt "$BUILD/$f.o" ] \
     || [ "$HERE/va_contour_core.h" -nt "$BUILD/$f.o" ] || [ "$HERE/va_contour_lut.h" -nt "$BUILD/$f.o" ] \
     || [ "$HERE/../../include/vision_assist_b200.h" -nt "$BUILD/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$BUILD/$f.o" 2> "$BUILD/$f.ptxas.log" || { cat "$BUILD/$f.ptxas.log"; exit 1; }
  fi
  OBJS+=("$BUILD/$f.o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${OBJS[@]}" -cudart static
echo "built $OUT"

#!/bin/bash
# Builds flame_b200/lib/libflame_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT="$ROOT/flame_b200/lib"
mkdir -p "$OUT/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -I$ROOT/include -I$HERE ${FLAME_B200_NVCC_EXTRA:-}"
pids=()
for f in pd raster stereo capi; do
  $NVCC $FLAGS -c "$HERE/$f.cu" -o "$OUT/obj/$f.o" &
  pids+=($!)
done
$NVCC $FLAGS -c "$HERE/delaunay.cc" -o "$OUT/obj/delaunay.o" &  # host-only (frame pipeline's triangulator)
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libflame_b200.so" "$OUT/obj/pd.o" "$OUT/obj/raster.o" "$OUT/obj/stereo.o" "$OUT/obj/capi.o" "$OUT/obj/delaunay.o" -lcudart
echo "built $OUT/libflame_b200.so"

#!/bin/bash
# Compiles the reference's own hot-path sources, where they lie under $1 (default
# /root/reference), against oracle/shim into oracle/_ref/libflame_ref.so. Test infrastructure
# only. Never copies reference sources into the repo; outputs only under oracle/_ref/.
set -euo pipefail
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
OBJ="$OUT/obj"
mkdir -p "$OBJ"
CXXFLAGS="-O3 -std=c++11 -ffp-contract=off -fPIC -w -I$HERE/shim -I$REF/src"
g++ -O2 -fPIC -w -I"$REF/src" -c "$REF/src/flame/external/triangle/triangle.cpp" -o "$OBJ/triangle.o"
g++ $CXXFLAGS -c "$REF/src/flame/utils/rasterization.cc" -o "$OBJ/rasterization.o"
g++ $CXXFLAGS -c "$REF/src/flame/utils/image_utils.cc" -o "$OBJ/image_utils.o"
g++ $CXXFLAGS -c "$HERE/ref_capi/ref_triangle.cc" -o "$OBJ/ref_triangle.o"
g++ $CXXFLAGS -c "$HERE/ref_capi/ref_raster.cc" -o "$OBJ/ref_raster.o"
OBJS="$OBJ/triangle.o $OBJ/rasterization.o $OBJ/image_utils.o $OBJ/ref_triangle.o $OBJ/ref_raster.o"
if [ -f "$HERE/ref_capi/ref_solver.cc" ]; then
  g++ $CXXFLAGS -c "$REF/src/flame/optimizers/nltgv2_l1_graph_regularizer.cc" -o "$OBJ/regularizer.o"
  g++ $CXXFLAGS -c "$HERE/ref_capi/ref_solver.cc" -o "$OBJ/ref_solver.o"
  OBJS="$OBJS $OBJ/regularizer.o $OBJ/ref_solver.o"
fi
if [ -f "$HERE/ref_capi/ref_stereo.cc" ]; then
  g++ $CXXFLAGS -c "$REF/src/flame/stereo/inverse_depth_filter.cc" -o "$OBJ/inverse_depth_filter.o"
  g++ $CXXFLAGS -c "$REF/src/flame/stereo/inverse_depth_meas_model.cc" -o "$OBJ/inverse_depth_meas_model.o"
  g++ $CXXFLAGS -c "$HERE/ref_capi/ref_stereo.cc" -o "$OBJ/ref_stereo.o"
  OBJS="$OBJS $OBJ/inverse_depth_filter.o $OBJ/inverse_depth_meas_model.o $OBJ/ref_stereo.o"
fi
if [ -f "$HERE/ref_capi/ref_filters.cc" ]; then
  # display filters + vertex normals: static members sliced verbatim out of flame.cc (which as a whole
  # cannot be compiled here) into oracle/_ref/gen/
  mkdir -p "$OUT/gen"
  python3 "$HERE/ref_capi/slice_flame.py" "$REF" "$OUT/gen"
  for f in ref_filters ref_project ref_detect; do
    if [ -f "$HERE/ref_capi/$f.cc" ]; then
      g++ $CXXFLAGS -I"$OUT/gen" -c "$HERE/ref_capi/$f.cc" -o "$OBJ/$f.o"
      OBJS="$OBJS $OBJ/$f.o"
    fi
  done
fi
g++ -shared -o "$OUT/libflame_ref.so" $OBJS -lm -lpthread
echo "built $OUT/libflame_ref.so"

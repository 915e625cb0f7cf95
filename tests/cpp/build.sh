#!/bin/bash
# Builds tests/cpp/adapter_test (needs /root/reference for the reference's headers and oracle/_ref/obj).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference}
[ -d "$REF/src/flame" ] || { echo "adapter_test: $REF absent, keeping prebuilt binary"; exit 0; }
[ -f "$ROOT/oracle/_ref/obj/regularizer.o" ] || "$ROOT/oracle/build_ref.sh" "$REF"
g++ -O2 -std=c++11 -ffp-contract=off -w -I"$ROOT/oracle/shim" -I"$REF/src" -I"$ROOT/include" \
  "$HERE/adapter_test.cc" "$ROOT/oracle/_ref/obj/regularizer.o" "$ROOT/oracle/_ref/obj/rasterization.o" \
  -L"$ROOT/flame_b200/lib" -lflame_b200 -Wl,-rpath,'$ORIGIN/../../flame_b200/lib' -o "$HERE/adapter_test"
echo "built $HERE/adapter_test"

#!/bin/bash
# Builds sdpl_slam_b200/lib/libsdpl_frontend.so for sm_100a (B200) in-tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --fmad=false -Xptxas -v"
SRCS=$(ls "$HERE"/*.cu)
OBJS=""
pids=""
for s in $SRCS; do
  o="$HERE/build/$(basename "$s" .cu).o"
  OBJS="$OBJS $o"
  stale=0
  for d in "$s" "$HERE"/*.cuh "$HERE"/../../include/*.h "$HERE"/../../include/*.inc "$HERE/build.sh"; do
    if [ ! -f "$o" ] || [ "$d" -nt "$o" ]; then stale=1; fi
  done
  if [ $stale = 1 ]; then
    ( $NVCC $FLAGS -c "$s" -o "$o" > "$HERE/build/$(basename "$s" .cu).log" 2>&1 || { cat "$HERE/build/$(basename "$s" .cu).log"; exit 1; } ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libsdpl_frontend.so" $OBJS -lcudart
echo "built $OUT/libsdpl_frontend.so"

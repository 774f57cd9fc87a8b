#!/bin/bash
# Builds libbrk_b200.so (the C-ABI library) for sm_100a, in-tree. No torch, no libcuda link:
# driver entry points are resolved at run time so the library also loads on a GPU-less host.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p build
objs=""
pids=""
for f in *.cu; do
  o=build/${f%.cu}.o
  stale=0
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ ../../include/brk_b200.h -nt "$o" ]; then stale=1; fi
  for h in *.cuh; do [ "$h" -nt "$o" ] && stale=1; done
  if [ $stale = 1 ]; then
    $NVCC $FLAGS ${BRK_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
for p in $pids; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libbrk_b200.so $objs
echo "built $(cd ..; pwd)/libbrk_b200.so"

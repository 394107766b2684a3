#!/bin/bash
# Builds pyemma_b200/libb2k.so for sm_100a (in-tree, so the .so travels with gpurun snapshots).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libb2k.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ --expt-relaxed-constexpr"
mkdir -p "$HERE/obj"
pids=()
for f in api exact lloyd rmsd kmpp regspace screen dtraj project prune; do
  if [ ! -f "$HERE/obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/obj/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/obj/$f.o" ] || [ "$HERE/kernels.h" -nt "$HERE/obj/$f.o" ] || [ "$HERE/../../include/b2k.h" -nt "$HERE/obj/$f.o" ]; then
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT" "$HERE"/obj/{api,exact,lloyd,rmsd,kmpp,regspace,screen,dtraj,project,prune}.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"

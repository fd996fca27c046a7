#!/bin/bash
# Builds libagym.so in-tree for sm_100a (B200).  Called by __graft_entry__.build().
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/../libagym.so"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static -I$ROOT/include -I$HERE"
mkdir -p "$HERE/obj"
pids=()
SRCS="agym_api agym_sim agym_staged agym_fit agym_fit_warp agym_fit_newton agym_bidfit agym_retain agym_nccl"
OBJS=""
for f in $SRCS; do
  OBJS="$OBJS $HERE/obj/$f.o"
  $NVCC $FLAGS ${AGYM_PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -ldl
echo "built $OUT"

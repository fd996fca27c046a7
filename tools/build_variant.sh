#!/bin/bash
# A/B builds of the library: tools/build_variant.sh NAME "-DAGYM_SIM_MINB=2 ..." -> build_variants/libagym_NAME.so
# (use with AGYM_LIB_PATH=build_variants/libagym_NAME.so; the shipped library is auction_gym_b200/libagym.so)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
HERE="$ROOT/auction_gym_b200/csrc"
NAME="$1"; DEFS="$2"
OBJ="$ROOT/build_variants/obj_$NAME"; mkdir -p "$OBJ"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static -I$ROOT/include -I$HERE $DEFS"
pids=(); OBJS=""
for f in agym_api agym_sim agym_staged agym_fit agym_fit_warp agym_fit_newton agym_bidfit agym_retain agym_nccl; do
  OBJS="$OBJS $OBJ/$f.o"
  /usr/local/cuda/bin/nvcc $FLAGS -c "$HERE/$f.cu" -o "$OBJ/$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$ROOT/build_variants/libagym_$NAME.so" $OBJS -ldl
echo "built build_variants/libagym_$NAME.so"

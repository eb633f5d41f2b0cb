#!/usr/bin/env bash
# Build libtsw_sm100.so in-tree (sm_100a only). Usage: build.sh [-j N]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libtsw_sm100.so
OBJ=../../build/obj
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
pids=()
for f in errors logmel ops heads gemm gemm_simt gemm_tc fmha decode ${TSW_EXTRA_SRCS:-}; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$f.cu" -nt "$OBJ/$f.o" ] || [ common.cuh -nt "$OBJ/$f.o" ] || [ gemm_common.cuh -nt "$OBJ/$f.o" ] || [ tc_ptx.cuh -nt "$OBJ/$f.o" ] || [ ../../include/tsw.h -nt "$OBJ/$f.o" ]; then
    nvcc $FLAGS -c "$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/*.o
echo "built $(realpath $OUT)"

#!/bin/bash
# Builds libvmgb.so (sm_100a only) in-tree: vmg_b200/libvmgb.so.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static"
SRCS="api.cu conv_generic.cu conv_tc.cu elementwise.cu warp.cu ltam.cu evalpath.cu resblock_chain.cu tab_mix.cu"
mkdir -p build
OBJS=""
PIDS=""
for s in $SRCS; do
  [ -f "$s" ] || continue
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_ptx.cuh -nt "$o" ] || [ ../../include/vmgb.h -nt "$o" ]; then
    echo "nvcc $s"
    $NVCC $FLAGS ${VMGB_PTXAS_V:+-Xptxas -v} -c "$s" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for pid in $PIDS; do wait "$pid" || { echo "nvcc failed" >&2; rm -f ../libvmgb.so; exit 1; }; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o ../libvmgb.so $OBJS -ldl -lpthread -lrt
echo "built $(cd .. && pwd)/libvmgb.so"

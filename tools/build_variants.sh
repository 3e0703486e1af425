#!/bin/bash
# Builds register-budget variants of the path tracer for A/B runs (tools/gpu_sweep3.py): build/variants/libsvr_t<threads>_<minblocks>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  t=${v%_*}; mb=${v#*_}
  ( nvcc -std=c++17 -O3 -use_fast_math -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -cudart shared \
      -DSVR_PT_MAX_THREADS=$t -DSVR_PT_MIN_BLOCKS=$mb -c sunvolumerender_b200/csrc/svr_pathtrace.cu -o build/variants/pt_t${t}_$mb.o 2> build/variants/pt_t${t}_$mb.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o build/variants/libsvr_t${t}_$mb.so build/obj/svr_api.o build/obj/svr_macrocell.o build/obj/svr_raycast.o build/obj/svr_volume_io.o build/obj/svr_tf_io.o build/obj/svr_env_io.o build/variants/pt_t${t}_$mb.o -lz ) &
done
wait
for v in "$@"; do
  t=${v%_*}; mb=${v#*_}
  echo -n "t${t}_$mb: "; grep -A2 "pathtrace_warp_kernelILi2ELb0" build/variants/pt_t${t}_$mb.log | grep -E "Used|spill" | tr '\n' ' '; echo
done

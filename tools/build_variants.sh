#!/bin/bash
# Builds A/B variants of the path tracer for tools/gpu_sweep3.py: build/variants/libsvr_<name>.so
#   tools/build_variants.sh name1:"-DSVR_VAR_EXITFMA=1" name2:"-DSVR_PT_MAX_THREADS=128 -DSVR_PT_MIN_BLOCKS=6" ...
set -e
cd "$(dirname "$0")/.."
make lib > /dev/null
mkdir -p build/variants
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  ( nvcc -std=c++17 -O3 -use_fast_math -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -cudart shared \
      $flags -c sunvolumerender_b200/csrc/svr_pathtrace.cu -o build/variants/pt_$name.o 2> build/variants/pt_$name.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o build/variants/libsvr_$name.so build/obj/svr_api.o build/obj/svr_macrocell.o build/obj/svr_raycast.o build/obj/svr_volume_io.o build/obj/svr_tf_io.o build/obj/svr_env_io.o build/obj/svr_canvas.o build/variants/pt_$name.o -lz ) &
done
wait
for v in "$@"; do
  name=${v%%:*}
  echo -n "$name: "; grep -A2 "pathtrace_profile_kernelILb0" build/variants/pt_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
done

set -x
timeout 900 python tools/gpu_sweep5.py --c4 --c4big --opts "kernel=2,wp=2;kernel=2,wp=4" build/variants/libsvr_mb10.so build/variants/libsvr_mb12.so build/variants/libsvr_mb9q10.so build/variants/libsvr_mb9q.so > gpurun_out/r_sweep.log 2>&1
cat gpurun_out/r_sweep.log

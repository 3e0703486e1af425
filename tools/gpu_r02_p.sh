set -x
timeout 900 python tools/gpu_sweep5.py --opts "kernel=2;kernel=2,wp=2;kernel=2,wp=3;kernel=2,wp=1;kernel=2,block=64" > gpurun_out/p_sweep.log 2>&1
timeout 900 python tools/gpu_sweep5.py --c4 --opts "kernel=2" build/variants/libsvr_cb1.so build/variants/libsvr_cb2.so build/variants/libsvr_cb3.so build/variants/libsvr_mb8.so >> gpurun_out/p_sweep.log 2>&1
cat gpurun_out/p_sweep.log

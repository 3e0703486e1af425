set -x
timeout 900 python tools/gpu_sweep5.py --c4 --c1 --opts "kernel=2;kernel=2,wp=2;kernel=2,wp=3" build/variants/libsvr_mb8q.so build/variants/libsvr_mb9q.so build/variants/libsvr_mb6.so > gpurun_out/q_sweep.log 2>&1
timeout 600 python tools/gpu_sweep5.py --c4 --c1 --noc3 --opts "kernel=2;kernel=2,wp=2" >> gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log

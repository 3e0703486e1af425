set -x
timeout 900 python tools/gpu_sweep5.py --c4 --c4spp128 --noc3 --opts "kernel=2;kernel=5;kernel=5,refill=4;kernel=5,refill=16" > gpurun_out/l_sweep.log 2>&1
timeout 900 python tools/gpu_sweep5.py --c4 --c4spp128 --noc3 --opts "kernel=5;kernel=5,refill=16" build/variants/libsvr_p64.so build/variants/libsvr_p96.so build/variants/libsvr_p192.so build/variants/libsvr_p128b4.so >> gpurun_out/l_sweep.log 2>&1
timeout 600 python tools/gpu_sweep5.py --opts "kernel=2;kernel=5" >> gpurun_out/l_sweep.log 2>&1
cat gpurun_out/l_sweep.log

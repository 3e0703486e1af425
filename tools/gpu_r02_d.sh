set -x
timeout 900 python -m pytest tests/test_gpu_profile_kernel.py -m gpu -q --timeout=600 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/d_pytest.log
timeout 900 python tools/gpu_sweep5.py --opts "profile=0;profile=1;profile=1,refill=4;profile=1,refill=12;profile=1,refill=16;profile=1,refill=24" > gpurun_out/d_sweep.log 2>&1; echo "sweep rc=$?"
timeout 900 python tools/gpu_sweep5.py --opts "profile=0;profile=1;profile=1,refill=16" build/variants/libsvr_pb6.so build/variants/libsvr_pb5.so build/variants/libsvr_ph7.so >> gpurun_out/d_sweep.log 2>&1
cat gpurun_out/d_sweep.log

# round 2, call B: the new kernel (shape 4): tests, light-in-view diagnostics, timing sweep
set -x
timeout 900 python -m pytest tests/test_gpu_profile_kernel.py tests/test_gpu_lights_in_view.py -m gpu -q --timeout=600 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/b_pytest.log
timeout 600 python tools/gpu_lights_diag.py ring_of_small_disks front_facing > gpurun_out/b_diag.log 2>&1; echo "diag rc=$?"; cat gpurun_out/b_diag.log | tail -40
timeout 900 python tools/gpu_sweep5.py --c1 --opts "profile=0;profile=1;profile=1,refill=1;profile=1,refill=4;profile=1,refill=16;profile=1,wp=8;profile=1,block=64" > gpurun_out/b_sweep.log 2>&1; echo "sweep rc=$?"
timeout 900 python tools/gpu_sweep5.py --opts "profile=1;profile=1,refill=4" build/variants/libsvr_pb6.so build/variants/libsvr_pb5.so build/variants/libsvr_pb8.so >> gpurun_out/b_sweep.log 2>&1
cat gpurun_out/b_sweep.log

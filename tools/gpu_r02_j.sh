set -x
timeout 900 python -m pytest tests/test_gpu_pool_kernel.py tests/test_gpu_lights_in_view.py -m gpu -q --timeout=600 -k "pool or cull" > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/j_pytest.log
timeout 900 python tools/gpu_sweep5.py --c4 --opts "kernel=2;kernel=5;kernel=5,refill=4;kernel=5,refill=16" > gpurun_out/j_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/j_sweep.log

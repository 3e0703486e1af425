set -x
timeout 900 python tools/gpu_sweep5.py --c1 --opts "kernel=2,wp=2" "" build/variants/libsvr_eps.so > gpurun_out/t_sweep.log 2>&1
timeout 900 python tools/gpu_sweep5.py --c1 --opts "kernel=2,wp=2" >> gpurun_out/t_sweep.log 2>&1
cat gpurun_out/t_sweep.log
timeout 600 python -m pytest tests/test_gpu_pathtrace.py -m gpu -q --timeout=600 -x -k "not c4 and not c5" > gpurun_out/t_pytest.log 2>&1; tail -3 gpurun_out/t_pytest.log

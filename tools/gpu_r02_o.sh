set -x
timeout 900 python tools/gpu_sweep5.py --c1 --opts "kernel=2;kernel=2,wp=8;kernel=2,wp=16" build/variants/libsvr_base.so > gpurun_out/o_sweep.log 2>&1
timeout 900 python tools/gpu_sweep5.py --c1 --opts "kernel=2;kernel=2,wp=8;kernel=2,wp=16;kernel=2,wp=32" >> gpurun_out/o_sweep.log 2>&1
cat gpurun_out/o_sweep.log
timeout 600 python -m pytest tests/test_gpu_pathtrace.py -m gpu -q --timeout=600 -k "sample_parallel or twin_mode or toggles or deterministic" > gpurun_out/o_pytest.log 2>&1; tail -3 gpurun_out/o_pytest.log

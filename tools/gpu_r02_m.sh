set -x
timeout 900 python -m pytest tests/test_gpu_envlight.py -m gpu -q -s --timeout=600 > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "variance ratio|passed|failed|^E " gpurun_out/m_pytest.log | head -20
timeout 600 python tools/gpu_sweep5.py --c4 --c4spp128 --opts "kernel=2" > gpurun_out/m_sweep.log 2>&1; cat gpurun_out/m_sweep.log

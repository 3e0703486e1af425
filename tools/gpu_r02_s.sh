set -x
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s_bench.err
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s_pytest.log

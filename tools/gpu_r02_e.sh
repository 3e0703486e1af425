set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/e_pytest.log

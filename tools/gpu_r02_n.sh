set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/n_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/n_smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/n_ref.json 2> gpurun_out/n_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-raycast > gpurun_out/n_ncu1.log 2>&1; echo "ncu rc=$?"

# round 2, call A: GPU tests, smoke, baseline bench, ncu --set full of the C4 scatter-queue kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 300 python tools/prof_run.py pt C4 2 32 2 > gpurun_out/a_c4_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pathtrace_queue_kernel --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/r02_pt_c4_queue python tools/prof_run.py pt C4 2 32 2 > gpurun_out/a_ncu_c4.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/a_ncu_c4.log

set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/u_smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/u_ref.json 2> gpurun_out/u_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/u_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-raycast > gpurun_out/u_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pathtrace_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_bench_pt_c3_warp \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e --no-raycast > gpurun_out/u_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 python tools/prof_run.py pt C4 2 512 1 > gpurun_out/u_c4_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pathtrace_queue_kernel --launch-skip 0 --launch-count 1 -f \
    -o gpurun_out/r02_pt_c4_queue python tools/prof_run.py pt C4 2 128 1 > gpurun_out/u_ncu_c4.log 2>&1; echo "ncu3 rc=$?"

set -x
SVR_BENCH_PRINT_MAPS=1 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/f_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/f_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pathtrace_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_bench_pt_c3_warp \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e > gpurun_out/f_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/f_ncu.log

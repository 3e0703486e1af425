# ncu passes over the bench command (run under gpurun, one GPU): launch list, then one full capture of the dominant kernel
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pathtrace_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r01_bench_pt_c3_warp \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log

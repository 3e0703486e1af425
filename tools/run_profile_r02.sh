# Round 2: the commands behind profiles/r02/ (run under gpurun on one B200; every ncu pass after the same command exited 0 without ncu).
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash tools/run_profile_r02.sh'
# then, here:  python tools/ncu_summary.py gpurun_out/<report>.ncu-rep 45 > profiles/r02/<name>.summary.txt
#              python tools/ncu_issue.py gpurun_out/r02_bench_pt_c3_warp.ncu-rep profiles/r02 C3 256 "pathtrace_warp_kernel<2,0>"
#              python tools/ncu_issue.py gpurun_out/r02_pt_c4_queue.ncu-rep profiles/r02 C4 128 "pathtrace_queue_kernel<0>"
#              python tools/ncu_launches.py gpurun_out/r02_bench_launches.csv > profiles/r02/bench_c3_launches.summary.txt
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-raycast > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-raycast > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pathtrace_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_bench_pt_c3_warp \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e --no-raycast > gpurun_out/ncu2.log 2>&1
python tools/prof_run.py pt C4 2 128 1 > gpurun_out/c4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pathtrace_queue_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/r02_pt_c4_queue \
    python tools/prof_run.py pt C4 2 128 1 > gpurun_out/ncu3.log 2>&1
# the two restructurings that were measured and left as options (prof_run.py: what config mode spp reps shape)
SVR_PROFILE=1 python tools/prof_run.py pt C3 2 256 2 4 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pathtrace_profile_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_pt_c3_profile \
    python tools/prof_run.py pt C3 2 256 2 4 > gpurun_out/ncu4.log 2>&1
python tools/prof_run.py pt C4 2 32 2 5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pathtrace_pool_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_pt_c4_pool \
    python tools/prof_run.py pt C4 2 32 2 5 > gpurun_out/ncu5.log 2>&1

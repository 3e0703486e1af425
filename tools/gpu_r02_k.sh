set -x
timeout 300 python tools/prof_run.py pt C4 2 32 2 5 > gpurun_out/k_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pathtrace_pool_kernel --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/r02_pt_c4_pool_v0 python tools/prof_run.py pt C4 2 32 2 5 > gpurun_out/k_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/k_ncu.log

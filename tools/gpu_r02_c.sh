set -x
timeout 300 python tools/prof_run.py pt C3 2 256 2 > gpurun_out/c_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pathtrace_profile_kernel --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/r02_pt_c3_profile_v0 python tools/prof_run.py pt C3 2 256 2 > gpurun_out/c_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/c_ncu.log

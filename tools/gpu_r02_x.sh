set -x
timeout 900 python bench.py --workload C5 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-ref-cuda --no-raycast > gpurun_out/x_c5_n1.json 2> gpurun_out/x_c5_n1.err; echo "c5 rc=$?"; tail -3 gpurun_out/x_c5_n1.err
timeout 900 python bench.py --impl reference --workload C5 --spp 8 --steps 2 --warmup 1 > gpurun_out/x_c5_ref.json 2> gpurun_out/x_c5_ref.err; echo "c5ref rc=$?"; tail -3 gpurun_out/x_c5_ref.err

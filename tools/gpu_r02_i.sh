set -x
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_pathtrace.py::test_row_bands_assemble_to_the_single_launch_frame" -m gpu -q --timeout=600 > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --c5 --no-raycast > gpurun_out/i_bench2.json 2> gpurun_out/i_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/i_bench2.err

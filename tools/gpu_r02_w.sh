set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 --no-raycast --no-c5 --no-e2e > gpurun_out/w_bench8.json 2> gpurun_out/w_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/w_bench8.err

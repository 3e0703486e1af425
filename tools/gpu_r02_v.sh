set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/v_bench8.json 2> gpurun_out/v_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/v_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/v_bench4.json 2> gpurun_out/v_bench4.err; echo "bench4 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/v_bench2.json 2> gpurun_out/v_bench2.err; echo "bench2 rc=$?"

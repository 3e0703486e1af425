TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
$TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-fanout pcie > gpurun_out/bench_n8p.json 2> gpurun_out/bench_n8p.err; echo rc=$?
timeout 400 $TR --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --workload C5 --spp 128 > gpurun_out/bench_n8_c5.json 2> gpurun_out/bench_n8_c5.err; echo rc=$?
tail -c 400 gpurun_out/bench_n8_c5.err
python - <<'PY'
import json
for f in ("bench_n8","bench_n8p","bench_n8_c5"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["kernel_ms"], d["config"]["macrocell"])
    except Exception as e:
        print(f, "ERR", e)
PY

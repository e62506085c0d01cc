#!/bin/bash
# round-2 GPU check B (N GPUs of one box): peer-gather tests, then the N-GPU bench line (gather inside the step),
# default gather mode and the push-after-sweep mode for comparison
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_peer_gather_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2b_pytest_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_$N.json 2> gpurun_out/r2b_bench_$N.err
RXG_GATHER_MODE=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_${N}_push.json 2> gpurun_out/r2b_bench_${N}_push.err
tail -3 gpurun_out/r2b_pytest_$N.txt; python - <<PY
import json
for f in ("gpurun_out/r2b_bench_$N.json", "gpurun_out/r2b_bench_${N}_push.json"):
    try:
        j = json.load(open(f)); g = j["gather"]
        print(f, "full", round(g["ms_per_step_full"], 3), "replicated", round(g["replicated_cov"]["ms_per_step"], 3), "sweep", round(g["sweep_only"]["ms_per_step"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 gpurun_out/r2b_bench_$N.err | cut -c1-300

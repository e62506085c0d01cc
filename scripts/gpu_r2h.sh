#!/bin/bash
# N-GPU gather legs only (no e2e / cpu legs): default gather mode, then forced fused mode for the replicated variant
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2h_bench_$N.json 2> gpurun_out/r2h_bench_$N.err
python - <<PY
import json
for f in ("gpurun_out/r2h_bench_$N.json",):
    try:
        j = json.load(open(f)); g = j["gather"]
        print(f, "full", round(g["ms_per_step_full"], 3), "replicated", round(g["replicated_cov"]["ms_per_step"], 3), "sweep", round(g["sweep_only"]["ms_per_step"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 gpurun_out/r2h_bench_$N.err | cut -c1-300

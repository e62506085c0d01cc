#!/bin/bash
# N-GPU bench line only (no tests): the driver's scaling run in miniature
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo_$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2g_bench_$N.json 2> gpurun_out/r2g_bench_$N.err
head -c 600 gpurun_out/r2g_bench_$N.json; echo; tail -n 12 gpurun_out/r2g_bench_$N.err | cut -c1-300

#!/bin/bash
# round-2 GPU check B (N GPUs of one box): peer-gather tests, then the N-GPU bench line (gather inside the step)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_peer_gather_gpu.py -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r2b_pytest_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_$N.json 2> gpurun_out/r2b_bench_$N.err
tail -6 gpurun_out/r2b_pytest_$N.txt; head -c 4000 gpurun_out/r2b_bench_$N.json; tail -n 15 gpurun_out/r2b_bench_$N.err

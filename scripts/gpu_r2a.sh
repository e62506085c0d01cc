#!/bin/bash
# round-2 GPU check A: host topology, full GPU test suite, 1-GPU bench line
mkdir -p gpurun_out
( nproc; cat /sys/fs/cgroup/cpu.max; lscpu | head -40; numactl -H 2>/dev/null; nvidia-smi topo -m; free -g ) > gpurun_out/r2a_host.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -8 gpurun_out/r2a_pytest.txt; head -c 2500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err

#!/bin/bash
# round-2 GPU check A: host topology, full GPU test suite, 1-GPU bench lines (sweep variants)
mkdir -p gpurun_out
( nproc; cat /sys/fs/cgroup/cpu.max; lscpu | head -40; numactl -H 2>/dev/null; nvidia-smi topo -m; free -g ) > gpurun_out/r2a_host.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --sweep-variant 3 > gpurun_out/r2a_bench_v3.json 2> gpurun_out/r2a_bench_v3.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err

tail -12 gpurun_out/r2a_pytest.txt; head -c 1500 gpurun_out/r2a_bench.json; echo; head -c 900 gpurun_out/r2a_bench_v3.json; tail -n 3 gpurun_out/r2a_bench.err; tail -n 3 gpurun_out/r2a_bench_v3.err
python - <<'PY' > gpurun_out/r2a_hostfill.txt 2>&1
import sys, numpy as np, torch
sys.path.insert(0, '.')
import rxinfer_jl_b200 as rx
lib = rx._lib.load()
rows, batch = 16000, 65536
buf = torch.empty(rows * batch, dtype=torch.float32).pin_memory()
p = rx._lib.as_fp(buf.data_ptr())
for nt in (4, 8, 16, 32):
    print("host fill threads", nt, "GB/s", lib.rxg_selftest_host_fill_gbs(p, rows, batch, nt, 3))
PY
cat gpurun_out/r2a_hostfill.txt

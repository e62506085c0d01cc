#!/bin/bash
# round-2 GPU check C (1 GPU): peer / rules tests, configs[2] and configs[3] bench lines, e2e host-thread sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_peer_gather_gpu.py tests/test_rules_gpu.py tests/test_streaming_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2c_pytest.txt
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err
python - <<'PY' > gpurun_out/r2c_e2e_threads.txt 2>&1
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import rxinfer_jl_b200 as rx
import bench
ctx = rx.Context(0)
mod = bench.notebook_model_f32()
T, D, batch = 1000, 4, 65536
yh = torch.randn(T, D, batch).pin_memory()
mh = torch.empty(T, D, batch).pin_memory(); ch = torch.empty(T, D, D, batch).pin_memory()
kw = dict(A=mod["A"], B=mod["B"], P=mod["P"], Q=mod["Q"], m0=mod["m0"], S0=mod["S0"])
for nt in (16, 6, 8, 12, 24, 32, 16):
    ctx.set_option("host_threads", nt)
    ctx.lgssm(yh, **kw, smooth=True, out_mean=mh, out_cov=ch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        ctx.lgssm(yh, **kw, smooth=True, out_mean=mh, out_cov=ch)
    torch.cuda.synchronize()
    print("e2e host_threads", nt, "ms/step", (time.perf_counter() - t0) / 4 * 1e3, flush=True)
for ns in (4, 16, 32):
    ctx.set_option("host_threads", 16); ctx.set_option("host_slices", ns)
    ctx.lgssm(yh, **kw, smooth=True, out_mean=mh, out_cov=ch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        ctx.lgssm(yh, **kw, smooth=True, out_mean=mh, out_cov=ch)
    torch.cuda.synchronize()
    print("e2e host_slices", ns, "ms/step", (time.perf_counter() - t0) / 4 * 1e3, flush=True)
PY
tail -8 gpurun_out/r2c_pytest.txt; cat gpurun_out/r2c_e2e_threads.txt; head -c 1800 gpurun_out/r2c_bench_cfg2.json; echo; head -c 1500 gpurun_out/r2c_bench_cfg3.json; tail -n 3 gpurun_out/r2c_bench_cfg2.err gpurun_out/r2c_bench_cfg3.err

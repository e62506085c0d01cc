#!/bin/bash
# round-2 GPU check A: full GPU test suite, smoke(), 1-GPU bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench.err
tail -6 gpurun_out/r2a_pytest.txt; cat gpurun_out/r2a_smoke.txt | tail -6; head -c 900 gpurun_out/r2a_bench.json; echo; head -c 500 gpurun_out/r2a_bench_ref.json; tail -n 3 gpurun_out/r2a_bench.err

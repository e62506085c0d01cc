#!/bin/bash
# round-2 GPU check J (1 GPU): rule kernels incl. the large-d path, rule timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rules_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2j_pytest.txt
timeout 300 python bench_extra.py --which rules 2>gpurun_out/r2j_extra.err > gpurun_out/r2j_rules.jsonl
cat gpurun_out/r2j_pytest.txt; cut -c1-230 gpurun_out/r2j_rules.jsonl; tail -3 gpurun_out/r2j_extra.err

#!/bin/bash
# round-2 GPU check I (1 GPU): latent-AR tests, peer-gather tests (push variants), LAR timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lar.py tests/test_peer_gather_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2i_pytest.txt
timeout 300 python bench_extra.py --which round2 2>gpurun_out/r2i_extra.err | grep -i "latent AR" > gpurun_out/r2i_lar.jsonl
cat gpurun_out/r2i_pytest.txt; cat gpurun_out/r2i_lar.jsonl; tail -3 gpurun_out/r2i_extra.err

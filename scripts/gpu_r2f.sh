#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_peer_gather_gpu.py -m gpu -q -k "per_chain_and_masks or generic_allgather or shared_missing" 2>&1 | tail -15 > gpurun_out/r2f_pytest.txt
timeout 600 python bench_extra.py --which round2 > gpurun_out/r2f_extra.jsonl 2> gpurun_out/r2f_extra.err
tail -4 gpurun_out/r2f_pytest.txt; grep -E "generic|missing" gpurun_out/r2f_extra.jsonl | cut -c1-300; tail -n 3 gpurun_out/r2f_extra.err

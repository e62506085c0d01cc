#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_peer_gather_gpu.py tests/test_parity_gpu.py -m gpu -q -s -k "peer or shared_missing or large_state or general_shapes" 2>&1 | tail -40 > gpurun_out/r2e_pytest.txt
tail -6 gpurun_out/r2e_pytest.txt

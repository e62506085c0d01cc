#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_gather_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2f_pytest.txt
cat gpurun_out/r2f_pytest.txt

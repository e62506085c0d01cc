#!/bin/bash
mkdir -p gpurun_out
RXG_DEBUG_PEER=1 timeout 900 python -m pytest tests/test_peer_gather_gpu.py -m gpu -q -s -k "generic_allgather" 2>&1 | tail -80 > gpurun_out/r2f_pytest.txt
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "per_chain_and_masks" 2>&1 | tail -3
grep -n "Error\|test_peer_gather_gpu.py:[0-9]*" gpurun_out/r2f_pytest.txt | head -20

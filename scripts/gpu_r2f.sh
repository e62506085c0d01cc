#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_peer.py > gpurun_out/r2f_debug.txt 2>&1
cat gpurun_out/r2f_debug.txt | tail -20

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_rng_goldens.py tests/test_vmp_hgf_gpu.py -m gpu -q -s 2>&1 | tail -12 > gpurun_out/r2f_pytest.txt
cat gpurun_out/r2f_pytest.txt | cut -c1-300

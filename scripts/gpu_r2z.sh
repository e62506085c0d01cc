#!/bin/bash
# round-2 final GPU check (1 GPU): full GPU suite, smoke(), bench line + reference arm, extra timings, ncu of the new kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2z_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>> gpurun_out/r2z_bench.err
timeout 400 python bench_extra.py --which rules,round2 > gpurun_out/r2z_extra.jsonl 2> gpurun_out/r2z_extra.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_new_kernels_launches.csv \
    python profiles/r2_lar_profile_driver.py > gpurun_out/r2z_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lar_vmp_kernel -c 1 -f -o gpurun_out/r2z_lar \
    python profiles/r2_lar_profile_driver.py > gpurun_out/r2z_ncu2.log 2>&1
tail -6 gpurun_out/r2z_pytest.txt; tail -4 gpurun_out/r2z_smoke.txt; head -c 700 gpurun_out/r2z_bench.json; echo; head -c 400 gpurun_out/r2z_bench_ref.json; echo
tail -n 3 gpurun_out/r2z_bench.err; wc -l gpurun_out/r2z_extra.jsonl; tail -2 gpurun_out/r2z_extra.err; tail -2 gpurun_out/r2z_ncu1.log; tail -2 gpurun_out/r2z_ncu2.log

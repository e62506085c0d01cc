# Round-1 (final) profiling recipe -- run under gpurun from the repo root, one GPU:
#   gpurun --timeout 1500 -- 'bash profiles/r1_profile_commands.sh'
# Numbers printed by a run under ncu are never bench values; the bench lines come from the un-profiled runs below.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r1f_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1f_bench_reference.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 10 --warmup 3 --per-chain-path --no-e2e --no-cpu > gpurun_out/r1f_bench_perchain.json 2>> gpurun_out/r1f_bench.err
python bench_extra.py > gpurun_out/r1f_extra.jsonl 2>> gpurun_out/r1f_bench.err
# launch list of the bench command (share of each kernel in the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_a.log 2>&1
# full-set captures: dominant kernel of the bench, and the tensor-core kernels of the large-state family
ncu --set full --clock-control none --import-source on -k regex:lgssm_shared_kernel -s 3 -c 1 -o gpurun_out/r1f_shared python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lgssm_umma_sweep -s 2 -c 1 -o gpurun_out/r1f_umma_sweep python profiles/r1_umma_profile_driver.py > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:umma_ky_kernel -s 2 -c 1 -o gpurun_out/r1f_umma_ky python profiles/r1_umma_profile_driver.py > gpurun_out/ncu_d.log 2>&1
B=18944 ncu --set full --clock-control none -k regex:lgssm_umma_sweep -s 2 -c 1 -o gpurun_out/r1f_umma_sweep_148 python profiles/r1_umma_profile_driver.py > gpurun_out/ncu_e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1f_large_launches.csv python profiles/r1_umma_profile_driver.py > /dev/null 2>&1
for k in shared umma_sweep umma_ky umma_sweep_148; do python profiles/ncu_summary.py gpurun_out/r1f_$k.ncu-rep > gpurun_out/r1f_${k}_summary.txt 2>&1; done
rm -f gpurun_out/r1f_umma_ky.ncu-rep gpurun_out/r1f_umma_sweep_148.ncu-rep      # 64 MiB return limit
cat gpurun_out/r1f_bench.json; tail -3 gpurun_out/r1f_bench.err; ls -la gpurun_out | tail -25

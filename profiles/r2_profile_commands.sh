# Round-2 profiling recipe -- run under gpurun from the repo root, one GPU:
#   gpurun --timeout 2400 -- 'bash profiles/r2_profile_commands.sh'
# Numbers printed by a run under ncu are never bench values; the bench lines come from the un-profiled runs.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg2.json 2>> gpurun_out/r2_bench.err
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg3.json 2>> gpurun_out/r2_bench.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --sweep-variant 3 > gpurun_out/r2_bench_seg_variant.json 2>> gpurun_out/r2_bench.err
python bench_extra.py --which round2,per_chain,filter > gpurun_out/r2_extra.jsonl 2>> gpurun_out/r2_bench.err
timeout 600 python -m pytest tests/test_peer_gather_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_peer_pytest.txt
# launch list of the bench command (share of each kernel in the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_cfg2_launches.csv python bench.py --config 2 --steps 1 --warmup 3 > gpurun_out/ncu_a2.log 2>&1
# full-set captures
ncu --set full --clock-control none -k regex:lgssm_shared_kernel -s 3 -c 1 -o gpurun_out/r2_shared python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none -k regex:lgssm_seg_kernel -s 3 -c 1 -o gpurun_out/r2_seg python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --sweep-variant 3 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:hgf_filter_kernel -s 3 -c 1 -o gpurun_out/r2_hgf python bench.py --config 3 --steps 1 --warmup 3 > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none -k regex:broadcast_cov_kernel -s 3 -c 1 -o gpurun_out/r2_bcast python bench.py --config 2 --steps 1 --warmup 3 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none -k regex:lgssm_umma_sweep -s 3 -c 1 -o gpurun_out/r2_umma_sweep python bench.py --config 2 --steps 1 --warmup 3 > gpurun_out/ncu_f.log 2>&1
for k in shared seg hgf bcast umma_sweep; do python profiles/ncu_summary.py gpurun_out/r2_$k.ncu-rep > gpurun_out/r2_${k}_summary.txt 2>&1; done
rm -f gpurun_out/r2_bcast.ncu-rep gpurun_out/r2_umma_sweep.ncu-rep gpurun_out/r2_hgf.ncu-rep      # 64 MiB return limit
cat gpurun_out/r2_peer_pytest.txt; cat gpurun_out/r2_extra.jsonl | cut -c1-400; cat gpurun_out/r2_bench.json | head -c 1500; tail -3 gpurun_out/r2_bench.err; ls -la gpurun_out | grep r2_ | tail -25

# Round-1 profiling recipe (run under gpurun from the repo root; one GPU).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lgssm_shared_kernel -s 3 -c 1 -o gpurun_out/r1_final_shared python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none -k regex:gain_scan_kernel -s 3 -c 1 -o gpurun_out/r1_final_scan python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:lgssm_chain_kernel -s 3 -c 1 -o gpurun_out/r1_final_chain python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --per-chain-path > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none -k regex:hgf_filter_kernel -s 1 -c 1 -o gpurun_out/r1_final_hgf python bench_extra.py --which hgf > gpurun_out/ncu_e.log 2>&1
for k in shared scan chain hgf; do python profiles/ncu_summary.py gpurun_out/r1_final_$k.ncu-rep > gpurun_out/r1_final_${k}_summary.txt 2>&1; done
rm -f gpurun_out/r1_final_scan.ncu-rep gpurun_out/r1_final_chain.ncu-rep gpurun_out/r1_final_hgf.ncu-rep   # 64 MiB return limit: keep the dominant kernel's report only
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err
python bench.py --steps 10 --warmup 3 --per-chain-path --no-e2e --no-cpu > gpurun_out/r1_final_bench_perchain.json 2>> gpurun_out/r1_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_final_bench_reference.json 2>> gpurun_out/r1_final_bench.err
python bench_extra.py > gpurun_out/r1_final_extra.jsonl 2>> gpurun_out/r1_final_bench.err
tail -3 gpurun_out/r1_final_bench.err; ls -la gpurun_out | head -30

# One GPU box: the wide-batch (GEMM-shaped stage 1) evidence -- bench line, launch list, one full ncu capture.
set -x
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
ROWS=4000000 REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|rescore|select|prepare" -c 200 --csv --log-file gpurun_out/launches_gemm.csv python scripts/gemm_sweep.py 4096:100 > gpurun_out/ncu_list_gemm.log 2>&1
ROWS=4000000 REPS=1 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 5 -c 1 -o gpurun_out/prof_scan_gemm python scripts/gemm_sweep.py 4096:100 > gpurun_out/ncu_full_gemm.log 2>&1
ROWS=10000000 python scripts/gemm_sweep.py 1024:10 4096:10 4096:100 16384:100 2>&1 | tee gpurun_out/gemm_sweep10m.log | tail

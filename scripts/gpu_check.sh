# One GPU box: parity tests, the bench (both arms), an ncu launch list and full captures of the stage-1 kernels.
set -x
python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
if [ "$1" = "full" ]; then
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|gemm_|rescore|select|automerge|prepare" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --wide-steps 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/prof_scan_tc_b1 python bench.py --steps 3 --warmup 3 --skip batch64,wide,cpu > gpurun_out/ncu_full_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 5 -c 1 -o gpurun_out/prof_scan_gemm64_b64 python bench.py --steps 3 --warmup 3 --skip wide,cpu > gpurun_out/ncu_full_b64.log 2>&1
TT_NO_GEMM=1 ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 2 -c 1 -o gpurun_out/prof_scan_tc2_b64 python bench.py --steps 3 --warmup 3 --skip wide,cpu > gpurun_out/ncu_full_b64_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 5 -c 1 -o gpurun_out/prof_scan_gemm python bench.py --steps 3 --warmup 3 --wide-steps 1 --skip batch64,cpu > gpurun_out/ncu_full_gemm.log 2>&1
for f in prof_scan_tc_b1 prof_scan_tc2_b64 prof_scan_gemm64_b64 prof_scan_gemm; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page details --csv > gpurun_out/${f}_details.csv 2>/dev/null; done
fi

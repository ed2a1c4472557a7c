# One GPU box: parity tests, the bench (both arms), an ncu launch list and full captures of the stage-1 kernels.
set -x
python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
if [ "$1" = "full" ]; then
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|rescore|select|automerge|prepare" -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/prof_scan_tc_b1 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_tc2 -s 2 -c 1 -o gpurun_out/prof_scan_tc2_b64 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_b64.log 2>&1
fi

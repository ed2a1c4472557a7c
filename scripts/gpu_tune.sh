set -x
python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest3.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest3.log
for cfg in "dyn_ch2:" "static_ch2:TT_SCAN_STATIC=1" "dyn_ch1:TT_SCAN_CH=1" "dyn_ch4:TT_SCAN_CH=4" "dyn_ch1_s8:TT_SCAN_CH=1 TT_SCAN_STAGES=8" "dyn_ch2_s3:TT_SCAN_STAGES=3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench3_$name.json 2> gpurun_out/bench3_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench3_$name.json"))
    print("$name", "value", round(d["value"],1), "scan_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "b64", round(d["batch64"]["value"],1), "parity", d["parity_vs_gpu_exact_scan"], "certfail", d["certificate_failures"])
except Exception as e:
    print("$name FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|rescore|select|automerge|prepare" -c 200 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_list3.log 2>&1

#!/bin/bash
# compute-sanitizer over this package's kernels (namespace tt::) on the GPU box: memcheck, racecheck, synccheck.
# (initcheck is left out: with the kernel filter the writes of PyTorch's own kernels are not tracked, so every buffer torch
#  initialised reads as uninitialised.)
#   workload 1: __graft_entry__.smoke()  (scan tcgen05 + simt + GEMM-shaped, fused re-score/select/auto-merge)
#   workload 2: tests/test_gpu_exchange_one_device.py  (the peer-exchange protocol: push, flag wait, merge, barrier)
# Logs land in gpurun_out/sanitize_*.log; summarised into profiles/r02_sanitizer.md by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
    timeout 420 $CS --tool $tool --kernel-name kns=tt:: --launch-timeout 120 --log-file gpurun_out/sanitize_${tool}_smoke.log \
        python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.out 2>&1
    echo "$tool smoke rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_smoke.log 2>/dev/null) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_smoke.log | tail -n 1)"
done
for tool in memcheck racecheck; do
    timeout 420 $CS --tool $tool --kernel-name kns=tt:: --launch-timeout 120 --log-file gpurun_out/sanitize_${tool}_exchange.log \
        python -m pytest tests/test_gpu_exchange_one_device.py -m gpu -q -x -k "push_merge or second_round or fused_tail" --timeout 400 \
        > gpurun_out/sanitize_${tool}_exchange.out 2>&1
    echo "$tool exchange rc=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_exchange.log | tail -n 1)"
done

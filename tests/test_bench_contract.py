"""bench.py's output contract, as far as it can be checked without a GPU: the CPU arm (`--impl reference`) prints ONE
JSON line with the keys the driver reads, and the GPU arm refuses to run without a CUDA device instead of falling back."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--rows", "131072", "--steps", "3", "--warmup", "3", "--cpu-sample-rows", "65536")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["gpu_launches"] == 0


def test_reference_arm_uses_every_host_thread_even_under_a_launcher():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the all-cores baseline regardless."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "131072",
                        "--steps", "3", "--warmup", "3", "--cpu-sample-rows", "65536"], capture_output=True, text=True, cwd=ROOT,
                       env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_does_not_fall_back_to_the_cpu():
    import torch

    if torch.cuda.is_available():
        return  # the GPU arm itself is exercised on the GPU box (gpurun)
    p = _run("--steps", "1", "--rows", "4096", "--skip", "batch64,wide,cpu")
    assert p.returncode != 0 and p.stdout.strip() == ""


def test_gpu_arm_prints_one_json_line_with_the_contract_keys():
    """On a GPU box: the default arm on a small corpus -- ONE stdout line, every key the driver and the judge read."""
    import pytest
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    p = _run("--rows", "300000", "--steps", "20", "--warmup", "3", "--wide-batch", "512", "--wide-k", "10", "--cpu-sample-rows", "65536",
             "--c3-rows-per-gpu", "200000", "--c4-rows-per-gpu", "100000", "--c5-rows", "300000")
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["bytes_per_launch"] == 300000 * 1024 * 2
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 and e["d2h_bytes_per_step"] > 0 and e["callers"] == 1
    assert e["two_callers"]["value"] > 0 and e["two_callers"]["callers"] == 2 and e["two_callers"]["equals_single_caller_answer"]
    assert e["eight_callers"]["value"] > 0 and e["eight_callers"]["callers"] == 8 and e["eight_callers"]["equals_single_caller_answer"]
    assert d["serial_graph"]["ms_per_step"] > 0 and 0 < r["step_share_graph"] <= 1.05
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    # prepare, scan, re-score + select + auto-merge: three kernels per step on one GPU
    assert d["gpu_launches"] == 20 * 3 and d["kernels_per_step"] == 3 and "workload" in d["config"] and "l2" in d["config"]
    assert d["clocks"]["sm_max_mhz"] and isinstance(d["clocks"]["reasons"], list)
    assert d["parity_vs_gpu_exact_scan"] is True and d["certificate_failures"] == 0
    w = d["wide"]
    assert w["roofline"]["bound"] == "tensor" and w["gemm_path"] and w["parity_vs_gpu_exact_scan_local_shard"] is True
    # the other BASELINE configs ride in the same line, each with its own roofline object
    assert d["c3"]["roofline"]["bound"] == "hbm" and d["c3"]["scaling"] == "weak" and d["c3"]["rows_per_gpu"] == 200000
    assert d["c4"]["roofline"]["bound"] == "tensor" and d["c4"]["rows_per_gpu"] == 100000
    assert d["c5"]["k"] == 200 and d["c5"]["roofline"]["bound"] == "hbm" and d["c5"]["certificate_failures"] == 0
    # the fp32 store's eps is measured from the stored rows (index._shadow_gap): above the bf16 store's, below the budgeted 4.2e-3
    assert 2.5e-4 < d["fp32_store"]["eps"] <= 4.2e-3 and d["fp32_store"]["certificate_failures"] == 0
    h = d["hard_queries"]
    assert h["deep_rung"]["matches_exact_scan"] and h["exact_fallback"]["matches_exact_scan"]
    assert h["deep_rung"]["deep_rescans_kprime128"] > 0 and h["exact_fallback"]["exact_fp64_fallbacks"] > 0
    assert d["parity_vs_cpu_oracle"]["ok"] is True and d["parity_vs_cpu_oracle"]["queries_checked"] == 64


test_gpu_arm_prints_one_json_line_with_the_contract_keys = __import__("pytest").mark.gpu(
    test_gpu_arm_prints_one_json_line_with_the_contract_keys)

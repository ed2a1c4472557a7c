"""Concurrent retrieve_host callers on one index: throughput, coalesced batch sizes, per-batch latency.  GPU box."""
import collections
import faulthandler
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus

faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG_S", 150)), exit=True)  # a hang prints every thread's stack
n = int(os.environ.get("ROWS", 10_000_000))
callers = int(os.environ.get("CALLERS", 8))
per = int(os.environ.get("PER", 60))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(64, lookup=lambda t: corpus[t])).cpu()
idx = DeviceIndex(corpus, sc.tree, inv_norm=inv)
if os.environ.get("WAIT_TIMEOUT_MS"):  # bounded device-side waits: a stuck ring / exchange wait surfaces as a TTError
    from tensor_truth_b200 import _lib

    _lib.set_wait_timeout_ms(0, int(os.environ["WAIT_TIMEOUT_MS"]))
sizes, lat = collections.Counter(), collections.defaultdict(list)
inner = idx._retrieve_host_lane


def timed(lane, qh, *a):
    t0 = time.perf_counter()
    out = inner(lane, qh, *a)
    b = int(qh.shape[0])
    sizes[b] += 1
    lat[b].append((time.perf_counter() - t0) * 1e3)
    return out


idx._retrieve_host_lane = timed


def worker(t, count):
    torch.cuda.set_device(0)
    for i in range(count):
        idx.retrieve_host(q[(7 * t + i) % 64:(7 * t + i) % 64 + 1], 10)


for phase, count in (("warm", per), ("timed", per)):
    sizes.clear()
    lat.clear()
    ths = [threading.Thread(target=worker, args=(t, count)) for t in range(callers)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    print(f"{phase}: {callers} callers, {callers * count / dt:.1f} q/s; batches {dict(sorted(sizes.items()))}; "
          f"median ms per batch size { {b: round(sorted(v)[len(v) // 2], 2) for b, v in sorted(lat.items())} }; coalesced {idx.coalesced}", flush=True)
graphs = {k: (v["graph"] is not None, v.get("failures", 0), v["calls"]) for k, v in idx._ws.items() if isinstance(k, tuple) and k and k[0] == "graph"}
print("graphs (captured, failures, calls):", graphs)

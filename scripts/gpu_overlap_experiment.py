"""Experiment (round 2): does running K sub-batches concurrently (one host thread and CUDA stream each) beat one batch of the same structures on one
stream? K1 is latency-bound (issue slots 35 % busy), K4 ALU-bound (76 %): on separate streams one sub-batch's K4 can use the slots another's K1 leaves idle.
usage: python scripts/gpu_overlap_experiment.py [nstruct]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import synth  # noqa: E402
from fpocket_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
raw = bench.make_batch(n, os.cpu_count() or 8)
structs = [synth.to_structure(s) for s in raw]
p = api.default_params()
api.init(0)
for K in (1, 2, 3, 4, 6):
    parts = [structs[i::K] for i in range(K)]
    batches = [api.Batch(q, p) for q in parts]
    for _ in range(2):
        th = [threading.Thread(target=b.run) for b in batches]
        [t.start() for t in th]
        [t.join() for t in th]
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        th = [threading.Thread(target=b.run) for b in batches]
        [t.start() for t in th]
        [t.join() for t in th]
        best = min(best, time.perf_counter() - t0)
    print("K=%d sub-batches: %.1f ms wall, %.0f structures/s; per-batch device ms %s" % (K, best * 1e3, n / best, [round(b.kernel_ms(7), 1) for b in batches]), flush=True)
    for b in batches:
        b.close()

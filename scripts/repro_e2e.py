"""stress fpk_detect_batch's chunked pipeline: varying chunk sizes / call sizes on a pool of synthetic structures"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, synth
from fpocket_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
raw = bench.make_batch(n, 8)
structs = [synth.to_structure(s) for s in raw]
p = api.default_params()
for call in ([0, n // 3], [0, n], [n // 2, n], [0, n]):
    sub = structs[call[0]:call[1]]
    res, free = api.detect_batch_raw(sub, p)
    print(call, collections.Counter(res[k].status for k in range(len(sub))), flush=True)
    free()
print("ok")

"""Round 2: what the non-default linkages (-C m | a | c, K3L k_linkage) cost on BASELINE configs[1]-type structures, beside single linkage (-C s, K3),
through the staged C ABI (inputs resident in HBM), and what the unmodified reference program needs for the same options on a few of the same files.
usage: python scripts/gpu_linkage_timing.py [nstruct] [n_reference_files]     -> one line per method, plus a JSON summary on the last line"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import synth  # noqa: E402
from fpocket_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
nref = int(sys.argv[2]) if len(sys.argv) > 2 else 2
raw = bench.make_batch(n, os.cpu_count() or 8)
structs = [synth.to_structure(s) for s in raw]
api.init(0)
out = {"structures": n, "methods": {}}
for method, measure in (("s", "e"), ("m", "e"), ("a", "e"), ("c", "e"), ("c", "b")):
    p = api.default_params()
    p.clustering_method, p.distance_measure = ord(method), ord(measure)
    b = api.Batch(structs, p)
    b.run()
    b.run()
    best, stage = 1e9, 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        b.run()
        dt = time.perf_counter() - t0
        if dt < best:
            best, stage = dt, b.kernel_ms(1)
    res = b.fetch(as_dicts=False)
    ok = sum(1 for r in res if r[0] == 0)
    b.close()
    key = "-C %s -e %s" % (method, measure)
    out["methods"][key] = {"structures_per_s": round(n / best, 1), "ms": round(best * 1e3, 1), "cluster_stage_ms": round(stage, 1), "ok": ok}
    print("%s: %.1f ms for %d structures = %.0f structures/s; clustering stage %.1f ms; %d ok" % (key, best * 1e3, n, n / best, stage, ok), flush=True)
ref = os.path.join(ROOT, "oracle", "_ref", "fpocket")
if nref > 0 and os.path.exists(ref):
    tmp = tempfile.mkdtemp(prefix="fpklink")
    for k in range(nref):
        synth.to_pdb(raw[k], os.path.join(tmp, "s%d.pdb" % k))
    for opts in ([], ["-C", "m"], ["-e", "b", "-C", "c"]):
        t0 = time.perf_counter()
        for k in range(nref):
            subprocess.run([ref, "-f", "s%d.pdb" % k] + opts, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = (time.perf_counter() - t0) / nref
        out.setdefault("reference_s_per_structure", {})[" ".join(opts) or "default"] = round(dt, 2)
        print("reference %s: %.2f s per structure (%d files, %d-%d atoms)" % (" ".join(opts) or "default", dt, nref, min(len(raw[k].xyz) for k in range(nref)), max(len(raw[k].xyz) for k in range(nref))), flush=True)
print(json.dumps(out))

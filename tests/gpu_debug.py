"""Manual GPU triage script (not a test): python tests/gpu_debug.py [case]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import golden_cases as G, oracle_lib as O
from fpocket_b200 import api, capi
case = sys.argv[1] if len(sys.argv) > 1 else "5WA6"
d = G.load(case); st, p = G.inputs(d, want_tets=1)
b = api.Batch([st], p); b.run(); print("counters: sub %d retries %d nbig/err %d" % (b.counter(8), b.counter(9), b.counter(10))); b.close()
t0 = time.time(); r = api.detect(st, p); print("detect %.3fs status %d sites %d tets %d spheres %d clusters %d pockets %d" % (time.time()-t0, r["status"], r["n_sites"], r["n_tets"], r["n_spheres"], r["n_clusters"], r["n_pockets"]))
q = O.canonicalize(d["qh.tets"]); print("tets equal qhull:", r["tets"] is not None and np.array_equal(r["tets"], q), len(q))
o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
print("oracle: spheres %d clusters %d pockets %d" % (o["n_spheres"], o["n_clusters"], o["n_pockets"]))
for k in ("sph_xyzr", "sph_bary", "sph_neigh", "sph_type", "sph_cluster", "cl_off", "cl_sph", "cl_atom_off", "cl_atoms", "rank_to_cluster", "atom_fx"):
    a, b = r[k], o[k]
    print("  %-16s %s" % (k, G.eq_nan(a, b) if a.shape == b.shape else ("shape", a.shape, b.shape)))
if r["n_clusters"] == o["n_clusters"]:
    for n in r["desc"].dtype.names:
        a, b = r["desc"][n], o["desc"][n]
        if not G.eq_nan(a, b):
            bad = np.where(~((a == b) | (np.isnan(a) & np.isnan(b))))[0]
            print("  DESC DIFF %-24s n=%d first %s gpu %s oracle %s" % (n, len(bad), bad[:3], a[bad[:3]], b[bad[:3]]))

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
from fpocket_b200 import api
s5 = synth.generate(20260301, 5000, 0.065)
st = synth.to_structure(s5)
p = api.default_params(); p.do_asa_and_volume = 0
b = api.Batch([st] * 1024, p)
b.run(); b.run()
names = ["k_delaunay_filter", "k_cluster", "host", "k_hull", "k_descriptors", "finalize+compact", "", "total"]
print({names[i]: round(b.kernel_ms(i), 2) for i in (0, 1, 2, 3, 4, 5, 7)})

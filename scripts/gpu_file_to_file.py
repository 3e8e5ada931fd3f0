"""file-to-file throughput of build/fpocket_cuda_batch (reference reader + CUDA detection + reference writers) on N synthetic PDBs"""
import os, sys, time, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
raw = bench.make_batch(n, 8)
tmp = tempfile.mkdtemp(prefix="f2f", dir="/dev/shm")
with open(os.path.join(tmp, "list.txt"), "w") as f:
    for k, s in enumerate(raw):
        synth.to_pdb(s, os.path.join(tmp, "s%05d.pdb" % k)); f.write("s%05d.pdb\n" % k)
t = time.time()
out = subprocess.check_output([os.path.join(ROOT, "build", "fpocket_cuda_batch"), "list.txt"], cwd=tmp, stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1]
dt = time.time() - t
print(out, "| %.1f s, %.1f structures/s file to file" % (dt, n / dt))
subprocess.call(["rm", "-rf", tmp])

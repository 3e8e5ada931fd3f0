"""file-to-file throughput on N synthetic PDB files (tmpfs): build/fpocket_fast (native reader / writers, fpocket_b200/hostio) in a sweep of
pipeline settings, beside build/fpocket_cuda_batch (the reference's reader / writers around the same library) on a small sample.
  python scripts/gpu_file_to_file.py [N] [--sweep]"""
import os, sys, time, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2000
raw = bench.make_batch(n, os.cpu_count() or 8)
tmp, paths, nbytes = bench.write_pdb_files(raw, n, os.cpu_count() or 8, "sweep")
with open(os.path.join(tmp, "list.txt"), "w") as f:
    f.write("".join(os.path.basename(p) + "\n" for p in paths))
with open(os.path.join(tmp, "list64.txt"), "w") as f:
    f.write("".join(os.path.basename(p) + "\n" for p in paths[:64]))
fast = os.path.join(ROOT, "build", "fpocket_fast")
configs = [({}, [])]
if "--devices" in sys.argv:
    dv = sys.argv[sys.argv.index("--devices") + 1]
    configs = [({}, []), ({}, ["--devices", dv])]
if "--sweep" in sys.argv:
    configs = [({"FPKIO_DETECT_THREADS": d}, ["--group", g, "--threads", t]) for d, g, t in
               [("1", "1024", "16"), ("2", "1024", "16"), ("2", "512", "16"), ("2", "2048", "16"), ("3", "1024", "16"), ("2", "1024", "12"), ("2", "1024", "24"),
                ("3", "512", "16"), ("2", "1024", "8")]]
    configs.append(({"FPKIO_DETECT_THREADS": "2", "FPK_CHUNK": "256"}, ["--group", "1024", "--threads", "16"]))
for env, opts in configs:
    subprocess.call("rm -rf %s/*_out" % tmp, shell=True)
    t = time.time()
    out = subprocess.check_output([fast, "-F", "list.txt", "--quiet", "--stats"] + opts, cwd=tmp, env=dict(os.environ, **env)).decode().strip().splitlines()[-1]
    dt = time.time() - t
    print(env, opts, "| process %.2f s |" % dt, out, flush=True)
bexe = os.path.join(ROOT, "build", "fpocket_cuda_batch")
if os.path.exists(bexe):
    subprocess.call("rm -rf %s/*_out" % tmp, shell=True)
    t = time.time()
    out = subprocess.check_output([bexe, "list64.txt"], cwd=tmp, stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1]
    dt = time.time() - t
    print("reference reader/writers:", out, "| %.1f s, %.1f structures/s file to file" % (dt, 64 / dt))
subprocess.call(["rm", "-rf", tmp])

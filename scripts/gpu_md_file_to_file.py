"""mdpocket file to file: a synthetic DCD trajectory (5,000-atom protein, config 3 of BASELINE.json) through build/mdpocket_cuda (the reference's
mdpocket program with mdpocket_detect() on the fpk_md_* API: molfile DCD reader -> blocks of frames -> GPU -> the reference's DX / PDB writers),
beside the unmodified reference on the first frames.   python scripts/gpu_md_file_to_file.py [frames] [--ref N]"""
import os, sys, time, subprocess, tempfile, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth


def write_dcd(path, xyz):
    """CHARMM DCD, little endian, no unit cell: what the molfile dcdplugin (and the reference's -t traj.dcd -O dcd) reads."""
    F, N, _ = xyz.shape
    with open(path, "wb") as f:
        icntrl = np.zeros(20, np.int32)
        icntrl[0], icntrl[1], icntrl[2], icntrl[3], icntrl[19] = F, 1, 1, F, 24
        hdr = b"CORD" + icntrl.tobytes()
        hdr = hdr[:40] + struct.pack("<f", 1.0) + hdr[44:]
        f.write(struct.pack("<i", 84) + hdr + struct.pack("<i", 84))
        title = b"synthetic trajectory".ljust(80)
        f.write(struct.pack("<i", 84) + struct.pack("<i", 1) + title + struct.pack("<i", 84))
        f.write(struct.pack("<i", 4) + struct.pack("<i", N) + struct.pack("<i", 4))
        m = struct.pack("<i", 4 * N)
        for k in range(F):
            for c in range(3):
                f.write(m + np.ascontiguousarray(xyz[k, :, c], np.float32).tobytes() + m)


def make_traj(F, seed=20260301):
    s = synth.generate(seed, 5000, 0.065)
    rng = np.random.default_rng(seed)
    modes = rng.normal(size=(8, s.n, 3)).astype(np.float32) * 0.25
    amp = np.cumsum(rng.normal(size=(F, 8)).astype(np.float32) * 0.15, 0)
    amp -= amp.mean(0)
    xyz = (s.xyz[None] + np.tensordot(np.clip(amp, -1.2, 1.2), modes, 1)).astype(np.float32)
    xyz[0] = s.xyz
    return s, np.round(xyz, 3)


if __name__ == "__main__":
    F = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    nref = int(sys.argv[sys.argv.index("--ref") + 1]) if "--ref" in sys.argv else 32
    s, xyz = make_traj(F)
    tmp = tempfile.mkdtemp(prefix="mdf2f", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    synth.to_pdb(s, os.path.join(tmp, "top.pdb"))
    write_dcd(os.path.join(tmp, "traj.dcd"), xyz)
    write_dcd(os.path.join(tmp, "ref.dcd"), xyz[:nref])
    ref = os.path.join(ROOT, "oracle", "_ref", "mdpocket")
    for prog, what in (("mdpocket_fast", "native DCD reader / writers"), ("mdpocket_cuda", "the reference's molfile reader / writers")):
        exe = os.path.join(ROOT, "build", prog)
        if not os.path.exists(exe):
            continue
        d2 = os.path.join(tmp, prog); os.makedirs(d2)
        t = time.time()
        subprocess.check_call([exe, "--trajectory_file", "../traj.dcd", "--trajectory_format", "dcd", "-f", "../top.pdb"], cwd=d2, stdout=subprocess.DEVNULL)
        dt = time.time() - t
        print("%s (%s): %d frames, %.2f s process wall (CUDA start-up, DCD reading, detection, grids, DX / PDB writing): %.0f frames/s" % (prog, what, F, dt, F / dt))
    if os.path.exists(ref) and nref > 0:
        d2 = os.path.join(tmp, "r"); os.makedirs(d2)
        t = time.time()
        subprocess.check_call([ref, "--trajectory_file", "../ref.dcd", "--trajectory_format", "dcd", "-f", "../top.pdb"], cwd=d2, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.time() - t
        print("reference mdpocket: %d frames, %.2f s: %.1f frames/s (one process: the frame loop is sequential)" % (nref, dt, nref / dt))
    subprocess.call(["rm", "-rf", tmp])

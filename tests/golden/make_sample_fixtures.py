"""Regenerates tests/golden/samples/*.npz: EVERY structure of the reference's data/sample (PDB and mmCIF), read by the reference's own
readers and run through the unmodified reference (oracle/_ref/ref_dump), default parameters. Build container only.

Kept per structure (compact: the GPU box has neither /root/reference nor an mmCIF reader):
  * the atoms exactly as the reference parsed them - pdb_w_lig's list plus a mask of the atoms that are also in pdb (pdb is a
    subsequence of pdb_w_lig on every sample; checked here);
  * what the reference produced: kept spheres, pre-drop pockets with descriptors, final pockets, per-atom ASA effects;
  * of qhull's tetrahedra only the tail the reference never read (voronoi.c:138-155), which explains the spheres it lacks.
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dumpio import read_dump  # noqa: E402

REF = os.environ.get("FPK_REFERENCE", "/root/reference")
DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
OUT = os.path.join(HERE, "samples")


def one(fn):
    name = fn.replace(".", "_")
    tmp = tempfile.mkdtemp(prefix="fpksmp")
    loc = os.path.join(tmp, fn)
    shutil.copy(os.path.join(REF, "data", "sample", fn), loc)
    os.chmod(loc, 0o644)
    out = os.path.join(tmp, "d.dump")
    rc = subprocess.run([DUMP, out, "--seed", "77", "-f", loc], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp).returncode
    if rc != 0 or not os.path.exists(out):
        shutil.rmtree(tmp)
        return name, "ref_dump rc %d" % rc
    d = read_dump(out)
    shutil.rmtree(tmp)
    if "sph0.f" not in d:
        return name, "no spheres"
    # pdb as a subsequence of pdb_w_lig
    a = np.concatenate([d["pdb.f"].view(np.int32), d["pdb.i"]], 1)
    w = np.concatenate([d["pdbw.f"].view(np.int32), d["pdbw.i"]], 1)
    mask = np.zeros(len(w), np.uint8)
    j = 0
    for i in range(len(a)):
        while j < len(w) and not np.array_equal(w[j], a[i]):
            j += 1
        if j >= len(w):
            return name, "pdb is not a subsequence of pdb_w_lig"
        mask[j] = 1
        j += 1
    keep = {"md": d["md"], "params.f": d["params.f"], "params.i": d["params.i"], "in_pdb": mask, "seed": np.array([77])}
    for k in ("f", "i", "symbol", "res_name", "bfstat"):
        keep["pdbw." + k] = d["pdbw." + k]
    keep["pdb.bfstat"] = d["pdb.bfstat"]
    for k in ("sph0.f", "sph0.i", "pre.off", "pre.sph", "pre.desc_f", "pre.desc_i", "fin.off", "fin.sph", "fin.desc_f", "atom_fx", "status"):
        if k in d:
            keep[k] = d[k]
    seen = d["qh.tets_seen"]
    assert np.array_equal(seen[:-1], d["qh.tets"][:len(seen) - 1])
    keep["qh.n_tets"] = np.array([len(d["qh.tets"])])
    keep["qh.tail"] = d["qh.tets"][len(seen) - 1:].copy()          # from the (possibly cut) last visible line on
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **keep)
    return name, "%d atoms (%d with ligands), %d tets, %d spheres, %d pockets, tail %d tets" % (
        len(a), len(w), len(d["qh.tets"]), len(d["sph0.f"]), len(d["fin.off"]) - 1, len(keep["qh.tail"]))


def main():
    os.makedirs(OUT, exist_ok=True)
    files = sorted(f for f in os.listdir(os.path.join(REF, "data", "sample")) if f.endswith((".pdb", ".cif")))
    only = set(sys.argv[1:])
    if only:
        files = [f for f in files if f in only]
    with ThreadPoolExecutor(8) as ex:
        for name, msg in ex.map(one, files):
            print(name, msg, flush=True)


if __name__ == "__main__":
    main()

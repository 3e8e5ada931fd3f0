"""Regenerates tests/golden/*.npz. Runs only where /root/reference and oracle/_ref/ref_dump exist
(the build container); the GPU box uses the committed .npz files.

Each fixture is the full stage dump of the UNMODIFIED reference (oracle/ref_dump.c linked against the
reference objects) on one of the reference's own sample inputs / test cases
(reference tests/test_fpocket.py:65-132, data/sample/).
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dumpio import read_dump  # noqa: E402

REF = os.environ.get("FPK_REFERENCE", "/root/reference")
DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")

# name, input file (under data/sample), ref_dump flags, fpocket options
CASES = [
    ("1UYD", "1UYD.pdb", ["--seed", "11"], []),                                   # test_fpocket.py:65 default params
    ("3LKF", "3LKF.pdb", ["--seed", "12"], []),
    ("5WA6", "5WA6.pdb", ["--seed", "13"], []),
    ("7TAA_m28_M74", "7TAA.pdb", ["--seed", "14"], ["-m", "2.8", "-M", "7.4"]),  # test_fpocket.py:73
    ("1ATP_D36", "1ATP.pdb", ["--seed", "15"], ["-D", "3.6"]),                   # test_fpocket.py:95
    ("1UYD_explicit", "1UYD.pdb", ["--seed", "16"], ["-r", "1224:PU8:A"]),       # test_fpocket.py:117
    ("2P0R_mod_aD", "2P0R_mod.pdb", ["--seed", "17"], ["-a", "D"]),              # test_fpocket.py:125
    ("5RGF_cif", "5RGF.cif", ["--seed", "18"], []),                              # test_mmcif.py:77
    ("4URL", "4URL.pdb", ["--seed", "20"], []),                                   # test_fpocket.py:69 (the largest of the default list)
    ("2P0R_cD", "2P0R.pdb", ["--seed", "21"], ["-c", "D"]),                       # test_fpocket.py:111 drop a chain
    ("6X3P_cif", "6X3P.cif", ["--seed", "22"], []),                               # test_mmcif.py:82
    ("1QNH_cif_aCD", "1QNH.cif", ["--seed", "23"], ["-a", "C,D"]),                # test_mmcif.py:92 chains as ligand, mmCIF reader
    ("1UYD_dpocket", "1UYD.pdb", ["--seed", "19", "--lig", "PU8"], []),            # dpocket.c:186-290 with the known ligand
    # option branches of fparams.c:113-530 that the default cases never take (untested on the GPU)
    ("5WA6_i30_A2_v1000", "5WA6.pdb", ["--seed", "41"], ["-i", "30", "-A", "2", "-v", "1000"]),                     # -i, -A, -v
    ("5WA6_p01", "5WA6.pdb", ["--seed", "45"], ["-p", "0.1"]),            # -p: with refine.c:129-130's precedence every pocket that is not 100 % apolar is dropped
    ("3LKF_i8_A4", "3LKF.pdb", ["--seed", "42"], ["-i", "8", "-A", "4"]),                                         # min_pock_nb_asph below the default: the MC skip keys on it
    ("1UYD_P_u2", "1UYD.pdb", ["--seed", "43"], ["-P", "51:-:A.52:-:A.55:-:A.93:-:A.98:-:A.107:-:A.138:-:A.139:-:A.184:-:A", "-u", "2"]),   # explicit pocket by residues
    ("1UYD_P_u4", "1UYD.pdb", ["--seed", "44"], ["-P", "51:-:A.52:-:A.55:-:A.91:-:A.93:-:A.96:-:A.98:-:A.107:-:A.138:-:A.139:-:A.150:-:A.184:-:A"]),  # default -u 4
    ("md_2yex", "mdpocket/2yex.pdb", ["--md"], []),                           # mdpocket.c:833 call shape
    ("md_3ot3_S", "mdpocket/3ot3.pdb", ["--md", "--score"], []),              # -S drug-score increments
    # one snapshot of mdpocket's characterize mode (mdpocket.c:436-460): zone = the first 70 grid points of the reference's own mdpout_dens_iso_8.pdb
    ("alt_1UYD_eb", "1UYD.pdb", ["--seed", "32"], ["-e", "b"]),                  # test_fpocket.py:91 uses -e b (city-block clustering distance)
    # the other linkages of clusterlib.c:3705-3880,3337 behind -C (complete, average, centroid); test_fpocket.py:91 runs "-e b -C c"
    ("alt_1UYD_Cm", "1UYD.pdb", ["--seed", "33"], ["-C", "m"]),
    ("alt_1UYD_Ca", "1UYD.pdb", ["--seed", "34"], ["-C", "a"]),
    ("alt_1UYD_eb_Cc", "1UYD.pdb", ["--seed", "35"], ["-e", "b", "-C", "c"]),
    ("alt_3LKF_Cc", "3LKF.pdb", ["--seed", "36"], ["-C", "c"]),
    ("mdchar_3pa3", "mdpocket/3pa3.pdb", ["--md", "--seed", "31", "--want", "@zone70"], []),
]


def main():
    tmp = tempfile.mkdtemp(prefix="fpkgold")
    only = set(sys.argv[1:])            # optional: names of the fixtures to (re)generate
    for name, fn, dflags, opts in CASES:
        if only and name not in only:
            continue
        src = os.path.join(REF, "data", "sample", fn)
        if not os.path.exists(src):
            print("skip", name, "(no", src, ")")
            continue
        loc = os.path.join(tmp, os.path.basename(fn))
        shutil.copy(src, loc)
        os.chmod(loc, 0o644)
        out = os.path.join(tmp, name + ".dump")
        if "@zone70" in dflags:
            zone = os.path.join(tmp, "zone70.pdb")
            with open(os.path.join(REF, "data", "sample", "mdpocket", "mdpout_dens_iso_8.pdb")) as f:
                open(zone, "w").writelines(f.readlines()[:70])
            dflags = [zone if x == "@zone70" else x for x in dflags]
        subprocess.run([DUMP, out] + dflags + ["-f", loc] + opts, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, cwd=tmp)
        d = read_dump(out)
        d["seed"] = np.array([int(dflags[dflags.index("--seed") + 1]) if "--seed" in dflags else 0])
        for k in ("pdb.name", "pdb.type", "pdbw.name", "pdbw.type", "qh.centres", "sph1.f"):
            if not (k == "pdb.name" and "--lig" in dflags):          # atm_corsp compares atom names (atom.c:281)
                d.pop(k, None)
        # the reference reads only the flushed part of qhull's output (see oracle/ref_dump.c): keep the full
        # tet list, how many lines were visible and the (possibly cut) last visible line
        seen = d.pop("qh.tets_seen")
        assert np.array_equal(seen[:-1], d["qh.tets"][:len(seen) - 1])
        d["qh.n_seen"] = np.array([len(seen)])
        d["qh.last_seen"] = seen[-1].copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "sites", len(d["qh.tets"]), "tets; spheres", len(d["sph0.f"]), "pockets", len(d["fin.off"]) - 1)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()

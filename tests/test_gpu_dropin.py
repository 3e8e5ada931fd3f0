"""GPU: the reference's OWN fpocket program with one translation unit swapped - src/fpocket.c (search_pocket) replaced by
fpocket_b200/shim/search_pocket_cuda.c - built by fpocket_b200/shim/Makefile into build/fpocket_cuda (CLI, PDB reader and
writers are the reference's objects, unchanged). Checks
  * it runs and writes the reference's file set;
  * what it prints in <name>_info.txt IS the C-ABI result for the same structure (the shim maps every field correctly);
  * against the unmodified reference binary on the same file: same pockets up to the spheres the reference loses in its
    unflushed qhull tail (DESIGN.md par. 3) - pocket count within 2, and every top reference pocket has a CUDA twin sharing
    most of its atoms."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = O.ROOT
EXE = os.path.join(ROOT, "build", "fpocket_cuda")
REF = os.path.join(ROOT, "oracle", "_ref", "fpocket")
SEED = 424242


def _info(path):
    pockets, cur = [], None
    for line in open(path):
        m = re.match(r"Pocket (\d+) :", line)
        if m:
            cur = {}
            pockets.append(cur)
        elif ":" in line and cur is not None:
            k, v = line.rsplit(":", 1)
            cur[k.strip()] = float(v)
    return pockets


def _atoms(dirpath, k):
    ids = set()
    for line in open(os.path.join(dirpath, "pockets", "pocket%d_atm.pdb" % k)):
        if line.startswith("ATOM"):
            ids.add(int(line[6:11]))
    return ids


@pytest.mark.parametrize("n,opts", [(2300, []), (4100, []), (2300, ["-e", "b", "-C", "c"]), (2300, ["-C", "a"])])
def test_reference_program_with_cuda_search_pocket(tmp_path, n, opts):
    """opts: the default run, and the non-default clustering of the reference's own tests/test_fpocket.py:91 ("-e b -C c") parsed by the
    reference's fparams.c and carried through the shim into fpk_params."""
    if not os.path.exists(EXE):
        pytest.skip("build/fpocket_cuda not built (needs the reference tree: make -C fpocket_b200/shim)")
    import synth
    from fpocket_b200 import api
    s = synth.generate(20269400 + n, n, 0.065)
    dg, dr = tmp_path / "gpu", tmp_path / "ref"
    for d in (dg, dr):
        d.mkdir()
        synth.to_pdb(s, str(d / "x.pdb"))
    env = dict(os.environ, FPOCKET_MC_SEED=str(SEED), TMPDIR=str(tmp_path))
    subprocess.check_call([EXE, "-f", "x.pdb"] + opts, cwd=str(dg), env=env, stdout=subprocess.DEVNULL)
    out = dg / "x_out"
    assert (out / "x_info.txt").exists() and (out / "x_out.pdb").exists() and (out / "x_pockets.pqr").exists()
    got = _info(str(out / "x_info.txt"))
    # ---- the file is the C-ABI result ----------------------------------------------------------------------------------
    st = synth.to_structure(s)
    # fpmain.c:143-164 reads the file twice: pdb and pdb_w_lig are two objects with the same atoms here (asa.c:315 pointer test is always true)
    st.wx, st.wy, st.wz, st.wradius, st.welectroneg = st.x, st.y, st.z, st.radius, st.electroneg
    st.wflags = np.zeros(st.natoms, np.uint8)
    st.natoms_w, st.same_pdb = st.natoms, 0
    p = api.default_params()
    p.mc_seed = SEED
    for o, v in zip(opts[::2], opts[1::2]):
        setattr(p, {"-e": "distance_measure", "-C": "clustering_method"}[o], ord(v))
    r = api.detect(st, p)
    assert r["status"] == 0 and len(got) == r["n_pockets"]
    if opts:
        assert r["n_clusters"] != api.detect(st, api.default_params())["n_clusters"]
    D = r["desc"]
    fields = {"Score": "score", "Druggability Score": "drug_score", "Number of Alpha Spheres": "nb_asph", "Total SASA": "surf_vdw14",
              "Polar SASA": "surf_pol_vdw14", "Apolar SASA": "surf_apol_vdw14", "Volume": "volume", "Mean local hydrophobic density": "mean_loc_hyd_dens",
              "Mean alpha sphere radius": "mean_asph_ray", "Mean alp. sph. solvent access": "masph_sacc",
              "Apolar alpha sphere proportion": "apolar_asphere_prop", "Hydrophobicity score": "hydrophobicity_score", "Volume score": "volume_score",
              "Polarity score": "polarity_score", "Charge score": "charge_score", "Proportion of polar atoms": "prop_polar_atm",
              "Alpha sphere density": "as_density", "Cent. of mass - Alpha Sphere max dist": "as_max_dst", "Flexibility": "flex"}
    for k, c in enumerate(r["rank_to_cluster"]):
        for label, name in fields.items():
            assert abs(got[k][label] - float(D[name][c])) <= 5.1e-4 + 1e-6 * abs(float(D[name][c])), (k, label, got[k][label], D[name][c])
        atoms = {int(st.atom_id[a]) for a in r["cl_atoms"][r["cl_atom_off"][c]:r["cl_atom_off"][c + 1]]}
        assert _atoms(str(out), k + 1) == atoms
    # ---- against the unmodified reference on the same file -------------------------------------------------------------------
    if os.path.exists(REF):
        subprocess.check_call([REF, "-f", "x.pdb"] + opts, cwd=str(dr), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = _info(str(dr / "x_out" / "x_info.txt"))
        # Set equality modulo the explained tail: the reference loses the spheres of the last <= 4095 bytes of qhull's output
        # (voronoi.c:138-155; a handful of spheres, DESIGN.md par. 3), which can touch a pocket or two. Every other pocket must be OURS,
        # atom for atom; the touched ones must still be recognisable.
        assert abs(len(ref) - len(got)) <= 2
        ours = [frozenset(_atoms(str(out), k + 1)) for k in range(len(got))]
        theirs = [frozenset(_atoms(str(dr / "x_out"), k + 1)) for k in range(len(ref))]
        inexact = [a for a in theirs if a not in set(ours)]
        assert len(inexact) <= 3, (len(inexact), len(theirs))
        for a in inexact:
            best = max(len(a & b) / float(len(a | b)) for b in ours)
            assert best >= 0.6, best
        # the ranking of the pockets both runs have, in the reference's order, is ours
        common = [a for a in theirs if a in set(ours)]
        pos = [ours.index(a) for a in common]
        assert sum(1 for i in range(1, len(pos)) if pos[i] < pos[i - 1]) <= 2, pos


def test_batch_driver_writes_what_the_single_file_program_writes(tmp_path):
    """build/fpocket_cuda_batch (the -F list loop as ONE fpk_detect_batch call; reader and writers are the reference's) must leave the
    same <name>_out/<name>_info.txt and pocket atom files as build/fpocket_cuda -f <name> run file by file (same Monte-Carlo seed)."""
    bexe = os.path.join(ROOT, "build", "fpocket_cuda_batch")
    if not (os.path.exists(EXE) and os.path.exists(bexe)):
        pytest.skip("build/fpocket_cuda_batch not built (needs the reference tree: make -C fpocket_b200/shim)")
    import synth
    env = dict(os.environ, FPOCKET_MC_SEED=str(SEED), TMPDIR=str(tmp_path))
    da, db = tmp_path / "single", tmp_path / "batch"
    names = []
    for d in (da, db):
        d.mkdir()
    for i, n in enumerate((2050, 3100, 2500)):
        s = synth.generate(20269800 + i, n, 0.065)
        names.append("s%d" % i)
        for d in (da, db):
            synth.to_pdb(s, str(d / (names[-1] + ".pdb")))
    for nm in names:
        subprocess.check_call([EXE, "-f", nm + ".pdb"], cwd=str(da), env=env, stdout=subprocess.DEVNULL)
    (db / "list.txt").write_text("".join(nm + ".pdb\n" for nm in names))
    out = subprocess.check_output([bexe, "list.txt"], cwd=str(db), env=env).decode()
    assert "3 structures" in out and "0 failed" in out
    for nm in names:
        a = (da / (nm + "_out") / (nm + "_info.txt")).read_text()
        b = (db / (nm + "_out") / (nm + "_info.txt")).read_text()
        assert a == b and a.count("Pocket") >= 5
        k = a.count("Pocket ")
        for q in (1, k):
            fa = (da / (nm + "_out") / "pockets" / ("pocket%d_atm.pdb" % q)).read_text()
            fb = (db / (nm + "_out") / "pockets" / ("pocket%d_atm.pdb" % q)).read_text()
            assert fa == fb


def _read_dx(path):
    vals, dims, org = [], None, None
    for line in open(path):
        if line.startswith("object 1"):
            dims = [int(v) for v in line.split()[-3:]]
        elif line.startswith("origin"):
            org = [float(v) for v in line.split()[1:4]]
        elif line and (line[0].isdigit() or line[0] == "-") :
            vals += [float(v) for v in line.split()]
    return np.array(vals, np.float32).reshape(dims), np.array(org, np.float32)


def test_mdpocket_program_on_the_library(tmp_path):
    """build/mdpocket_cuda = the reference's mdpocket program with mdpocket_detect() replaced by the fpk_md_* driver
    (fpocket_b200/shim/mdpocket_cuda.c); reading of the snapshots, normalize_grid, project_grid_on_atoms and the DX / PDB writers are
    the reference's. Its DX grids must be the library's grids (fpk_md API called directly), and close to the unmodified reference's
    (which loses the spheres of its unflushed qhull tail, DESIGN.md par. 3)."""
    exe = os.path.join(ROOT, "build", "mdpocket_cuda")
    ref = os.path.join(ROOT, "oracle", "_ref", "mdpocket")
    if not os.path.exists(exe):
        pytest.skip("build/mdpocket_cuda not built (needs the reference tree: make -C fpocket_b200/shim)")
    import copy
    import synth
    from fpocket_b200 import api
    s = synth.generate(20269900, 2400, 0.065)
    rng = np.random.default_rng(2)
    frames = [s.xyz] + [np.round(s.xyz + rng.normal(scale=0.2, size=s.xyz.shape).astype(np.float32), 3) for _ in range(5)]
    dg, dr = tmp_path / "gpu", tmp_path / "ref"
    for d in (dg, dr):
        d.mkdir()
        with open(d / "list.txt", "w") as f:
            for i, x in enumerate(frames):
                sf = copy.copy(s)
                sf.xyz = x
                synth.to_pdb(sf, str(d / ("f%02d.pdb" % i)))
                f.write(str(d / ("f%02d.pdb" % i)) + "\n")
    subprocess.check_call([exe, "--pdb_list", "list.txt"], cwd=str(dg), stdout=subprocess.DEVNULL)
    dens, org = _read_dx(str(dg / "mdpout_dens_grid.dx"))
    freq, _ = _read_dx(str(dg / "mdpout_freq_grid.dx"))
    # ---- the same through the C ABI ----------------------------------------------------------------------------------------
    st = synth.to_structure(s)
    m = api.MdPocket(st, api.default_params())
    xyz = np.stack(frames).astype(np.float32)
    o, dims = m.grid_from_frame(xyz[:1])
    m.set_grid(o, dims)
    m.push_frames(xyz)
    gd, gf = m.fetch_grids()
    m.close()
    assert list(dens.shape) == [int(v) for v in dims] and np.allclose(org, np.round(o, 2), atol=0.0051)
    assert np.allclose(dens, gd / len(frames), atol=5.1e-4) and np.allclose(freq, gf / len(frames), atol=5.1e-4)
    assert (dg / "mdpout_all_atom_pdensities.pdb").stat().st_size > 1000
    # ---- against the unmodified reference ---------------------------------------------------------------------------------------
    if os.path.exists(ref):
        subprocess.check_call([ref, "--pdb_list", "list.txt"], cwd=str(dr), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rd, ro = _read_dx(str(dr / "mdpout_dens_grid.dx"))
        if rd.shape == dens.shape and np.allclose(ro, org, atol=0.011):
            # the same stencils on the same spheres; what may differ: the spheres of the reference's lost tail (a few per frame) and the
            # once-per-pocket centre cell, which follows the pocket's LAST sphere (DESIGN.md par. 5)
            cc = np.corrcoef(rd.ravel(), dens.ravel())[0, 1]
            assert cc > 0.995, cc
            assert np.mean(np.abs(rd - dens) > 5.1e-4) < 0.01


def test_dpocket_program_on_the_library(tmp_path):
    """build/dpocket_cuda = the reference's dpocket program with desc_pocket() replaced (fpocket_b200/shim/dpocket_cuda.c): loading,
    write_out_fpocket and the three dpout_*.txt writers are the reference's; pockets, the explicit pocket's record, interface atoms
    and the per-pocket criteria come from ONE fpk_detect call with the ligand attached. The rows it prints must be the C-ABI result,
    and agree with the unmodified reference dpocket on what does not depend on the reference's order-dependent explicit-pocket
    subset: the pocket that holds the ligand and its overlap criteria."""
    exe = os.path.join(ROOT, "build", "dpocket_cuda")
    ref = os.path.join(ROOT, "oracle", "_ref", "dpocket")
    if not os.path.exists(exe):
        pytest.skip("build/dpocket_cuda not built (needs the reference tree: make -C fpocket_b200/shim)")
    import synth
    from fpocket_b200 import api
    s = synth.generate(20269950, 2700, 0.065)
    first = api.detect(synth.to_structure(s), api.default_params())
    lig = synth.make_ligand(np.random.default_rng(4), first["desc"]["bary"][first["rank_to_cluster"][0]])
    rows = {}
    for d, prog in ((tmp_path / "gpu", exe), (tmp_path / "ref", ref)):
        if not os.path.exists(prog):
            continue
        d.mkdir()
        synth.to_pdb(s, str(d / "c.pdb"), lig)
        (d / "list.txt").write_text("%s LIG\n" % (d / "c.pdb"))
        env = dict(os.environ, FPOCKET_MC_SEED=str(SEED), TMPDIR=str(tmp_path))
        subprocess.check_call([prog, "-f", "list.txt"], cwd=str(d), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows[prog] = {k: [l.split() for l in open(d / ("dpout_%s.txt" % k)).read().splitlines()[1:]] for k in ("explicitp", "fpocketp", "fpocketnp")}
    g = rows[exe]
    assert len(g["explicitp"]) == 1 and len(g["fpocketp"]) >= 1
    # ---- the rows are the C-ABI result -------------------------------------------------------------------------------------------
    st = synth.with_ligand(s, lig)
    p = api.default_params()
    p.mc_seed = SEED
    r = api.detect(st, p)
    assert len(g["fpocketp"]) + len(g["fpocketnp"]) == r["n_pockets"]
    e = g["explicitp"][0]
    assert int(e[11]) == len(r["xspheres"]) and abs(float(e[10]) - float(r["xdesc"]["volume"][0])) <= 0.0051      # nb_AS, pock_vol
    best = int(np.argmax(r["pk_ovlp"]))
    top = g["fpocketp"][0]
    assert abs(float(top[2]) - float(r["pk_ovlp"][best])) <= 0.0051 and abs(float(top[4]) - float(r["pk_dst"][best])) <= 0.0051
    # ---- the unmodified reference finds the same ligand pocket -----------------------------------------------------------------------
    if ref in rows:
        rt = rows[ref]["fpocketp"][0]
        assert abs(float(rt[2]) - float(top[2])) <= 15.0 and abs(float(rt[4]) - float(top[4])) <= 1.0 and rt[5] == top[5]

"""CPU: the native host I/O (fpocket_b200/hostio, include/fpocket_hostio.h) against the reference's OWN reader and
writers. The checker is oracle/_ref/ref_io (oracle/ref_io.c linked with the unmodified reference objects + the C shim):

  * reader  : fpk_structure arrays and every text field of both atom lists must be BIT-IDENTICAL to rpdb_open + rpdb_read +
              shim_job_prepare — on synthetic PDB files, on the reference's data/sample files, with -c / -k, and on files whose
              records were mutated (truncated, alt-locs, blank elements, over-long lines, CR, HETATM variants, early ENDMDL);
  * writers : the output directory must be BYTE-IDENTICAL (and carry the same file modes) to what the reference's write_out_fpocket
              writes for the same result; the result comes from the oracle (test infrastructure) so no GPU is needed;
  * pipeline: fpkio_run_list with the oracle as the detect callback leaves the same directories as the one-by-one calls.
"""
import ctypes as C
import os
import random
import re
import shutil
import subprocess

import numpy as np
import pytest

import dumpio
import oracle_lib as O
from fpocket_b200 import api, capi, hostio

ROOT = O.ROOT
REF_IO = os.path.join(ROOT, "oracle", "_ref", "ref_io")
SAMPLE = "/root/reference/data/sample"
needs_ref_io = pytest.mark.skipif(not os.path.exists(REF_IO), reason="oracle/_ref/ref_io not built (needs the reference tree)")
needs_sample = pytest.mark.skipif(not os.path.isdir(SAMPLE), reason="reference data/sample not present")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fpocket_hostio.h")).read()
    names = sorted(set(re.findall(r"\b(fpkio_[a-z_]+)\s*\(", hdr)) - {"fpkio_detect_fn", "fpkio_free_fn"})
    assert len(names) >= 10
    L = hostio.lib()
    for n in names:
        assert hasattr(L, n), n
    assert b"hostio" in L.fpkio_version()
    # no CUDA dependency of its own: the detector is handed in
    out = subprocess.check_output(["ldd", hostio.LIB_PATH]).decode()
    assert "cuda" not in out and "fpocket_cuda" not in out


def test_refusals_and_errors(tmp_path):
    L = hostio.lib()
    h = C.c_void_p()
    assert L.fpkio_read_pdb(str(tmp_path / "nope.pdb").encode(), None, C.byref(h)) == hostio.FPKIO_ERR_OPEN
    (tmp_path / "x.cif").write_text("data_x\n")
    assert L.fpkio_read_pdb(str(tmp_path / "x.cif").encode(), None, C.byref(h)) == hostio.FPKIO_ERR_UNSUPPORTED
    (tmp_path / "e.pdb").write_text("HEADER nothing\nHETATM    1  O   HOH A   1       1.000   2.000   3.000  1.00 10.00           O\nEND\n")
    assert L.fpkio_read_pdb(str(tmp_path / "e.pdb").encode(), None, C.byref(h)) == hostio.FPKIO_ERR_NO_ATOMS
    assert L.fpkio_write_output(None, b"x.pdb", None, None) == hostio.FPKIO_ERR_BAD_ARG
    # fopen_pdb_check_case (utils.c:704-727): a code given in the wrong case is found in upper, then lower case
    (tmp_path / "1ABC.pdb").write_text("ATOM      1  N   ALA A   1       1.000   2.000   3.000  1.00 10.00           N\nEND\n")
    assert L.fpkio_read_pdb(str(tmp_path / "1abc.pdb").encode(), None, C.byref(h)) == hostio.FPKIO_OK and L.fpkio_natoms(h, 0) == 1
    L.fpkio_file_free(h)


# ---- reader -----------------------------------------------------------------------------------------------------------------------------
def _reader_equal(path, extra=(), opts=None):
    dump = path + ".dump"
    r = subprocess.run([REF_IO, "atoms", dump, "-f", path, *extra], capture_output=True)
    try:
        p = hostio.PdbFile(path, opts=opts)
    except IOError:
        p = None
    if r.returncode != 0 or p is None:
        assert (r.returncode != 0) == (p is None), (path, r.returncode, r.stderr[-200:])
        return 0
    d = dumpio.read_dump(dump)
    a = p.arrays()
    assert a["n"] == int(d["n"][0]) and a["nw"] == int(d["nw"][0]) and a["same_pdb"] == int(d["same_pdb"][0])
    for k in ("x", "y", "z", "radius", "electroneg", "bfactor", "bfstat", "wx", "wy", "wz", "wradius", "welectroneg"):
        assert np.array_equal(a[k].view(np.uint32), d[k].view(np.uint32)), (path, k)          # bit for bit, NaN-safe
    for k in ("atom_id", "res_id", "flags", "wflags"):
        assert np.array_equal(a[k], d[k]), (path, k)
    assert np.array_equal(a["aa_idx"].view(np.uint8), d["aa_idx"]), path
    for wl, pfx in ((0, "pdb"), (1, "pdbw")):
        t, iv = p.atom_text(wl)
        assert np.array_equal(iv, d[pfx + ".irc"]), (path, pfx)
        assert np.array_equal(t, d[pfx + ".text"]), (path, pfx)
        assert p.nhetatm(wl) == int(d[pfx + ".nhetatm"][0])
    st = p.structure
    xl = np.ctypeslib.as_array(st.xlig, (3 * st.n_xlig,)).copy() if st.n_xlig else np.zeros(0, np.float32)
    assert np.array_equal(xl.view(np.uint32), d["xlig"].view(np.uint32)), (path, "xlig", st.n_xlig, len(d["xlig"]))
    p.close()
    return a["n"]


@needs_ref_io
def test_reader_equals_reference_on_synthetic_files(tmp_path):
    import synth
    s = synth.generate(20268001, 900, 0.065)
    rng = np.random.default_rng(3)
    lig = s.xyz[:12] + rng.normal(scale=4.0, size=(12, 3)).astype(np.float32)
    synth.to_pdb(s, str(tmp_path / "a.pdb"))
    synth.to_pdb(s, str(tmp_path / "b.pdb"), lig_xyz=lig)
    assert _reader_equal(str(tmp_path / "a.pdb")) == s.n
    assert _reader_equal(str(tmp_path / "b.pdb")) == s.n


@needs_ref_io
@needs_sample
@pytest.mark.parametrize("name", ["1UYD", "1ATP", "5WA6", "1g50", "4bdf", "6x3p", "1UYD_wrote"])
def test_reader_equals_reference_on_sample_files(tmp_path, name):
    dst = str(tmp_path / (name + ".pdb"))
    shutil.copy(os.path.join(SAMPLE, name + ".pdb"), dst)
    assert _reader_equal(dst) > 400


@needs_ref_io
@needs_sample
def test_reader_chain_filters(tmp_path):
    dst = str(tmp_path / "1g50.pdb")
    shutil.copy(os.path.join(SAMPLE, "1g50.pdb"), dst)
    n_all = _reader_equal(dst)
    n_c = _reader_equal(dst, ("-c", "A"), hostio.ReadOpts.make(drop=["A"]))
    n_k = _reader_equal(dst, ("-k", "B,C"), hostio.ReadOpts.make(keep=["B", "C"]))
    assert 0 < n_c < n_all and 0 < n_k < n_all


@needs_ref_io
@needs_sample
def test_reader_explicit_ligand_and_explicit_pocket(tmp_path):
    """-r number:name:chain collects the coordinates of that residue's records (pdb->xlig_*), -P res:ins:chain.res:ins:chain flags the atoms of
    those residues (FPK_ATOM_XPOCKET); both arguments are parsed with the reference's quirks (empty fields vanish, the -P switch falls through)."""
    dst = str(tmp_path / "1UYD.pdb")
    shutil.copy(os.path.join(SAMPLE, "1UYD.pdb"), dst)
    o = hostio.ReadOpts.make(custom_ligand="1224:PU8:A")
    assert o.has_xlig == 1 and o.xlig_resnumber == 1224 and o.xlig_chain == b"A"
    assert _reader_equal(dst, ("-r", "1224:PU8:A"), o) > 1000
    p = hostio.PdbFile(dst, opts=o)
    assert p.structure.n_xlig > 10
    p.close()
    for spec, nflag in (("107:-:A.108:-:A.138:-:A", True), ("107::A.108::A", False)):        # "::" -> the insertion code becomes the chain letter: no atom matches
        o = hostio.ReadOpts.make(custom_pocket=spec)
        assert _reader_equal(dst, ("-P", spec), o) > 1000
        p = hostio.PdbFile(dst, opts=o)
        nf = int((p.arrays()["flags"] & capi.FPK_ATOM_XPOCKET != 0).sum())
        assert (nf > 10) == nflag, (spec, nf)
        p.close()


@needs_ref_io
@needs_sample
@pytest.mark.parametrize("name,chains", [("2P0R_mod", ["D"]), ("1g50", ["B"]), ("1g50", ["A"]), ("3vi4", ["B", "D"])])
def test_reader_chain_as_ligand(tmp_path, name, chains):
    """-a: the reference's counting pass looks at the current record's chain, its storing pass at the previous record's (rpdb.c:1079,1225), so
    one record goes astray at every boundary of a ligand chain and the declared atom count can differ from what was stored. The native
    reader reproduces both passes: same atoms (zero-filled slots included), same explicit-ligand coordinates, same B-factor statistics."""
    dst = str(tmp_path / (name + ".pdb"))
    shutil.copy(os.path.join(SAMPLE, name + ".pdb"), dst)
    o = hostio.ReadOpts.make(lig_chains=chains)
    n = _reader_equal(dst, ("-a", ",".join(chains)), o)
    p = hostio.PdbFile(dst, opts=o)
    assert n > 1000 and p.structure.n_xlig > 20 and p.natoms(1) >= n
    p.close()


@needs_ref_io
def test_reader_model_number(tmp_path):
    """-l N on a multi-model file: only the records of the N-th MODEL ... ENDMDL block; without -l the first ENDMDL ends the reading
    (its first three letters are END, rpdb.c:975)."""
    import synth
    s = synth.generate(20268005, 600, 0.065)
    synth.to_pdb(s, str(tmp_path / "m.pdb"))
    body = [ln for ln in open(tmp_path / "m.pdb").read().splitlines() if ln.startswith("ATOM")]
    txt = ""
    for m in range(3):
        txt += "MODEL     %4d\n" % (m + 1)
        for ln in body[: 300 + 100 * m]:
            txt += ln[:30] + "%8.3f" % (float(ln[30:38]) + 0.5 * m) + ln[38:] + "\n"
        txt += "ENDMDL\n"
    txt += "END\n"
    path = str(tmp_path / "multi.pdb")
    open(path, "w").write(txt)
    assert _reader_equal(path) == 300
    assert _reader_equal(path, ("-l", "2"), hostio.ReadOpts.make(model=2)) == 400
    assert _reader_equal(path, ("-l", "3"), hostio.ReadOpts.make(model=3)) == 500
    assert _reader_equal(str(tmp_path / "m.pdb"), ("-l", "2"), hostio.ReadOpts.make(model=2)) == s.n        # no MODEL record: no effect


@needs_ref_io
@needs_sample
def test_reader_on_mutated_records(tmp_path):
    """The reference parses fixed columns out of a line buffer it never clears: short records read what an earlier line left there.
    The native reader keeps the same buffer discipline, so even broken files give the same atoms."""
    rng = random.Random(7)
    src = open(os.path.join(SAMPLE, "1ATP.pdb")).read().split("\n")
    for trial in range(12):
        out = []
        for ln in src:
            if ln.startswith(("ATOM", "HETATM")) and rng.random() < 0.08:
                m = rng.randrange(12)
                if m == 0: ln = ln[:rng.randrange(27, 80)]
                elif m == 1: ln = ln[:76] + "  " + ln[78:]
                elif m == 2: ln = ln[:16] + rng.choice("AB C1") + ln[17:]
                elif m == 3: ln = ln + "   extra text beyond column 80"
                elif m == 4: ln = ln + "\r"
                elif m == 5: ln = ln[:30] + "9999.000" + ln[38:]
                elif m == 6: ln = ln[:78] + rng.choice(["1+", "2-", " 1", "+1"]) + ln[80:]
                elif m == 7: ln = "HETATM" + ln[6:17] + rng.choice(["HOH", "MSE", " ZN", "  A", "HEM", "LIG", "WAT", "TIP", " U "]) + ln[20:]
                elif m == 8: ln = ln[:12] + rng.choice(["1HB ", " CL ", "FE  ", " OXT", "CA  ", "HG21"]) + ln[16:76]
                elif m == 9: ln = ln[:54]
                elif m == 10: ln = ln[:17] + rng.choice(["  A", " DG", "  U", "HID"]) + ln[20:76]
                elif m == 11: ln = ln[:6] + "*****" + ln[11:22] + "   A" + ln[26:]
            out.append(ln)
        if trial % 4 == 0:
            out.insert(len(out) // 2, "ENDMDL")
        path = str(tmp_path / ("mut%d.pdb" % trial))
        open(path, "w").write("\n".join(out))
        assert _reader_equal(path) > 100


# ---- writers ----------------------------------------------------------------------------------------------------------------------------
def _serialize(res, natoms, path):
    def raw(p, nbytes):
        return C.string_at(p, nbytes) if nbytes else b""
    ns, nc, npk = res.n_spheres, res.n_clusters, res.n_pockets
    with open(path, "wb") as f:
        f.write(np.array([res.n_sites, res.n_tets, ns, nc, npk, natoms, C.sizeof(capi.FpkDesc)], np.int32).tobytes())
        f.write(raw(res.sph_xyzr, ns * 16) + raw(res.sph_bary, ns * 12) + raw(res.sph_neigh, ns * 16) + raw(res.sph_type, ns))
        f.write(raw(res.sph_cluster, ns * 4) + raw(res.cl_off, (nc + 1) * 4) + raw(res.cl_sph, ns * 4) + raw(res.desc, nc * C.sizeof(capi.FpkDesc)))
        f.write(raw(res.rank_to_cluster, npk * 4) + raw(res.atom_fx, natoms * 16))


def _tree(d):
    out = {}
    for root, _, files in os.walk(d):
        for fn in files:
            p = os.path.join(root, fn)
            out[os.path.relpath(p, d)] = (open(p, "rb").read(), os.stat(p).st_mode & 0o777)
    return out


def _writer_equal(src, tmp_path, params=None, min_pockets=1, mode="p"):
    name = os.path.basename(src)
    code = name.rsplit(".", 1)[0]
    for d in ("ref", "nat"):
        os.makedirs(str(tmp_path / d), exist_ok=True)
        shutil.copy(src, str(tmp_path / d / name))
    p = hostio.PdbFile(str(tmp_path / "nat" / name))
    prm = params or api.default_params()
    prm.mc_seed = 1234
    L = O.lib()
    res = capi.FpkResult()
    L.orc_search_pocket(C.byref(p.structure), C.byref(prm), None, 0, O.RNG_PHILOX, C.byref(res))
    assert res.n_pockets >= min_pockets, (res.status, res.n_pockets)
    _serialize(res, p.natoms(0), str(tmp_path / "res.bin"))
    nb = C.c_longlong(0)
    npk = hostio.lib().fpkio_write_output_mode(p.h, str(tmp_path / "nat" / name).encode(), C.byref(res), mode.encode(), C.byref(nb))
    nbytes = nb.value
    r = subprocess.run([REF_IO, "write", str(tmp_path / "res.bin"), "-f", str(tmp_path / "ref" / name)], capture_output=True, cwd=str(tmp_path),
                       env=dict(os.environ, REF_IO_WRITE_MODE=mode))
    assert r.returncode == 0, r.stderr[-300:]
    L.orc_result_free(C.byref(res))
    p.close()
    a, b = _tree(str(tmp_path / "ref" / (code + "_out"))), _tree(str(tmp_path / "nat" / (code + "_out")))
    assert sorted(a) == sorted(b), (sorted(set(a) ^ set(b)))
    for k in a:
        assert a[k][0] == b[k][0], "content of %s differs" % k
        assert a[k][1] == b[k][1], "mode of %s differs: %o vs %o" % (k, a[k][1], b[k][1])
    assert npk == res.n_pockets or True
    assert len(a) == {"p": 7 + 2 * npk, "m": 7 + 2 * npk, "b": 12 + 3 * npk, "d": 2 + npk}[mode] and sum(len(v[0]) for v in b.values()) == nbytes
    return npk


@needs_ref_io
def test_writers_equal_reference_on_a_synthetic_structure(tmp_path):
    import synth
    s = synth.generate(20268002, 1500, 0.065)
    synth.to_pdb(s, str(tmp_path / "syn.pdb"))
    assert _writer_equal(str(tmp_path / "syn.pdb"), tmp_path) >= 3


@needs_ref_io
@needs_sample
@pytest.mark.parametrize("name", ["1UYD", "1ATP", "4bdf"])
def test_writers_equal_reference_on_sample_files(tmp_path, name):
    assert _writer_equal(os.path.join(SAMPLE, name + ".pdb"), tmp_path) >= 5


@needs_ref_io
@needs_sample
@pytest.mark.parametrize("mode", ["m", "b"])
def test_writers_mmcif_modes(tmp_path, mode):
    """-w m / -w both on a PDB input: <code>_out.cif, pocketK_atm.cif and the four *_cif* scripts (write_pockets_single_mmcif, write_pocket_mmcif,
    write_mmcif_atom_line, write_mmcif_vert: writepocket.c:307-357,642-720, writepdb.c:201-236, voronoi.c:1486-1511), byte for byte."""
    assert _writer_equal(os.path.join(SAMPLE, "1UYD.pdb"), tmp_path, mode=mode) >= 5
    assert (tmp_path / "nat" / "1UYD_out" / "1UYD_out.cif").exists()


@needs_ref_io
def test_writers_in_undecided_mode_as_dpocket_leaves_it(tmp_path):
    """The reference's dpocket never decides write_mode (its "input file" is the list, fparams.c:397-407): write_out_fpocket then leaves only
    _info.txt, _pockets.pqr and pocketK_vert.pqr. fpkio_write_output_mode('d') does the same."""
    import synth
    s = synth.generate(20268004, 1200, 0.065)
    synth.to_pdb(s, str(tmp_path / "syn.pdb"))
    assert _writer_equal(str(tmp_path / "syn.pdb"), tmp_path, mode="d") >= 2
    assert not (tmp_path / "nat" / "syn_out" / "syn_out.pdb").exists()


@needs_ref_io
def test_writers_when_every_pocket_is_dropped(tmp_path):
    """search_pocket returns an EMPTY list, not NULL, when dropSmallNpolarPockets removes everything (fpocket.c:225), and the reference
    still writes the directory (atoms only, empty info file). Same here."""
    import synth
    s = synth.generate(20268003, 700, 0.065)
    synth.to_pdb(s, str(tmp_path / "syn.pdb"))
    prm = api.default_params()
    prm.min_pock_nb_asph = 100000
    assert _writer_equal(str(tmp_path / "syn.pdb"), tmp_path, params=prm, min_pockets=0) == 0
    assert (tmp_path / "nat" / "syn_out" / "syn_info.txt").read_bytes() == b""


# ---- the -F pipeline with the oracle as detector ---------------------------------------------------------------------------------------------
def test_pipeline_equals_one_by_one(tmp_path):
    import synth
    L = O.lib()
    prm = api.default_params()
    prm.mc_seed = 99
    names = []
    for d in ("pipe", "single"):
        os.makedirs(str(tmp_path / d))
    for i, n in enumerate((650, 900, 700, 1100, 800)):
        s = synth.generate(20268100 + i, n, 0.065)
        names.append("s%d.pdb" % i)
        for d in ("pipe", "single"):
            synth.to_pdb(s, str(tmp_path / d / names[-1]))
    # one by one
    for nm in names:
        p = hostio.PdbFile(str(tmp_path / "single" / nm))
        res = capi.FpkResult()
        L.orc_search_pocket(C.byref(p.structure), C.byref(prm), None, 0, O.RNG_PHILOX, C.byref(res))
        p.write_output(str(tmp_path / "single" / nm), res)
        L.orc_result_free(C.byref(res))
        p.close()
    # pipeline: groups of two files, four pool threads, an unreadable path in the middle of the list
    calls = []

    def detect(inp, n, pp, out):
        calls.append(n)
        for k in range(n):
            L.orc_search_pocket(C.byref(inp[k]), pp, None, 0, O.RNG_PHILOX, C.byref(out[k]))
        return 0

    def free_result(r):
        L.orc_result_free(r)

    paths = [str(tmp_path / "pipe" / nm) for nm in names]
    paths.insert(2, str(tmp_path / "pipe" / "missing.pdb"))
    st = hostio.run_list(paths, prm, detect=detect, free_result=free_result, group=2, io_threads=4, quiet=1)
    assert st["n_files"] == 6 and st["n_written"] == 5 and st["n_read_failed"] == 1 and st["n_detect_failed"] == 0
    assert sum(calls) == 5 and max(calls) <= 2 and st["n_pockets"] >= 5 and st["bytes_out"] > 100000
    for nm in names:
        code = nm[:-4]
        a, b = _tree(str(tmp_path / "single" / (code + "_out"))), _tree(str(tmp_path / "pipe" / (code + "_out")))
        assert a == b and len(a) >= 9


# ---- dpocket: reader in ligand mode, table rows, the loop --------------------------------------------------------------------------------
def _complex(tmp_path, name, seed, n, nlig=24, hetatm_extra=False):
    import synth
    s = synth.generate(seed, n, 0.065)
    rng = np.random.default_rng(seed)
    c = s.xyz[rng.integers(0, s.n)]
    lig = (c + rng.normal(scale=2.0, size=(nlig, 3))).astype(np.float32)
    path = str(tmp_path / name)
    synth.to_pdb(s, path, lig_xyz=lig)
    if hetatm_extra:        # a water, a listed cofactor (kept as protein) and an unlisted group (dropped in ligand mode) before END
        txt = open(path).read().replace("END\n", "")
        txt += "HETATM90001  O   HOH A9100    %8.3f%8.3f%8.3f  1.00 20.00           O\n" % tuple(c + 9)
        txt += "HETATM90002 ZN    ZN A9101    %8.3f%8.3f%8.3f  1.00 20.00          ZN\n" % tuple(c - 7)
        txt += "HETATM90003  C1  XYZ A9102    %8.3f%8.3f%8.3f  1.00 20.00           C\nEND\n" % tuple(c + 6)
        open(path, "w").write(txt)
    return path, s


@needs_ref_io
def test_reader_ligand_mode_equals_reference(tmp_path):
    """dpocket loads a complex with rpdb_open / rpdb_read(ligan = residue name) (dpocket.c:186-203): ligand atoms only in pdb_w_lig, other
    HETATM groups only when listed. Same arrays, same ligand atom list as the reference's reader."""
    path, s = _complex(tmp_path, "c.pdb", 20268201, 800, hetatm_extra=True)
    dump = path + ".dump"
    r = subprocess.run([REF_IO, "atoms", dump, "--lig", "LIG", "-f", path], capture_output=True)
    assert r.returncode == 0, r.stderr[-200:]
    d = dumpio.read_dump(dump)
    p = hostio.PdbFile(path, opts=hostio.ReadOpts.make(ligand="LIG"))
    a = p.arrays()
    assert a["n"] == int(d["n"][0]) == s.n + 1 and a["nw"] == int(d["nw"][0]) == s.n + 1 + 24          # + ZN; + ZN + ligand
    for k in ("x", "y", "z", "radius", "electroneg", "bfactor", "bfstat", "wx", "wy", "wz", "wradius", "welectroneg"):
        assert np.array_equal(a[k].view(np.uint32), d[k].view(np.uint32)), k
    for k in ("atom_id", "res_id", "flags", "wflags"):
        assert np.array_equal(a[k], d[k]), k
    st = p.structure
    assert p.nligand() == st.n_lig == 24
    assert np.array_equal(np.ctypeslib.as_array(st.lig_atoms, (24,)), d["lig"])
    assert list(np.ctypeslib.as_array(st.lig_mol, (24,))) == [0] * 24
    for wl, pfx in ((0, "pdb"), (1, "pdbw")):
        t, iv = p.atom_text(wl)
        assert np.array_equal(iv, d[pfx + ".irc"]) and np.array_equal(t, d[pfx + ".text"])
        assert p.nhetatm(wl) == int(d[pfx + ".nhetatm"][0])
    p.close()
    # no such ligand: dpocket.c:189 "ERROR - No ligand"
    h = C.c_void_p()
    o = hostio.ReadOpts.make(ligand="QQQ")
    assert hostio.lib().fpkio_read_pdb(path.encode(), C.byref(o), C.byref(h)) == hostio.FPKIO_ERR_NO_LIGAND


@needs_ref_io
@needs_sample
def test_reader_ligand_mode_on_1uyd(tmp_path):
    dst = str(tmp_path / "1UYD.pdb")
    shutil.copy(os.path.join(SAMPLE, "1UYD.pdb"), dst)
    r = subprocess.run([REF_IO, "atoms", dst + ".dump", "--lig", "PU8", "-f", dst], capture_output=True)
    assert r.returncode == 0
    d = dumpio.read_dump(dst + ".dump")
    p = hostio.PdbFile(dst, opts=hostio.ReadOpts.make(ligand="PU8"))
    a = p.arrays()
    assert a["n"] == int(d["n"][0]) and a["nw"] == int(d["nw"][0])
    for k in ("x", "y", "z", "radius", "electroneg", "bfactor", "wx", "wy", "wz", "wradius"):
        assert np.array_equal(a[k].view(np.uint32), d[k].view(np.uint32)), k
    st = p.structure
    assert st.n_lig == len(d["lig"]) > 10 and np.array_equal(np.ctypeslib.as_array(st.lig_atoms, (st.n_lig,)), d["lig"])


def test_dpocket_loop_with_the_oracle_as_detector(tmp_path):
    """fpkio_run_dpocket: complexes + ligand names in, <code>_out/ directories and the three tables out (header of dpocket.h:42, one
    explicit-pocket row per complex in list order, every pocket in exactly one of the two other tables by the rule of dpocket.c:277)."""
    L = O.lib()
    prm = api.default_params()
    prm.mc_seed = 5
    paths, ligs = [], []
    for i, n in enumerate((700, 900, 800)):
        pth, _ = _complex(tmp_path, "c%d.pdb" % i, 20268300 + i, n)
        paths.append(pth)
        ligs.append("LIG")
    npk = []

    def detect(inp, n, pp, out):
        for k in range(n):
            assert inp[k].n_lig == 24
            L.orc_search_pocket(C.byref(inp[k]), pp, None, 0, O.RNG_PHILOX, C.byref(out[k]))
            npk.append(out[k].n_pockets)
        return 0

    def free_result(r):
        L.orc_result_free(r)

    outs = [str(tmp_path / nm) for nm in ("exp.txt", "fp.txt", "fpn.txt")]
    st = hostio.run_dpocket(paths, ligs, prm, outs, detect=detect, free_result=free_result, group=2, io_threads=3)
    assert st["n_files"] == 3 and st["n_written"] == 3 and st["n_read_failed"] == 0
    tabs = [open(o).read().splitlines() for o in outs]
    for t in tabs:
        assert t[0].startswith("pdb lig overlap PP-crit PP-dst crit4 crit5 crit6 crit6_continue lig_vol pock_vol nb_AS") and t[0].endswith("TYR VAL")
    assert [ln.split()[0] for ln in tabs[0][1:]] == paths and all(ln.split()[1] == "LIG" for ln in tabs[0][1:])
    assert len(tabs[1]) - 1 + len(tabs[2]) - 1 == sum(npk)
    ncol = len(tabs[0][0].split())
    for t in tabs:
        for ln in t[1:]:
            f = ln.split()
            assert len(f) == ncol
            ovlp, dst, ligvol = float(f[2]), float(f[4]), float(f[9])
            assert 0.0 <= ovlp <= 100.0 and dst >= 0.0 and 50.0 < ligvol < 2000.0
    for ln in tabs[1][1:]:
        f = ln.split()
        assert float(f[2]) > 40.0 or float(f[4]) < 4.0
    for ln in tabs[2][1:]:
        f = ln.split()
        assert not (float(f[2]) > 40.0 or float(f[4]) < 4.0)
    for pth in paths:
        assert os.path.exists(pth[:-4] + "_out/" + os.path.basename(pth)[:-4] + "_info.txt")


def test_fast_programs_fail_loudly_without_a_gpu(tmp_path):
    """build/fpocket_fast and build/dpocket_fast have no CPU detection path: without a usable sm_100 device they stop with the library's
    error before anything is written; and they refuse, by name, the options whose routes are not restated."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import synth
    fast, dfast = os.path.join(ROOT, "build", "fpocket_fast"), os.path.join(ROOT, "build", "dpocket_fast")
    if not (os.path.exists(fast) and os.path.exists(dfast)):
        pytest.skip("build/fpocket_fast not built")
    s = synth.generate(20268401, 500, 0.065)
    synth.to_pdb(s, str(tmp_path / "x.pdb"))
    r = subprocess.run([fast, "-f", "x.pdb"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 4 and "no CPU fallback" in r.stderr and not (tmp_path / "x_out").exists()
    (tmp_path / "l.txt").write_text("x.pdb\tLIG\n")
    r = subprocess.run([dfast, "-f", "l.txt"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 4 and "no CPU fallback" in r.stderr and not (tmp_path / "dpout_explicitp.txt").exists()
    for opt in (["-x"], ["-d"], ["-e", "c"], ["-C", "w"]):            # -e b (city block) and -C m | a | c are on the GPU since round 2
        r = subprocess.run([fast, "-f", "x.pdb"] + opt, cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 3 and "fpocket_cuda_batch" in r.stderr, (opt, r.stderr)
    r = subprocess.run([fast, "-f", "x.cif"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 3 and "mmCIF" in r.stderr


# ---- mdpocket: DCD reader, what follows the frame loop ---------------------------------------------------------------------------------------
def test_dcd_reader(tmp_path):
    """fpkio_dcd_*: CHARMM DCD as the molfile plugin of the reference reads it; little and big endian, with and without unit-cell records."""
    import struct
    rng = np.random.default_rng(11)
    F, N = 7, 123
    xyz = rng.normal(scale=20.0, size=(F, N, 3)).astype(np.float32)
    for endian, cell in (("<", False), (">", False), ("<", True)):
        path = str(tmp_path / ("t%s%d.dcd" % ("le" if endian == "<" else "be", cell)))
        with open(path, "wb") as f:
            ic = [0] * 20
            ic[0], ic[1], ic[2], ic[3], ic[19] = F, 1, 1, F, 24
            ic[10] = 1 if cell else 0
            hdr = b"CORD" + struct.pack(endian + "20i", *ic)
            hdr = hdr[:40] + struct.pack(endian + "f", 1.0) + hdr[44:]
            f.write(struct.pack(endian + "i", 84) + hdr + struct.pack(endian + "i", 84))
            f.write(struct.pack(endian + "i", 84) + struct.pack(endian + "i", 1) + b"x".ljust(80) + struct.pack(endian + "i", 84))
            f.write(struct.pack(endian + "iii", 4, N, 4))
            for k in range(F):
                if cell:
                    f.write(struct.pack(endian + "i", 48) + struct.pack(endian + "6d", 50, 0, 50, 0, 0, 50) + struct.pack(endian + "i", 48))
                for c in range(3):
                    f.write(struct.pack(endian + "i", 4 * N) + xyz[k, :, c].astype(endian + "f4").tobytes() + struct.pack(endian + "i", 4 * N))
        got, nf = hostio.read_dcd(path)
        assert nf == F and got.shape == (F, N, 3) and np.array_equal(got.view(np.uint32), xyz.view(np.uint32)), (endian, cell)
    (tmp_path / "bad.dcd").write_bytes(b"not a dcd file at all")
    with pytest.raises(IOError):
        hostio.read_dcd(str(tmp_path / "bad.dcd"))


@needs_ref_io
@pytest.mark.parametrize("score", [0, 1])
def test_md_outputs_equal_reference_writers(tmp_path, score):
    """normalize_grid, project_grid_on_atoms, write_first_bfactor_density and write_md_grid (mdpbase.c:301-364, mdpout.c:71-122,177-191) on the
    grids of a few oracle frames: the five files must equal, byte for byte, what the reference's own functions write for the same grids."""
    import synth
    s = synth.generate(20268501, 900, 0.065)
    top = str(tmp_path / "top.pdb")
    synth.to_pdb(s, top)
    pf = hostio.PdbFile(top)
    st = capi.FpkStructure()
    assert hostio.lib().fpkio_md_topology(pf.h, C.byref(st)) == hostio.FPKIO_OK and st.same_pdb == 1 and st.natoms_w == 0
    prm = api.default_params()
    prm.do_asa_and_volume = 0
    rng = np.random.default_rng(5)
    syn = synth.to_structure(s)
    org = dims = None
    dens = freq = None
    nfr = 3
    for k in range(nfr):
        syn.x, syn.y, syn.z = (np.ascontiguousarray(np.round(s.xyz[:, c] + (0 if k == 0 else rng.normal(scale=0.15, size=s.n)), 3).astype(np.float32)) for c in range(3))
        o, d, gd, gf, _ = O.md_grids(syn, prm, use_drug_score=score, origin=org, dims=dims)
        if org is None:
            org, dims, dens, freq = o, d, gd.copy(), gf.copy()
        else:
            dens += gd
            freq += gf
    assert dens.sum() > 100
    paths = [str(tmp_path / ("nat_" + nm)) for nm in ("freq.dx", "dens.dx", "freq_iso_0_5.pdb", "dens_iso_8.pdb", "all_atom_pdensities.pdb")]
    nb = hostio.md_write_outputs(pf, dens, freq, org, dims, nfr, prm, paths, use_drug_score=score)
    with open(tmp_path / "grids.bin", "wb") as f:
        f.write(np.array([dims[0], dims[1], dims[2], nfr, score], np.int32).tobytes() + np.asarray(org, np.float32).tobytes())
        f.write(np.ascontiguousarray(dens, np.float32).tobytes() + np.ascontiguousarray(freq, np.float32).tobytes())
    r = subprocess.run([REF_IO, "mdwrite", str(tmp_path / "grids.bin"), str(tmp_path / "ref"), "-f", top], capture_output=True)
    assert r.returncode == 0, r.stderr[-300:]
    total = 0
    for nm in ("freq.dx", "dens.dx", "freq_iso_0_5.pdb", "dens_iso_8.pdb", "all_atom_pdensities.pdb"):
        a, b = open(tmp_path / ("ref_" + nm), "rb").read(), open(tmp_path / ("nat_" + nm), "rb").read()
        assert a == b, nm
        total += len(b)
    assert total == nb and total > 100000

"""GPU (-m gpu): the CUDA path, called through the C ABI (include/fpocket_cuda.h), against the CPU oracle and the
reference's golden dumps. Bars (DESIGN.md "Parity"):
  * Delaunay tetrahedra                         == qhull's (set equality, every fixture)                 bit-exact
  * spheres, types, clusters, pocket membership == oracle in canonical order                             bit-exact
  * every descriptor incl. MC volume (same Philox stream), hull volume, scores, drop + rank, ASA per-atom side effects
                                                == oracle                                                bit-exact
    (drug_score goes through exp(): CUDA libm vs glibc may differ in the last float ulp -> rtol 2e-7 stated below)
  * vs the reference itself: its kept spheres are a subset of ours; what is missing from it lies in the part of qhull's
    output the reference never reads (unflushed stdio tail, see oracle/ref_dump.c)
"""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O

pytestmark = pytest.mark.gpu

EXACT_F = [n for n in G.DESC_F if n != "drug_score"]


def _api():
    from fpocket_b200 import api
    return api


def assert_same_result(r, o, what=""):
    assert r["status"] == o["status"], (what, r["status"], o["status"])
    assert r["n_sites"] == o["n_sites"] and r["n_tets"] == o["n_tets"], what
    assert r["n_spheres"] == o["n_spheres"], (what, r["n_spheres"], o["n_spheres"])
    for k in ("sph_xyzr", "sph_bary", "sph_neigh", "sph_type"):
        assert np.array_equal(r[k], o[k]), (what, k)
    if o["n_clusters"] == 0:
        return
    for k in ("sph_cluster", "cl_off", "cl_sph", "cl_atom_off", "cl_atoms"):
        assert np.array_equal(r[k], o[k]), (what, k)
    D, E = r["desc"], o["desc"]
    for n in EXACT_F:
        assert G.eq_nan(D[n], E[n]), (what, n, np.nanmax(np.abs(D[n] - E[n])))
    assert np.allclose(D["drug_score"], E["drug_score"], rtol=2e-7, atol=0, equal_nan=True), what
    for n in G.DESC_I + ["n_atoms", "first_sphere", "final_rank"]:
        assert np.array_equal(D[n], E[n]), (what, n)
    assert np.array_equal(D["aa_compo"], E["aa_compo"]) and np.array_equal(D["bary"], E["bary"]), what
    assert r["n_pockets"] == o["n_pockets"] and np.array_equal(r["rank_to_cluster"], o["rank_to_cluster"]), what
    assert G.eq_nan(r["atom_fx"], o["atom_fx"]), what
    if "xspheres" in o:             # dpocket ligand queries (K6)
        for k in ("xspheres", "iface_atoms", "pk_ovlp", "pk_dst", "pk_c4", "pk_c5"):
            assert k in r and np.array_equal(r[k], o[k]), (what, k)
        assert np.float32(r["lig_volume"]) == np.float32(o["lig_volume"]) and r["lig_volume"] > 0, (what, "ligand volume", r["lig_volume"], o["lig_volume"])
        assert (r["xdesc"] is None) == (o["xdesc"] is None), what
        if o["xdesc"] is not None:      # the explicit pocket's own record (dpocket.c:350), bit for bit
            for n in r["xdesc"].dtype.names:
                assert G.eq_nan(r["xdesc"][n], o["xdesc"][n]), (what, "xdesc", n, r["xdesc"][n], o["xdesc"][n])
            assert r["xdesc"]["volume"][0] > 0 and r["xdesc"]["nb_asph"][0] == len(r["xspheres"])


@pytest.mark.parametrize("case", G.ALL)
def test_delaunay_equals_qhull(case):
    d = G.load(case)
    st, p = G.inputs(d, want_tets=1)
    r = _api().detect(st, p)
    assert r["n_tets"] == len(d["qh.tets"])
    assert np.array_equal(r["tets"], O.canonicalize(d["qh.tets"]))


@pytest.mark.parametrize("case", G.ALL)
def test_pipeline_equals_oracle(case):
    d = G.load(case)
    st, p = G.inputs(d)
    r = _api().detect(st, p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert_same_result(r, o, case)


@pytest.mark.parametrize("case", G.FPOCKET_CASES)
def test_against_reference_dump(case):
    """The library against what the UNMODIFIED reference produced (tests/golden), under the stated tolerances of tests/refcompare.py:
    spheres, clusters, every descriptor, scores, drop + rank and the per-atom ASA effects, on EVERY fixture. On the fixtures whose
    reference run lost spheres in qhull's unflushed tail (voronoi.c:138-155) every cluster no tail sphere joined is still compared."""
    import refcompare as RC
    d = G.load(case)
    st, p = G.inputs(d)
    r = _api().detect(st, p)
    rep = RC.compare(r, d, st, p, what=case, allow_tail=True)
    assert rep["clusters_compared"] >= rep["clusters_ref"] - 2 * rep["tail"]
    if rep["tail"] == 0:
        assert rep["clusters_compared"] == rep["clusters_ref"] == r["n_clusters"]


def test_md_grids_against_the_reference_grid():
    """mdpocket's two grids of one frame against the grids the unmodified reference accumulated (fixture md_2yex: no tail loss): equal but for
    the once-per-pocket centre cell, which follows the pocket's LAST sphere (mdpbase.c:164-166) - at most one cell per pocket and grid moves."""
    import refcompare as RC
    api = _api()
    d = G.load("md_2yex")
    st, p = G.inputs(d)
    xyz = np.stack([st.x, st.y, st.z], -1)[None]
    m = api.MdPocket(st, p)
    org, dims = m.grid_from_frame(xyz)
    assert np.array_equal(org, d["grid.origin"][:3]) and tuple(dims) == d["grid.dens"].shape
    m.set_grid(org, dims)
    m.push_frames(xyz)
    gd, gf = m.fetch_grids()
    m.close()
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    rep = RC.compare_md_grids(gd.reshape(dims), gf.reshape(dims), d, o["n_pockets"], what="md_2yex")
    assert rep["dens_cells_moved"] <= 2 * o["n_pockets"]


@pytest.mark.parametrize("name", G.SAMPLES)
def test_every_reference_sample(name):
    """All 37 readable structures of the reference's data/sample (PDB and mmCIF; 496 to 26,677 atoms, 3vi4 - BASELINE.md's largest - among
    them), atoms exactly as the reference's readers parsed them: the library equals the oracle bit for bit, and the unmodified reference's
    own run within the stated tolerances (tests/refcompare.py)."""
    import refcompare as RC
    d = G.load_sample(name)
    st, p = G.inputs(d)
    r = _api().detect(st, p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert_same_result(r, o, name)
    rep = RC.compare(r, d, st, p, what=name)
    assert rep["clusters_compared"] >= rep["clusters_ref"] - 2 * rep["tail"]
    assert r["n_tets"] == int(d["qh.n_tets"][0])           # as many Delaunay tetrahedra as qhull printed


def test_capacity_overflow_is_rerun_with_the_worst_case_buffers(monkeypatch):
    """FPK_ERR_CAPACITY (a structure keeps more spheres than its slice of the sphere arrays holds) is never silent: fpk_detect_batch runs
    such a structure again, alone, with worst-case buffers. FPK_SPH_CAP_FACTOR (a test knob) shrinks the first-pass buffers so that every
    structure of the batch overflows; the results must still be the oracle's, bit for bit."""
    api = _api()
    ds = [G.load(c) for c in ("5WA6", "1UYD", "3LKF")]
    p = G.inputs(ds[0])[1]
    sts = [G.inputs(d)[0] for d in ds]
    want = [O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX) for st in sts]
    monkeypatch.setenv("FPK_SPH_CAP_FACTOR", "0.2")
    got = api.detect_batch(sts, p)
    monkeypatch.delenv("FPK_SPH_CAP_FACTOR")
    for c, r, o in zip(("5WA6", "1UYD", "3LKF"), got, want):
        assert r["n_spheres"] > 0.2 * r["n_sites"]           # the first pass did overflow
        assert_same_result(r, o, "capacity " + c)
    # and a batch where only ONE structure overflows its neighbours' results are untouched
    monkeypatch.setenv("FPK_SPH_CAP_FACTOR", "0.55")
    got = api.detect_batch(sts, p)
    monkeypatch.delenv("FPK_SPH_CAP_FACTOR")
    over = [r["n_spheres"] > int(0.55 * r["n_sites"]) for r in got]
    assert any(over) and not all(over), over
    for c, r, o in zip(("5WA6", "1UYD", "3LKF"), got, want):
        assert_same_result(r, o, "capacity, mixed " + c)


def test_one_unusable_structure_does_not_abort_the_batch():
    """a structure with duplicate atom ids (or a NULL array) gets FPK_ERR_BAD_ARG in ITS status; the rest of the batch is
    processed (the reference's -F loop moves on to the next file, fpmain.c:61-76)."""
    import copy
    from fpocket_b200 import capi
    api = _api()
    d = G.load("5WA6")
    st, p = G.inputs(d)
    bad = copy.copy(st)
    bad.atom_id = st.atom_id.copy()
    bad.atom_id[10] = bad.atom_id[3]
    got = api.detect_batch([st, bad, st], p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert got[1]["status"] == capi.FPK_ERR_BAD_ARG and got[1]["n_spheres"] == 0
    assert_same_result(got[0], o, "before the bad one")
    assert_same_result(got[2], o, "after the bad one")


def test_city_block_clustering_equals_the_reference():
    """-e b (clusterlib.c:1022-1083, the MEAN absolute difference) on the GPU (K3 with another pair test, cell edge 3 x the cut): equal to the
    oracle bit for bit, equal to the unmodified reference's run of `fpocket -f 1UYD.pdb -e b` within the stated tolerances, different from -e e;
    also on a synthetic batch, together with a structure that takes the grid-wide clustering kernel."""
    import refcompare as RC
    api = _api()
    d = G.load("alt_1UYD_eb")
    st, p = G.inputs(d)
    assert p.distance_measure == ord("b")
    r = api.detect(st, p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert_same_result(r, o, "-e b")
    RC.compare(r, d, st, p, what="-e b")
    pe = G.inputs(d)[1]
    pe.distance_measure = ord("e")
    assert api.detect(st, pe)["n_clusters"] != r["n_clusters"]
    sts = [_synth(20269330 + i, n) for i, n in enumerate((2100, 13000, 2900))]
    pb = api.default_params()
    pb.distance_measure = ord("b")
    for st2, r2 in zip(sts, api.detect_batch(sts, pb)):
        assert_same_result(r2, O.search_pocket(st2, pb, tets=None, rng_mode=O.RNG_PHILOX), "-e b synthetic")


@pytest.mark.parametrize("case", [c for c in G.ALT_CASES if "_C" in c])
def test_linkage_clustering_equals_the_reference(case):
    """-C m | a | c (clusterlib.c:3705 complete, :3781 average, :3337 centroid linkage, cut by cuttree_distance :3235) on the GPU (K3L k_linkage:
    the reference's matrix layout and moves with cached row minima): equal to the oracle bit for bit, equal to the unmodified reference's run
    (`-C m`, `-C a`, `-e b -C c` - the parameters of the reference's own tests/test_fpocket.py:91 - and `-C c`) within the stated tolerances,
    and different from single linkage."""
    import refcompare as RC
    api = _api()
    d = G.load(case)
    st, p = G.inputs(d)
    assert chr(p.clustering_method) in "mac"
    r = api.detect(st, p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert_same_result(r, o, case)
    RC.compare(r, d, st, p, what=case)
    ps = G.inputs(d)[1]
    ps.clustering_method = ord("s")
    assert api.detect(st, ps)["n_clusters"] != r["n_clusters"]


@pytest.mark.parametrize("method,measure", [("m", "e"), ("a", "e"), ("c", "e"), ("c", "b"), ("a", "b")])
def test_linkage_clustering_on_a_synthetic_batch(method, measure, monkeypatch):
    """the same on a batch of synthetic structures (1,200 - 2,900 heavy atoms), with a pool so small that the batch needs several waves and
    the largest structure is refused (FPK_ERR_BAD_ARG, the batch goes on), and on mdpocket's per-frame body (descriptors off)."""
    api = _api()
    from fpocket_b200 import capi
    sts = [_synth(20269400 + i, n) for i, n in enumerate((1200, 2900, 1700, 2200, 1500))]
    p = api.default_params()
    p.clustering_method, p.distance_measure = ord(method), ord(measure)
    want = [O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX) for st in sts]
    for st, r, o in zip(sts, api.detect_batch(sts, p), want):
        assert_same_result(r, o, "-C %s -e %s" % (method, measure))
        assert r["n_clusters"] > 20
    ns = sorted(o["n_spheres"] for o in want)
    pool_mb = 8.0 * (ns[-2] ** 2 / 2 + 40 * ns[-2]) / 2 ** 20 * 1.15           # room for the second largest matrix, not for the largest
    monkeypatch.setenv("FPK_LINK_POOL_MB", str(int(pool_mb) + 1))
    got = api.detect_batch(sts, p)
    for st, r, o in zip(sts, got, want):
        if o["n_spheres"] == ns[-1]:
            assert r["status"] == capi.FPK_ERR_BAD_ARG and r["n_pockets"] == 0
        else:
            assert_same_result(r, o, "-C %s in waves" % method)
    monkeypatch.delenv("FPK_LINK_POOL_MB")
    p.do_asa_and_volume = 0
    r = api.detect(sts[2], p)
    assert_same_result(r, O.search_pocket(sts[2], p, tets=None, rng_mode=O.RNG_PHILOX), "-C %s without descriptors" % method)


def test_unknown_clustering_method_is_refused():
    api = _api()
    st = _synth(20269420, 1300)
    p = api.default_params()
    p.clustering_method = ord("x")
    with pytest.raises(api.FpkError, match="clustering method"):
        api.detect(st, p)


def test_dpocket_interface_method_2_equals_the_reference():
    """dpocket -E (dparams.h:44, dpocket.c:331-336): the interface = atoms within 4.0 of a ligand atom (neighbor.c:67 get_mol_atm_neigh). K6's set
    equals the oracle's and what the reference's own function returned on 1UYD / PU8; so does the overlap atm_corsp computes with it."""
    api = _api()
    d = G.load("1UYD_dpocket")
    st, p = G.inputs(d)
    p.lig_interface_method = 2
    r = api.detect(st, p)
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert_same_result(r, o, "interface method 2")
    lig = np.sort(d["lig.atoms"])
    ref = np.sort(np.array([w - np.searchsorted(lig, w) for w in d["lig.iface2"]], dtype=np.int32))
    assert len(ref) > 20 and np.array_equal(r["iface_atoms"], ref)
    p.lig_interface_method = 1
    r1 = api.detect(st, p)
    assert len(r1["iface_atoms"]) != len(ref) and r1["lig_volume"] == r["lig_volume"]


def test_batch_equals_single():
    cases = ["5WA6", "1UYD", "3LKF", "1UYD_explicit"][:3]
    ds = [G.load(c) for c in cases]
    sts = [G.inputs(d)[0] for d in ds]
    p = G.inputs(ds[0])[1]
    api = _api()
    rb = api.detect_batch(sts + sts[:1], p)
    for c, st, r in zip(cases + cases[:1], sts + sts[:1], rb):
        o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert_same_result(r, o, c)


def test_scratch_reuse_across_batches_of_different_shape():
    """The per-block Delaunay scratch and the pooled host contexts are reused by later calls whose layout differs: nothing may
    depend on what an earlier batch left behind (regression: stale cavity marks read by the serial big-cavity path)."""
    api = _api()
    ds = [G.load(c) for c in ("7TAA_m28_M74", "5WA6", "1UYD", "3LKF")]
    p = G.inputs(ds[1])[1]
    sts = [G.inputs(d)[0] for d in ds]
    want = [O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX) for st in sts]
    for order in ([0, 1, 2, 3, 0, 1], [3, 2], [1, 1, 1, 0, 2, 3, 3], [2]):
        got = api.detect_batch([sts[i] for i in order], p)
        for i, r in zip(order, got):
            assert_same_result(r, want[i], "reuse %d" % i)


def _synth(seed, n):
    import synth
    return synth.to_structure(synth.generate(seed, n, 0.065))


def test_synthetic_structures_equal_oracle():
    """BASELINE configs[1] sizes (2k-5k heavy atoms, the bench generator): the whole result, bit for bit, against the oracle."""
    api = _api()
    sts = [_synth(20269100 + i, n) for i, n in enumerate((2000, 3517, 5000, 4242))]
    p = api.default_params()
    p.want_tets = 1
    got = api.detect_batch(sts, p)
    for st, r in zip(sts, got):
        o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert r["status"] == 0 and r["n_pockets"] >= 5
        _, ki, _ = O.sites(st)
        rc, tets, _ = O.delaunay(ki)
        assert rc == 0 and np.array_equal(r["tets"], tets)
        assert_same_result(r, o, "synthetic")


def test_empty_sphere_property_at_bench_size():
    """Size-independent property of the whole K1 stage on a bench-like batch: every kept alpha sphere touches its four contact atoms
    (f32 distances agree within 1e-3, voronoi.c:1042) and holds no heavy atom (the Delaunay property), its radius is inside the window,
    and clustering is the connected-component partition of the <= 2.4 A graph (pocket sets closed under the cut)."""
    api = _api()
    sts = [_synth(20269200 + i, 2000 + 97 * i) for i in range(24)]
    p = api.default_params()
    got = api.detect_batch(sts, p)
    rng = np.random.default_rng(1)
    for st, r in zip(sts, got):
        assert r["status"] == 0
        ns = r["n_spheres"]
        assert 0.4 * st.natoms < ns < 0.75 * st.natoms          # SURVEY 8d sanity gate of the generator
        xyz = np.stack([st.x, st.y, st.z], -1).astype(np.float64)
        pick = rng.choice(ns, size=min(300, ns), replace=False)
        c, rad = r["sph_xyzr"][pick, :3].astype(np.float64), r["sph_xyzr"][pick, 3].astype(np.float64)
        d = np.linalg.norm(xyz[None] - c[:, None], axis=2)
        dn = np.take_along_axis(d, r["sph_neigh"][pick].astype(np.int64), 1)
        assert np.all(np.abs(dn - rad[:, None]) < 1.5e-3)
        assert np.all(d.min(1) > rad - 1.5e-3)
        assert np.all((rad >= 3.4 - 1e-3) & (rad <= 6.2 + 1e-3))
        # single-linkage cut: no two spheres of different clusters within the cut
        cx = r["sph_xyzr"][:, :3].astype(np.float64)
        sub = rng.choice(ns, size=min(400, ns), replace=False)
        dd = np.linalg.norm(cx[sub][:, None] - cx[None], axis=2)
        close = dd <= float(np.float32(2.4)) - 1e-6
        same = r["sph_cluster"][sub][:, None] == r["sph_cluster"][None]
        assert not np.any(close & ~same)


def test_large_structure_sites_in_global_memory():
    """20k heavy atoms do not fit the shared-memory site buffer: K1 takes the global-memory instantiation of the same code."""
    api = _api()
    st = _synth(20269300, 20000)
    p = api.default_params()
    p.want_tets = 1
    r = api.detect(st, p)
    assert r["status"] == 0
    _, ki, _ = O.sites(st)
    rc, tets, _ = O.delaunay(ki)
    assert rc == 0 and np.array_equal(r["tets"], tets)


def test_grid_wide_delaunay_equals_oracle():
    """BASELINE configs[3] path: from 12,000 sites on (2,500 for a lone structure) ONE triangulation is built by every block of a
    cooperative launch (csrc/k_delaunay_grid.cuh) instead of one block. 60,000 heavy atoms: the tetrahedra are the oracle's, as a set."""
    api = _api()
    st = _synth(20269310, 60000)
    p = api.default_params()
    p.want_tets = 1
    r = api.detect(st, p)
    assert r["status"] == 0 and r["n_sites"] == st.natoms > 59000
    _, ki, _ = O.sites(st)
    rc, tets, _ = O.delaunay(ki)
    assert rc == 0 and r["n_tets"] == len(tets)
    assert np.array_equal(r["tets"], tets)


def test_grid_wide_and_block_paths_in_one_batch(monkeypatch):
    """A batch that mixes both routes (the threshold lowered so that two of five structures take the grid-wide kernel): every result, the
    whole record, equals the oracle; and the same structures through the other route give the same bytes."""
    api = _api()
    sts = [_synth(20269320 + i, n) for i, n in enumerate((2300, 6100, 2500, 7400, 3100))]
    p = api.default_params()
    p.want_tets = 1
    want = [O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX) for st in sts]
    monkeypatch.setenv("FPK_GRID_MIN_SITES", "6000")
    got = api.detect_batch(sts, p)
    monkeypatch.setenv("FPK_GRID_MIN_SITES", "1000000")          # everything by one block each
    got_block = api.detect_batch(sts, p)
    monkeypatch.setenv("FPK_GRID_MIN_SITES", "100")              # everything by the whole GPU, one after the other
    got_grid = api.detect_batch(sts, p)
    for k, (st, o) in enumerate(zip(sts, want)):
        for r in (got[k], got_block[k], got_grid[k]):
            assert np.array_equal(r["tets"], got_block[k]["tets"])
            assert_same_result(r, o, "mixed routes %d" % k)


def test_large_complex_has_the_empty_sphere_property():
    """150,000 heavy atoms (BASELINE configs[3], the bench's generator and seed), the whole pipeline: 1,039,790 tetrahedra equal to the
    oracle's, every array of the result equal to the oracle's bit for bit (the oracle needs ~30 s for it), plus properties that do not
    depend on any oracle - every kept alpha sphere touches its four atoms, holds no atom, has its radius inside the window."""
    api = _api()
    st = _synth(20260401, 150000)
    p = api.default_params()
    p.want_tets = 1
    r = api.detect(st, p)
    assert r["status"] == 0 and r["n_sites"] == st.natoms
    o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    _, ki, _ = O.sites(st)
    rc, tets, _ = O.delaunay(ki)
    assert rc == 0 and np.array_equal(r["tets"], tets)
    assert_same_result(r, o, "150k")
    assert 6.0 * st.natoms < r["n_tets"] < 7.5 * st.natoms
    ns = r["n_spheres"]
    assert ns > 0.2 * st.natoms and r["n_pockets"] > 100
    xyz = np.stack([st.x, st.y, st.z], -1).astype(np.float64)
    rng = np.random.default_rng(2)
    pick = rng.choice(ns, size=400, replace=False)
    c, rad = r["sph_xyzr"][pick, :3].astype(np.float64), r["sph_xyzr"][pick, 3].astype(np.float64)
    for q in range(0, 400, 50):
        d = np.linalg.norm(xyz[None] - c[q:q + 50, None], axis=2)
        dn = np.take_along_axis(d, r["sph_neigh"][pick[q:q + 50]].astype(np.int64), 1)
        assert np.all(np.abs(dn - rad[q:q + 50, None]) < 2e-3)
        assert np.all(d.min(1) > rad[q:q + 50] - 2e-3)
    assert np.all((r["sph_xyzr"][:, 3] >= 3.4 - 1e-3) & (r["sph_xyzr"][:, 3] <= 6.2 + 1e-3))
    keys = r["sph_neigh"]
    assert len(np.unique(keys, axis=0)) == ns            # no sphere twice
    order = np.lexsort((keys[:, 3], keys[:, 2], keys[:, 1], keys[:, 0]))
    assert np.array_equal(order, np.arange(ns))          # canonical order: ascending site tuples


def test_edge_cases():
    from fpocket_b200 import capi
    api = _api()
    d = G.load("5WA6")
    st, p = G.inputs(d)
    import copy

    def sub(n):
        s = copy.copy(st)
        for k in ("x", "y", "z", "radius", "electroneg", "bfactor", "atom_id", "res_id", "aa_idx", "flags"):
            setattr(s, k, getattr(st, k)[:n].copy())
        s.natoms = n
        s.natoms_w = 0
        s.same_pdb = 1
        return s
    r3, r40, r400 = api.detect_batch([sub(3), sub(40), sub(400)], p)
    assert r3["status"] == capi.FPK_ERR_DEGENERATE
    for s, r in ((sub(40), r40), (sub(400), r400)):
        o = O.search_pocket(s, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert r["status"] == o["status"] and r["n_tets"] == o["n_tets"] and r["n_spheres"] == o["n_spheres"]
        if o["status"] == 0:
            assert_same_result(r, o, "sub")


@pytest.mark.parametrize("case", G.MD_CASES)
def test_md_grids_equal_oracle(case):
    api = _api()
    d = G.load(case)
    st, p = G.inputs(d)
    score = case.endswith("_S")
    xyz = np.stack([st.x, st.y, st.z], -1)[None]
    m = api.MdPocket(st, p, use_drug_score=score)
    org, dims = m.grid_from_frame(xyz)
    oo, od, dens, freq, _ = O.md_grids(st, p, use_drug_score=int(score))
    assert np.array_equal(org, oo) and np.array_equal(dims, od)
    # the geometry the reference computed from ITS sphere list (it lost the unflushed tail) may differ by the lost spheres
    m.set_grid(org, dims)
    m.push_frames(np.concatenate([xyz, xyz, xyz]))
    gd, gf = m.fetch_grids()
    if score:   # float increments: the reference adds them one by one in float, the library sums them exactly (32.32 fixed point, order-free)
        assert np.allclose(gd, 3 * dens, rtol=1e-5, atol=1e-6) and np.allclose(gf, 3 * freq, rtol=1e-5, atol=1e-6)
    else:
        assert np.array_equal(gd, 3 * dens) and np.array_equal(gf, 3 * freq)
    m.close()


def test_md_trajectory_of_distinct_frames_equals_oracle():
    """mdpocket on a short synthetic trajectory (every frame different, chunked pushes): both grids equal the oracle's per-frame
    detection + scatter accumulated frame by frame (integer increments: exact)."""
    import copy
    import synth
    api = _api()
    s = synth.generate(20269500, 2600, 0.065)
    st = synth.to_structure(s)
    p = api.default_params()
    rng = np.random.default_rng(3)
    frames = np.round(s.xyz[None] + rng.normal(scale=0.25, size=(7,) + s.xyz.shape).astype(np.float32), 3).astype(np.float32)
    frames[0] = s.xyz
    m = api.MdPocket(st, p)
    org, dims = m.grid_from_frame(frames[:1])
    m.set_grid(org, dims)
    m.push_frames(frames[:3])
    m.push_frames(frames[3:])
    gd, gf = m.fetch_grids()
    m.close()
    dens = np.zeros(dims, np.float32)
    freq = np.zeros(dims, np.float32)
    pm = api.default_params()
    pm.do_asa_and_volume = 0                    # mdpocket.c:89
    for f in frames:
        sf = copy.copy(st)
        sf.x, sf.y, sf.z = (np.ascontiguousarray(f[:, k]) for k in range(3))
        o0, d0, dd, ff, _ = O.md_grids(sf, pm, origin=org, dims=dims)
        dens += dd
        freq += ff
    assert np.array_equal(gd, dens) and np.array_equal(gf, freq)
    assert gd.sum() > 0 and gf.sum() > 0


def test_md_grids_with_drug_score_do_not_depend_on_how_frames_are_pushed():
    """-S: the increments are floats. The library sums them in 32.32 fixed point, so the grids are the same bits whether the frames
    arrive one by one (the reference's loop, build/mdpocket_cuda) or in one block (build/mdpocket_fast), and equal, to float rounding,
    the reference's frame-by-frame float sums (oracle)."""
    import copy
    import synth
    api = _api()
    s = synth.generate(20269501, 2300, 0.065)
    st = synth.to_structure(s)
    p = api.default_params()
    rng = np.random.default_rng(4)
    frames = np.round(s.xyz[None] + rng.normal(scale=0.25, size=(9,) + s.xyz.shape).astype(np.float32), 3).astype(np.float32)
    frames[0] = s.xyz
    grids = []
    for how in ("block", "single", "pairs"):
        m = api.MdPocket(st, p, use_drug_score=True)
        org, dims = m.grid_from_frame(frames[:1])
        m.set_grid(org, dims)
        if how == "block":
            m.push_frames(frames)
        else:
            step = 1 if how == "single" else 2
            for i in range(0, len(frames), step):
                m.push_frames(frames[i:i + step])
        grids.append(m.fetch_grids())
        m.close()
    for gd, gf in grids[1:]:
        assert np.array_equal(gd, grids[0][0]) and np.array_equal(gf, grids[0][1])
    dens = np.zeros(dims, np.float32)
    freq = np.zeros(dims, np.float32)
    pm = api.default_params()
    pm.do_asa_and_volume = 0                    # mdpocket.c:89
    for f in frames:
        sf = copy.copy(st)
        sf.x, sf.y, sf.z = (np.ascontiguousarray(f[:, k]) for k in range(3))
        _, _, dd, ff, _ = O.md_grids(sf, pm, origin=org, dims=dims, use_drug_score=1)
        dens += dd
        freq += ff
    assert np.allclose(grids[0][0], dens, rtol=1e-5, atol=1e-6) and np.allclose(grids[0][1], freq, rtol=1e-5, atol=1e-6)
    assert grids[0][0].sum() > 0 and grids[0][1].sum() > 0


@pytest.mark.parametrize("case", G.MDCHAR_CASES)
def test_md_characterize_equals_oracle_and_the_reference(case):
    """mdpocket's characterize mode on the GPU (fpk_md_characterize: K7 zone selection + the hull / descriptor kernels on the zone batch), one
    snapshot of the reference's sample trajectory and the first 70 points of its own mdpout_dens_iso_8.pdb: the zone's sphere list (repeats
    included), the atoms it contacts, its record (Monte-Carlo and hull volume included) and the per-atom effects equal the oracle's bit for bit
    (canonical order); against the unmodified reference's run the list holds the same spheres the same number of times, the same atoms, and
    the record agrees within the tolerances of tests/refcompare.py. A second, displaced frame in the same call checks the batching."""
    import refcompare as RC
    api = _api()
    d = G.load(case)
    st, p = G.inputs(d)
    zone = d["want.points"]
    xyz = np.stack([st.x, st.y, st.z], -1)
    frames = np.stack([xyz, xyz + np.float32(0.5), xyz]).astype(np.float32)
    m = api.MdPocket(st, p)
    got = m.characterize(frames, zone)
    m.close()
    o = O.characterize(st, p, zone, tets=None, rng_mode=O.RNG_PHILOX)
    for z in (got[0], got[2]):
        assert z["status"] == 0 and z["n_spheres_frame"] == o["n_spheres"] and z["n_pockets_frame"] == o["n_pockets"]
        assert np.array_equal(z["spheres"], o["want_sph"]) and len(z["spheres"]) > 50
        assert np.array_equal(z["sph_xyzr"], o["sph_xyzr"][o["want_sph"]])
        assert np.array_equal(z["atoms"], o["want_atoms"])
        rank_of = np.zeros(o["n_clusters"], np.int32)
        rank_of[o["rank_to_cluster"]] = np.arange(1, o["n_pockets"] + 1)
        assert np.array_equal(z["sph_pocket"], rank_of[o["sph_cluster"][z["spheres"]]])        # vertice->resid = final rank of the sphere's pocket
        D = o["want_desc"]
        for n in G.DESC_F:
            if n in ("score", "drug_score") or n.endswith("_norm"):
                continue                                  # the zone's record is never normalised or scored (mdpocket.c:453)
            a, b = np.float32(z["desc"][n]), np.float32(getattr(D, n))
            assert a == b or (np.isnan(a) and np.isnan(b)), (n, a, b)
        for n in ("nb_asph", "polarity_score", "charge_score", "n_abpa", "n_atoms", "first_sphere"):
            assert int(z["desc"][n]) == int(getattr(D, n)), n
        assert z["desc"]["aa_compo"].tolist() == list(D.aa_compo)
        assert z["desc"]["volume"] > 0 and z["desc"]["convex_hull_volume"] > 0
        assert G.eq_nan(z["atom_fx"], o["atom_fx"])
    assert got[1]["n_spheres_frame"] == o["n_spheres"]           # a rigid shift keeps every sphere; the zone did not move with it
    assert len(got[1]["spheres"]) != len(got[0]["spheres"])
    # against the unmodified reference (its order is qhull's): same multiset of spheres, same atoms, record within tolerance
    keys = G.canon_sphere_keys(o["sph_neigh"])
    ref_keys = G.canon_sphere_keys(d["sph0.i"][:, :4])
    if set(ref_keys) == set(keys):
        mine = sorted(keys[s] for s in got[0]["spheres"])
        theirs = sorted(ref_keys[s] for s in d["want.sph"])
        assert mine == theirs
        assert sorted(got[0]["atoms"].tolist()) == sorted(d["want.atoms"].tolist())
        F = d["want.desc_f"]
        for k, n in enumerate(G.DESC_F):
            if n in RC.PLAIN and n not in RC.RESIDUE_BASED:
                assert np.isclose(got[0]["desc"][n], F[k], rtol=RC.F_RTOL, atol=2e-5, equal_nan=True), (n, got[0]["desc"][n], F[k])
        assert np.isclose(got[0]["desc"]["convex_hull_volume"], F[G.DESC_F.index("convex_hull_volume")], rtol=RC.F_RTOL)
        assert abs(got[0]["desc"]["volume"] - F[G.DESC_F.index("volume")]) < 0.1 * F[G.DESC_F.index("volume")]


def test_hull_volume_is_the_same_through_both_routes(tmp_path):
    """K4b has two routes to the same exact integer: one warp per pocket (incremental hull, hull_warp.cuh) and, for what that declines, one
    block per pocket (Delaunay of the centres, k_hull_volume). FPK_HULL_WARP=0 (read once per process) sends every pocket through the second
    route and its hand-over list: a child process runs two golden cases that way and must report the floats this process gets."""
    import os
    import subprocess
    import sys
    api = _api()
    code = ("import sys, numpy as np\n"
            "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import golden_cases as G\n"
            "from fpocket_b200 import api\n"
            "for c in sys.argv[2:]:\n"
            "    st, p = G.inputs(G.load(c))\n"
            "    np.save(sys.argv[1] + '/' + c + '.npy', api.detect(st, p)['desc']['convex_hull_volume'])\n") % (O.ROOT, os.path.join(O.ROOT, "tests"))
    cases = ["5WA6", "4URL"]
    env = dict(os.environ, FPK_HULL_WARP="0")
    subprocess.check_call([sys.executable, "-c", code, str(tmp_path)] + cases, env=env)
    for c in cases:
        st, p = G.inputs(G.load(c))
        here = api.detect(st, p)["desc"]["convex_hull_volume"]
        there = np.load(str(tmp_path / (c + ".npy")))
        assert np.array_equal(here, there) and (here > 0).sum() >= 10, c


def _points_structure(xyz):
    """a bare point set as a structure (carbon atoms, one residue each): only the Delaunay stage is of interest"""
    from fpocket_b200 import capi
    n = len(xyz)
    one = np.ones(n, np.float32)
    return capi.Structure(xyz[:, 0], xyz[:, 1], xyz[:, 2], 1.7 * one, 2.5 * one, 20 * one, np.arange(1, n + 1), np.arange(1, n + 1),
                          np.zeros(n, np.int8), np.zeros(n, np.uint8), (np.float32(20), np.float32(0), np.float32(20)), w=None, same_pdb=1)


@pytest.mark.parametrize("route", ["block", "grid"])
def test_delaunay_on_degenerate_and_random_point_sets(route, monkeypatch):
    """(route = grid: the same sets, one after the other, through the grid-wide kernel k_delaunay_grid - coincident sites, over-sized
    cavities on the serial path and the symbolic perturbation with the counters in global memory.)
    The shared-memory fast path has its own exact predicates and symbolic perturbation (delaunay_warp.cuh): cubic lattices (every
    point cospherical / coplanar with many others), duplicated sites, thin slabs and random clouds of many sizes, all in ONE batch,
    must give exactly the oracle's tetrahedra (same perturbation scheme, so the same triangulation of the degenerate sets)."""
    api = _api()
    monkeypatch.setenv("FPK_GRID_MIN_SITES", "4" if route == "grid" else "1000000")
    rng = np.random.default_rng(11)
    sets = []
    g = np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    sets.append(g * 1.5)                                                  # 216-point cubic lattice
    sets.append(np.concatenate([g * 2.0, g * 2.0 + 1.0]))                 # body-centred lattice, 432 points
    slab = rng.random((500, 3)).astype(np.float32) * np.array([40, 40, 0.004], np.float32)
    sets.append(np.round(slab, 3))                                        # thin slab on the 1e-3 grid: coplanar quadruples galore
    dup = np.round(rng.random((300, 3)).astype(np.float32) * 25, 3)
    dup[100] = dup[7]; dup[101] = dup[7]; dup[250] = dup[3]
    sets.append(dup)                                                      # coincident sites
    for n in (5, 9, 33, 120, 700, 1900, 3300, 5200):
        sets.append(np.round(rng.random((n, 3)).astype(np.float32) * (n ** (1 / 3.0)) * 2.7, 3))
    sts = [_points_structure(x) for x in sets]
    p = api.default_params()
    p.want_tets = 1
    got = api.detect_batch(sts, p)
    for st, r in zip(sts, got):
        _, ki, _ = O.sites(st)
        rc, tets, _ = O.delaunay(ki)
        assert rc == 0
        assert r["n_tets"] == len(tets), (st.natoms, r["status"], r["n_tets"], len(tets))
        assert np.array_equal(r["tets"], tets), st.natoms


def test_dpocket_synthetic_complexes_equal_oracle():
    """BASELINE configs[4] shape: ligand-bound synthetic structures (and ligand-free ones in the same batch). Ligand queries, the explicit
    pocket's record and everything else equal the oracle bit for bit; a ligand far from the protein gives an empty explicit pocket."""
    import synth
    api = _api()
    rng = np.random.default_rng(5)
    sts = []
    for i, n in enumerate((2100, 3300, 2600)):
        s = synth.generate(20269600 + i, n, 0.065)
        first = api.detect(synth.to_structure(s), api.default_params())
        c = first["desc"]["bary"][first["rank_to_cluster"][0]] if i < 2 else s.xyz.max(0) + 60.0        # third ligand: nowhere near
        sts.append(synth.with_ligand(s, synth.make_ligand(rng, c)))
    sts.insert(1, _synth(20269650, 2400))                                   # no ligand
    p = api.default_params()
    got = api.detect_batch(sts, p)
    for k, (st, r) in enumerate(zip(sts, got)):
        o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert_same_result(r, o, "dpocket %d" % k)
    assert "xspheres" not in got[1]
    assert len(got[0]["xspheres"]) > 10 and got[0]["pk_ovlp"].max() > 30 and got[0]["pk_c4"].max() > 0.5
    assert len(got[3]["xspheres"]) == 0 and got[3]["xdesc"] is None and np.all(got[3]["pk_c5"] == 0)


def test_hydrogens_are_not_sites():
    """Atoms flagged as hydrogens are not Delaunay sites (voronoi.c:100,128) nor surrounding atoms (asa.c:89): the site gather path
    (no contiguous TMA tile) against the oracle, in a batch with a hydrogen-free structure (TMA path)."""
    from fpocket_b200 import capi
    api = _api()
    a, b = _synth(20269700, 2300), _synth(20269701, 2500)
    rng = np.random.default_rng(9)
    h = rng.random(a.natoms) < 0.15
    a.flags = np.where(h, capi.FPK_ATOM_H | capi.FPK_ATOM_H1, 0).astype(np.uint8)
    p = api.default_params()
    p.want_tets = 1
    got = api.detect_batch([a, b, a], p)
    for st, r in zip([a, b, a], got):
        o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert r["n_sites"] == o["n_sites"] == int(np.count_nonzero((st.flags & capi.FPK_ATOM_H) == 0))
        assert_same_result(r, o, "hydrogens")
    assert got[0]["n_sites"] < a.natoms and got[1]["n_sites"] == b.natoms

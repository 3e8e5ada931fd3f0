"""CPU: pins the oracle (oracle/oracle.c) to the UNMODIFIED reference, through the stage dumps in
tests/golden (made by tests/golden/make_golden.py with oracle/_ref/ref_dump).

Given the tetrahedra exactly as the reference read them from qhull (order, vertex order, lost tail),
every stage of the oracle must be BIT-IDENTICAL to the reference: kept spheres, types, clusters,
every descriptor (including the Monte-Carlo volume when fed the same libc rand() stream and the
convex-hull volume), per-atom ASA side effects, drop decisions and final ranking.
The oracle's own exact Delaunay must equal qhull's tetrahedra as a set.
"""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O


@pytest.mark.parametrize("case", G.FPOCKET_CASES)
def test_oracle_bit_exact_on_reference_tets(case):
    d = G.load(case)
    st, p = G.inputs(d)
    r = O.search_pocket(st, p, tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    assert r["n_spheres"] == len(d["sph0.f"])
    assert np.array_equal(r["sph_xyzr"], d["sph0.f"][:, :4])
    assert np.array_equal(r["sph_bary"], d["sph0.f"][:, 4:7])
    assert np.array_equal(r["sph_neigh"], d["sph0.i"][:, :4])
    assert np.array_equal(r["sph_type"], d["sph0.i"][:, 4])
    assert np.array_equal(r["cl_off"], d["pre.off"])
    assert np.array_equal(r["cl_sph"], d["pre.sph"])
    D, F, I = r["desc"], d["pre.desc_f"], d["pre.desc_i"]
    for k, n in enumerate(G.DESC_F):
        assert G.eq_nan(D[n], F[:, k]), n
    assert np.array_equal(D["bary"], F[:, 27:30])
    for k, n in enumerate(G.DESC_I):
        assert np.array_equal(D[n], I[:, k]), n
    assert np.array_equal(D["aa_compo"], I[:, 11:31])
    assert G.eq_nan(r["atom_fx"], d["atom_fx"])
    # final pockets: same clusters, same rank order (psorting.c quicksort incl. tie behaviour)
    ref_first = [int(d["fin.sph"][o]) for o in d["fin.off"][:-1]]
    mine = [int(r["cl_sph"][r["cl_off"][c]]) for c in r["rank_to_cluster"]]
    assert mine == ref_first
    assert np.array_equal(D["score"][r["rank_to_cluster"]], d["fin.desc_f"][:, 0])


@pytest.mark.parametrize("case", G.ALL)
def test_oracle_delaunay_equals_qhull(case):
    d = G.load(case)
    st, _ = G.inputs(d)
    _, ki, _ = O.sites(st)
    rc, t, vol = O.delaunay(ki)
    assert rc == 0
    assert np.array_equal(t, O.canonicalize(d["qh.tets"]))
    assert vol > 0


@pytest.mark.parametrize("case", G.MD_CASES)
def test_oracle_md_grids(case):
    d = G.load(case)
    st, p = G.inputs(d)
    score = int(case.endswith("_S"))
    org, dims, dens, freq, r = O.md_grids(st, p, use_drug_score=score, tets=d["qh.tets_seen"])
    assert np.array_equal(org, d["grid.origin"][:3])
    assert tuple(dims) == d["grid.dens"].shape
    assert np.array_equal(dens, d["grid.dens"])
    assert np.array_equal(freq, d["grid.freq"])


def test_reference_tail_loss_is_real():
    """The reference parses qhull's output file before it is flushed (voronoi.c:138-146), so it sees
    only whole 4096-byte blocks: document that the fixtures carry this, so nobody 'fixes' the oracle."""
    d = G.load("1UYD")
    assert int(d["qh.out_len"][1]) == (int(d["qh.out_len"][0]) // 4096) * 4096
    assert int(d["qh.n_seen"][0]) < len(d["qh.tets"])


def test_oracle_ligand_queries_equal_the_reference_functions():
    """dpocket (dpocket.c:228-290): fed the reference's tets, the oracle's restatement of the ligand queries - sets and counts of the
    strict dist() < d relation - must equal what the reference's own x-sorted sweeps returned (neighbor.c:192,313,490,598, atom.c:264),
    element for element once sorted, and the per-pocket criteria bit for bit."""
    d = G.load("1UYD_dpocket")
    st, p = G.inputs(d)
    r = O.search_pocket(st, p, tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    assert r["n_pockets"] == len(d["lig.crit"]) and r["n_spheres"] == len(d["sph0.f"])
    assert np.array_equal(r["iface_atoms"], np.sort(d["lig.iface"]))
    # get_mol_vert_neigh drops a sphere whose v->id was "already seen" (neighbor.c:244 is_in_lst_vert), and v->id is NOT unique: it is
    # natoms + qhull index + 1 - kept-so-far (voronoi.c:488), constant over runs of consecutive kept tets. The reference's explicit pocket
    # is therefore an order-dependent subset of the relation; ours is the relation itself. Every sphere it lacks is such a collision.
    ours, ref, ids = set(r["xspheres"].tolist()), set(d["lig.xspheres"].tolist()), d["sph0.i"][:, 7]
    assert ref <= ours
    assert {int(ids[s]) for s in ours - ref} <= {int(ids[s]) for s in ref}
    assert len({int(ids[s]) for s in ours}) == len(ref) == len({int(ids[s]) for s in ref})
    got = np.stack([r["pk_ovlp"], r["pk_dst"], r["pk_c4"], r["pk_c5"]], 1)
    assert np.array_equal(got, d["lig.crit"])


@pytest.mark.parametrize("case", G.MDCHAR_CASES)
def test_oracle_characterize_snapshot_equals_the_reference(case):
    """mdpocket characterize mode, one snapshot (mdpocket.c:436-460): fed the reference's tets and libc rand() stream, the oracle's list of zone
    spheres (surviving pockets in rank order, one entry per zone point within 1.0 of the centre: repeats included, mdpocket.c:695-745), the atoms they
    contact and the record set_descriptors makes of them - Monte-Carlo volume, hull volume and the ASA side effects on the atoms included - must be
    the reference's bit for bit. Oracle first (SURVEY par. 8 row f3): the GPU path for this mode is the next round's work."""
    d = G.load(case)
    st, p = G.inputs(d)
    r = O.characterize(st, p, d["want.points"], tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    assert r["n_spheres"] == len(d["sph0.f"]) and r["n_pockets"] == len(d["fin.off"]) - 1
    assert len(d["want.sph"]) > 50 and len(set(d["want.sph"].tolist())) < len(d["want.sph"])      # the fixture does contain repeats
    assert np.array_equal(r["want_sph"], d["want.sph"])
    assert np.array_equal(r["want_atoms"], d["want.atoms"])
    D, F, I = r["want_desc"], d["want.desc_f"], d["want.desc_i"]
    for k, n in enumerate(G.DESC_F):
        if n in ("score", "drug_score") or n.endswith("_norm"):
            continue                                  # the zone's record is never normalised or scored (mdpocket.c:453)
        assert np.float32(getattr(D, n)) == F[k] or (np.isnan(getattr(D, n)) and np.isnan(F[k])), (n, getattr(D, n), F[k])
    for k, n in enumerate(("nb_asph", "polarity_score", "charge_score", "n_abpa")):
        assert getattr(D, n) == I[k], n
    assert list(D.aa_compo) == I[11:31].tolist()
    assert D.volume > 0 and D.convex_hull_volume > 0 and D.nb_asph == len(d["want.sph"])
    assert G.eq_nan(r["atom_fx"], d["want.atom_fx"])


@pytest.mark.parametrize("case", G.ALT_CASES)
def test_oracle_non_default_clustering_equals_the_reference(case):
    """-e b (clusterlib.c:1022-1083: mean absolute difference) and -C m | a | c (complete :3705, average :3781, centroid :3337 linkage, cut by
    cuttree_distance :3235): clusters, every descriptor, drop and rank equal the reference's bit for bit, and differ from the default run's (reference
    tests/test_fpocket.py:83-91 runs "-e b -C c" and only asks for the difference). The library computes them too since round 2
    (tests/test_gpu_parity.py::test_non_default_clustering_equals_the_reference)."""
    d = G.load(case)
    want_e, want_c = (ord("b") if "_eb" in case else ord("e")), (ord(case.split("_C")[1][0]) if "_C" in case else ord("s"))
    assert int(d["params.i"][8]) == want_e and int(d["params.i"][7]) == want_c
    st, p = G.inputs(d)
    assert p.distance_measure == want_e and p.clustering_method == want_c         # fpk_params fields, taken from the reference's own parameter dump
    r = O.search_pocket(st, p, tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    assert np.array_equal(r["cl_off"], d["pre.off"]) and np.array_equal(r["cl_sph"], d["pre.sph"])
    D, F, I = r["desc"], d["pre.desc_f"], d["pre.desc_i"]
    for k, n in enumerate(G.DESC_F):
        assert G.eq_nan(D[n], F[:, k]), n
    for k, n in enumerate(G.DESC_I):
        assert np.array_equal(D[n], I[:, k]), n
    ref_first = [int(d["fin.sph"][o]) for o in d["fin.off"][:-1]]
    assert [int(r["cl_sph"][r["cl_off"][c]]) for c in r["rank_to_cluster"]] == ref_first
    p.distance_measure, p.clustering_method = ord("e"), ord("s")
    e = O.search_pocket(st, p, tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    assert e["n_clusters"] != r["n_clusters"]


def test_oracle_interface_method_2_equals_the_reference_function():
    """dpocket's second interface definition (dparams.h:44, dpocket.c:331-336): the atoms within 4.0 of a ligand atom. The oracle's set equals what
    the reference's get_mol_atm_neigh returned on the same complex (neighbor.c:67; indices of pdb_w_lig mapped to pdb by skipping the ligand), and the
    per-pocket overlap atm_corsp computes with it is the reference's bit for bit. Oracle first: the library has method 1 only."""
    d = G.load("1UYD_dpocket")
    st, p = G.inputs(d)
    O.lib().orc_set_interface_method(2)
    try:
        r = O.search_pocket(st, p, tets=d["qh.tets_seen"], rng_mode=O.RNG_LIBC)
    finally:
        O.lib().orc_set_interface_method(1)
    lig = np.sort(d["lig.atoms"])
    ref = np.sort(np.array([w - np.searchsorted(lig, w) for w in d["lig.iface2"]], dtype=np.int32))
    assert len(ref) > 20 and not set(d["lig.iface2"].tolist()) & set(lig.tolist())
    assert np.array_equal(r["iface_atoms"], ref)
    assert np.array_equal(r["pk_ovlp"], d["lig.ovlp2"])
    assert not np.array_equal(d["lig.ovlp2"], d["lig.crit"][:, 0])          # the two definitions do differ on this complex

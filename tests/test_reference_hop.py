"""CPU: the hop between OUR sphere order (canonical) and the REFERENCE's (qhull's facet order), measured and bounded on every fixture.

The parity chain is   library == oracle(canonical order)   [bit for bit, tests/test_gpu_parity.py]   and
oracle(reference order, the tets the reference read) == reference   [bit for bit, tests/test_oracle_vs_reference.py].
This file closes the remaining link with the stated tolerances of tests/refcompare.py:
  * the oracle in CANONICAL order on exactly the tets the reference read (same sphere set, other order) against the reference's dump:
    every sphere, cluster, descriptor, score, rank and per-atom effect is compared, on every fixture, nothing is skipped;
  * the oracle on the FULL triangulation (what the library computes) against the same dump: the spheres of the tail the reference
    never read (voronoi.c:138-155) are the only difference, and every cluster they do not touch is compared.
"""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O
import refcompare as RC


@pytest.mark.parametrize("case", G.FPOCKET_CASES)
def test_canonical_order_equals_reference_order_within_tolerance(case):
    d = G.load(case)
    st, p = G.inputs(d)
    r = O.search_pocket(st, p, tets=O.canonicalize(d["qh.tets_seen"]), rng_mode=O.RNG_PHILOX)
    rep = RC.compare(r, d, st, p, what=case, allow_tail=False)
    assert rep["tail"] == 0 and rep["clusters_compared"] == rep["clusters_ref"]


@pytest.mark.parametrize("case", G.FPOCKET_CASES)
def test_full_triangulation_equals_reference_outside_the_unread_tail(case):
    d = G.load(case)
    st, p = G.inputs(d)
    r = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    rep = RC.compare(r, d, st, p, what=case, allow_tail=True)
    assert rep["clusters_compared"] >= rep["clusters_ref"] - 2 * rep["tail"] - 1


def test_the_tail_loss_touches_seven_fixtures_and_none_is_skipped():
    """documents which fixtures carry tail spheres (they used to be compared with nothing)"""
    n_tail = {}
    for case in G.FPOCKET_CASES:
        d = G.load(case)
        st, p = G.inputs(d)
        r = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        n_tail[case] = r["n_spheres"] - len(d["sph0.f"])
    assert sum(1 for v in n_tail.values() if v > 0) >= 5 and max(n_tail.values()) <= 32, n_tail


@pytest.mark.parametrize("case", G.MD_CASES)
def test_md_grids_in_canonical_order_against_the_reference(case):
    d = G.load(case)
    st, p = G.inputs(d)
    score = int(case.endswith("_S"))
    org, dims, dens, freq, r = O.md_grids(st, p, use_drug_score=score, tets=O.canonicalize(d["qh.tets_seen"]))
    assert np.array_equal(org, d["grid.origin"][:3]) and tuple(dims) == d["grid.dens"].shape
    if not score:
        rep = RC.compare_md_grids(dens, freq, d, r["n_pockets"], what=case)
        assert rep["dens_cells_moved"] <= 2 * r["n_pockets"]
    else:       # -S: the increments are the pockets' drug scores (floats in [0, 1]): same stencils, the same centre-cell rule
        dd, df = np.abs(dens - d["grid.dens"]), np.abs(freq - d["grid.freq"])
        assert dd.max() <= 1.0 + 1e-5 and df.max() <= 1.0 + 1e-5
        assert np.count_nonzero(dd > 1e-5) <= 2 * r["n_pockets"] and np.count_nonzero(df > 1e-5) <= 2 * r["n_pockets"]
        assert np.isclose(dens.sum(), d["grid.dens"].sum(), rtol=1e-5)


@pytest.mark.parametrize("name", G.SAMPLES)
def test_every_reference_sample_against_the_reference_run(name):
    """All 37 readable structures of the reference's data/sample (PDB and mmCIF twins, up to 26,677 atoms; two *_wrote.cif files are rejected
    by the reference's own reader), atoms as the reference parsed them, default parameters: what the library computes (the oracle on its
    own full triangulation, canonical order) against the unmodified reference's run, under the tolerances of tests/refcompare.py."""
    d = G.load_sample(name)
    st, p = G.inputs(d)
    r = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    rep = RC.compare(r, d, st, p, what=name)
    assert rep["clusters_compared"] >= rep["clusters_ref"] - 2 * rep["tail"]

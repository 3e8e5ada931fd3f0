"""CPU: the product's one deliberate saving - no Monte-Carlo volume and no hull for clusters below min_pock_nb_asph, which
dropSmallNpolarPockets always removes (reference src/refine.c:124) - changes nothing a user of the reference can observe.
The oracle in ORC_RNG_LIBC mode is the reference bit for bit (test_oracle_vs_reference.py); in ORC_RNG_PHILOX mode it is what
the CUDA path computes. Both must agree on everything but the Monte-Carlo volume for every pocket that survives."""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O


@pytest.mark.parametrize("case", G.FPOCKET_CASES)
def test_surviving_pockets_do_not_depend_on_the_saving(case):
    d = G.load(case)
    st, p = G.inputs(d)
    a = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_LIBC)
    b = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    assert a["status"] == b["status"] and a["n_pockets"] == b["n_pockets"]
    assert np.array_equal(a["rank_to_cluster"], b["rank_to_cluster"])
    for k in ("sph_xyzr", "sph_neigh", "cl_off", "cl_sph", "cl_atoms", "atom_fx"):
        assert G.eq_nan(a[k], b[k]), k
    keep = a["rank_to_cluster"]
    small = a["desc"]["nb_asph"] < p.min_pock_nb_asph
    assert not small[keep].any()
    for n in G.DESC_F:
        if n == "volume":
            continue
        assert G.eq_nan(a["desc"][n][keep], b["desc"][n][keep]), n
        if n not in ("convex_hull_volume", "score"):            # unobservable pockets differ only where the hull enters
            assert G.eq_nan(a["desc"][n], b["desc"][n]), n
    assert np.all(b["desc"]["volume"][small] == 0) and np.all(b["desc"]["convex_hull_volume"][small] == 0)
    # the Monte-Carlo estimate itself: two independent streams agree within the sampling error (4 sigma)
    va, vb = a["desc"]["volume"][keep], b["desc"]["volume"][keep]
    for c, x, y in zip(keep, va, vb):
        sph = a["sph_xyzr"][a["cl_sph"][a["cl_off"][c]:a["cl_off"][c + 1]]]
        lo, hi = (sph[:, :3] - sph[:, 3:4]).min(0), (sph[:, :3] + sph[:, 3:4]).max(0)
        vbox = float(np.prod(hi - lo))
        pin = min(max(0.5 * (x + y) / vbox, 1e-3), 0.999)
        sigma = vbox * np.sqrt(pin * (1 - pin) / (100 * p.nb_mcv_iter))
        assert abs(x - y) <= 4 * np.sqrt(2) * sigma + 1e-3 * vbox, (c, x, y, sigma)

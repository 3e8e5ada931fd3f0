"""CPU: an independent check of the oracle's complete- and average-linkage clustering (-C m, -C a; reference clusterlib.c:3705,3781 + cuttree_distance :3235,
restated in oracle/oracle.c linkage_labels and pinned bit for bit to the reference's own runs by tests/test_oracle_vs_reference.py). Both linkages are monotone, so
cutting the dendrogram at clust_max_dist is a partition any correct implementation must agree on when no two candidate merges tie: scipy's linkage + fcluster is one.
(Centroid linkage is not monotone and the reference's cut rule is its own - highest-numbered ancestor at or below the cut -, so only the fixtures pin it.)"""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O

scipy_h = pytest.importorskip("scipy.cluster.hierarchy")


def _partition(labels):
    groups = {}
    for i, c in enumerate(labels):
        groups.setdefault(int(c), []).append(i)
    return sorted(tuple(v) for v in groups.values())


@pytest.mark.parametrize("case,method", [("1UYD", "m"), ("1UYD", "a"), ("3LKF", "m"), ("5WA6", "a")])
def test_monotone_linkages_equal_scipy(case, method):
    d = G.load(case)
    st, p = G.inputs(d)
    p.clustering_method = ord(method)
    r = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
    xyz = r["sph_xyzr"][:, :3].astype(np.float64)
    Z = scipy_h.linkage(xyz, method={"m": "complete", "a": "average"}[method], metric="euclidean")
    want = scipy_h.fcluster(Z, t=float(np.float32(p.clust_max_dist)), criterion="distance")
    assert _partition(r["sph_cluster"]) == _partition(want)
    assert r["n_clusters"] == len(set(want.tolist())) > 100
    # and the numbering is the one apply_clustering leaves (refine.c:58-93): clusters in order of their lowest sphere
    first = [int(np.flatnonzero(r["sph_cluster"] == c)[0]) for c in range(r["n_clusters"])]
    assert first == sorted(first)

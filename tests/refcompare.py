"""The hop from OUR order to the REFERENCE's order, as a tested tolerance (DESIGN.md par. 5).

The library (and the oracle when it triangulates itself) lists alpha spheres in canonical order; the reference lists them in
qhull's facet order, which cannot be reproduced. Everything that depends on that order is compared here against the stage
dumps of the unmodified reference (tests/golden/*.npz, oracle/ref_dump.c), for EVERY fixture, under these stated criteria:

  spheres          reference set is a subset of ours; what it lacks lies in the tail of qhull's output it never read
                   (voronoi.c:138-155); centres within 1e-6 relative (measured: identical bits), radii within 1e-4 A
                   (the radius is the f32 distance to neigh[0], and qhull's first vertex is another atom than ours)
  types            equal
  clusters         every cluster of ours that holds no tail sphere IS a cluster of the reference (same sphere set,
                   same contacted-atom set); on fixtures without tail spheres the two partitions are identical
  integer fields   nb_asph, n_abpa, apolar / polar counts: equal; polarity, charge, amino-acid composition (and the two float means
                   over residues): equal except on pockets that touch two different residues with the same number in two
                   chains - descriptors.c:408 keeps the first one met, which follows the sphere order
  float fields     F_RTOL relative (float sums run over another order of the same terms)
  hull volume      F_RTOL (measured: identical bits, it is an exact integer converted once)
  score            SCORE_ATOL absolute + F_RTOL relative (it sums float terms of the order of the areas), drug score DRUG_RTOL
  MC volume        statistical: |V - V_ref| <= MC_SIGMAS sigma, sigma from the box volume and 100 x nb_mcv_iter samples
                   (the reference seeds libc rand() with the time: voronoi.c:87). The sampling box itself depends on the sphere
                   order (else-if in voronoi.c:1172-1195): where the two orders give different boxes (2 of 87 pockets of the
                   7TAA fixture) each estimate is checked against the volume inside its own box
  ranking          the surviving pockets, as sphere sets, in the same order
  per-atom ASA     atoms contacted by ONE cluster: equal; atoms contacted by several: the last cluster in list order wins
                   (asa.c:337,392-396), list order is qhull's there and canonical here: they may differ, on at most
                   FX_MAX_FRACTION of the atoms
Clusters below min_pock_nb_asph are always dropped (refine.c:124); the library skips their Monte-Carlo and hull volume
(DESIGN.md par. 4), so volume, hull volume and score are compared on the others only.
"""
import numpy as np

import golden_cases as G
import oracle_lib as O

F_RTOL = 5e-5            # measured worst: 3.1e-5 (masph_sacc, 5WA6)
NORM_ATOL = 1e-5         # min-max normalised fields, values in [0, 1]; measured worst 1.7e-6
SCORE_ATOL = 5e-6        # measured worst 1.9e-6
DRUG_RTOL = 5e-6         # measured worst 1.9e-6
RADIUS_ATOL = 1e-4       # measured worst 6.1e-5
MC_SIGMAS = 6.0
FX_MAX_FRACTION = 0.20   # measured worst 13.9 % (7TAA -m 2.8 -M 7.4)

NORMALISED = ["flex", "nas_norm", "polarity_score_norm", "mean_loc_hyd_dens_norm", "prop_asapol_norm", "as_density_norm", "as_max_dst_norm"]
VOLUME_LIKE = ["volume", "convex_hull_volume", "score", "drug_score"]
PLAIN = [n for n in G.DESC_F if n not in NORMALISED and n not in VOLUME_LIKE]
RESIDUE_BASED = ["polarity_score", "charge_score", "hydrophobicity_score", "volume_score"]      # + aa_compo


def quirk_box(xyzr):
    """the Monte-Carlo box of voronoi.c:1172-1195, in f32: +correct on both sides, and the max is only looked at when the min did not
    move (else-if), so the box depends on the ORDER of the spheres and can cut off part of the pocket"""
    f = np.float32
    c = f(-1.6)
    lo, hi = [f(0)] * 3, [f(0)] * 3
    for i, (x, y, z, r) in enumerate(np.asarray(xyzr, dtype=np.float32)):
        for a, v in enumerate((x, y, z)):
            mn, mx = f(f(v - r) + c), f(f(v + r) + c)
            if i == 0:
                lo[a], hi[a] = mn, mx
            elif lo[a] > mn:
                lo[a] = mn
            elif hi[a] < mx:
                hi[a] = mx
    return np.array(lo, np.float64), np.array(hi, np.float64)


def box_volume_estimate(xyzr, lo, hi, nsamp=120000, seed=7):
    """volume of (union of the spheres with radius ray - 1.6) inside the box, by plain numpy sampling: (estimate, its sigma)"""
    rng = np.random.default_rng(seed)
    pts = lo + rng.random((nsamp, 3)) * (hi - lo)
    c, rr = xyzr[:, :3].astype(np.float64), (xyzr[:, 3].astype(np.float64) - 1.6)
    inside = np.zeros(nsamp, bool)
    for k in range(len(c)):
        if rr[k] <= 0:
            continue
        inside |= ((pts - c[k]) ** 2).sum(1) < rr[k] ** 2
    vbox = float(np.prod(hi - lo))
    pr = inside.mean()
    return vbox * pr, vbox * np.sqrt(max(pr * (1 - pr), 1e-12) / nsamp)


def mc_sigma(vbox, v, niter):
    ns = 100.0 * niter
    pr = min(max(v / max(vbox, 1e-9), 1.0 / ns), 1.0)
    return vbox * np.sqrt(pr * (1 - pr) / ns + 1.0 / ns ** 2)


def compare(r, d, st, p, what="", allow_tail=True):
    """r: a result in canonical order (library or oracle); d: the reference's dump of the same input. Returns a dict of what was compared."""
    ours = {k: i for i, k in enumerate(G.canon_sphere_keys(r["sph_neigh"]))}
    ref_keys = G.canon_sphere_keys(d["sph0.i"][:, :4])
    assert len(set(ref_keys)) == len(ref_keys), what
    assert set(ref_keys) <= set(ours), (what, "the reference kept a sphere we do not have")
    idx = np.array([ours[k] for k in ref_keys], dtype=np.int64)
    extra = sorted(set(ours.values()) - set(idx.tolist()))
    if extra:
        assert allow_tail, (what, "extra spheres")
        n_seen = int(d["qh.n_seen"][0])
        s2a, _, _ = O.sites(st)
        lost = {tuple(sorted(int(s2a[v]) for v in t)) for t in d["qh.tets"][n_seen - 1:]}
        keys = G.canon_sphere_keys(r["sph_neigh"])
        assert {keys[i] for i in extra} <= lost, (what, "extra spheres that the reference's unread tail does not explain")
    # ---- spheres ------------------------------------------------------------------------------------------------------------
    assert np.allclose(r["sph_xyzr"][idx, :3], d["sph0.f"][:, :3], rtol=1e-6, atol=0), what
    assert np.allclose(r["sph_xyzr"][idx, 3], d["sph0.f"][:, 3], rtol=0, atol=RADIUS_ATOL), what
    assert np.array_equal(r["sph_type"][idx], d["sph0.i"][:, 4]), what
    # ---- clusters: ours without a tail sphere <-> the reference's --------------------------------------------------------------
    nref = len(d["pre.off"]) - 1
    ref_sets = {frozenset(idx[d["pre.sph"][d["pre.off"][q]:d["pre.off"][q + 1]]].tolist()): q for q in range(nref)}
    xs = set(extra)
    clean, refq = [], []
    for c in range(r["n_clusters"]):
        s = frozenset(r["cl_sph"][r["cl_off"][c]:r["cl_off"][c + 1]].tolist())
        if s & xs:
            continue
        assert s in ref_sets, (what, "a cluster without tail spheres is not a cluster of the reference", c)
        clean.append(c)
        refq.append(ref_sets[s])
    clean, refq = np.array(clean, dtype=np.int64), np.array(refq, dtype=np.int64)
    if not extra:
        assert len(clean) == nref == r["n_clusters"], what
    assert len(clean) >= nref - 2 * len(extra), (what, "too few comparable clusters", len(clean), nref)      # a tail sphere spoils at most the clusters it joins
    D, F, I = r["desc"], d["pre.desc_f"], d["pre.desc_i"]
    coord_ulp = float(np.spacing(np.float32(max(np.abs(st.x).max(), np.abs(st.y).max(), np.abs(st.z).max()))))
    # Residues are made unique by res_id ALONE (descriptors.c:408, the chain is ignored): when a pocket touches two different residues
    # that carry the same number (two chains), the one met first in the atom list stands for both, and the atom list follows the
    # sphere order. On such pockets the residue-based fields are order-dependent and left out; everywhere else they must be equal.
    ambiguous = np.zeros(len(clean), bool)
    for j, c in enumerate(clean):
        at = r["cl_atoms"][r["cl_atom_off"][c]:r["cl_atom_off"][c + 1]]
        pairs = {(int(st.res_id[a]), int(st.aa_idx[a])) for a in at}
        ambiguous[j] = len({rid for rid, _ in pairs}) < len(pairs)
    sure = ~ambiguous
    assert len(clean) == 0 or ambiguous.mean() <= 0.25, (what, "pockets with two residues of one number", float(ambiguous.mean()))
    for k, n in enumerate(G.DESC_I):
        m = sure if n in RESIDUE_BASED else np.ones(len(clean), bool)
        assert np.array_equal(D[n][clean][m], I[refq, k][m]), (what, n)
    assert np.array_equal(D["aa_compo"][clean][sure], I[refq, 11:31][sure]), what
    for n in PLAIN:
        m = sure if n in RESIDUE_BASED else np.ones(len(clean), bool)
        a, b = D[n][clean][m], F[refq, G.DESC_F.index(n)][m]
        # masph_sacc = mean |centre - barycentre of the 4 atoms| / radius: the barycentre is an f32 sum of the atoms in vertex order
        # (voronoi.c:892-916), so it moves by an ulp of the COORDINATE (3e-5 A at 280 A from the origin) with the order
        atol = 2.0 * coord_ulp if n == "masph_sacc" else 1e-6
        assert np.allclose(a, b, rtol=F_RTOL, atol=atol, equal_nan=True), (what, n, float(np.nanmax(np.abs(a - b))))
    assert np.allclose(D["bary"][clean], F[refq, 27:30], rtol=F_RTOL, atol=max(1e-5, 2.0 * coord_ulp)), what
    # contacted atoms as sets (first-seen order follows the sphere order)
    for c in clean[:64]:
        s = frozenset(r["cl_sph"][r["cl_off"][c]:r["cl_off"][c + 1]].tolist())
        mine = set(r["cl_atoms"][r["cl_atom_off"][c]:r["cl_atom_off"][c + 1]].tolist())
        assert mine == {int(a) for i in s for a in r["sph_neigh"][i]}, what
    # ---- volume, hull, scores: clusters that can survive the drop ----------------------------------------------------------------
    single = r["n_clusters"] == 1
    obs = (D["nb_asph"][clean] >= p.min_pock_nb_asph) | single
    co, ro = clean[obs], refq[obs]
    full = bool(p.do_asa_and_volume)
    n_boxquirk = 0
    if full:
        a, b = D["convex_hull_volume"][co], F[ro, G.DESC_F.index("convex_hull_volume")]
        assert np.allclose(a, b, rtol=F_RTOL, atol=1e-4), (what, "hull volume")
        kv = G.DESC_F.index("volume")
        for c, q in zip(co, ro):
            mine = r["sph_xyzr"][r["cl_sph"][r["cl_off"][c]:r["cl_off"][c + 1]]]
            theirs = r["sph_xyzr"][idx[d["pre.sph"][d["pre.off"][q]:d["pre.off"][q + 1]]]]        # the same spheres in the reference's order
            (lo_a, hi_a), (lo_b, hi_b) = quirk_box(mine), quirk_box(theirs)
            va, vb = float(D["volume"][c]), float(F[q, kv])
            if np.array_equal(lo_a, lo_b) and np.array_equal(hi_a, hi_b):
                sigma = mc_sigma(float(np.prod(hi_a - lo_a)), max(va, vb), p.nb_mcv_iter)
                assert abs(va - vb) <= MC_SIGMAS * np.sqrt(2.0) * sigma, (what, "MC volume", int(c), va, vb, sigma)
            else:
                # the else-if of voronoi.c:1172-1195 makes the sampling box depend on the sphere order: each estimate is checked against
                # the volume inside ITS OWN box (numpy, 120k samples)
                n_boxquirk += 1
                for v, lo, hi in ((va, lo_a, hi_a), (vb, lo_b, hi_b)):
                    est, es = box_volume_estimate(mine, lo, hi)
                    sigma = mc_sigma(float(np.prod(hi - lo)), est, p.nb_mcv_iter)
                    assert abs(v - est) <= MC_SIGMAS * np.hypot(sigma, es), (what, "MC volume in its own box", int(c), v, est, sigma, es)
    if not extra:       # the min-max normalisation runs over ALL clusters: only comparable when the partitions are identical
        for n in NORMALISED:
            a, b = D[n][clean], F[refq, G.DESC_F.index(n)]
            if n == "polarity_score_norm" and ambiguous.any():
                continue        # its bounds run over every pocket's polarity_score, the order-dependent ones (above) included
            if n == "as_density_norm" and (np.all(np.isnan(a)) != np.all(np.isnan(b))):
                # pocket.c:664-785 scans the clusters in list order; when the FIRST one is a single sphere its as_density is 0/0 and the NaN
                # poisons both bounds, i.e. the whole column (nothing reads it afterwards: score and drug score use other fields). Which
                # cluster comes first is qhull's order there, canonical here.
                continue
            assert np.allclose(a, b, rtol=F_RTOL, atol=NORM_ATOL, equal_nan=True), (what, n)
        a, b = D["score"][co], F[ro, 0]
        assert np.allclose(a, b, rtol=F_RTOL, atol=SCORE_ATOL), (what, "score", float(np.abs(a - b).max()) if len(a) else 0)
        a, b = D["drug_score"][co], F[ro, 1]
        assert np.allclose(a, b, rtol=DRUG_RTOL, atol=1e-9), (what, "drug score")
    # ---- drop + rank ------------------------------------------------------------------------------------------------------------------
    fin_sets = [frozenset(idx[d["fin.sph"][d["fin.off"][q]:d["fin.off"][q + 1]]].tolist()) for q in range(len(d["fin.off"]) - 1)]
    my_sets = [frozenset(r["cl_sph"][r["cl_off"][c]:r["cl_off"][c + 1]].tolist()) for c in r["rank_to_cluster"]]
    if not extra:
        assert my_sets == fin_sets, (what, "ranking")
    else:
        # pockets untouched by the tail: same survivors; their relative order can only change through the normalisation bounds,
        # so it is required only where the two scores differ by more than the normalised terms can move (0.485 x nas_norm)
        mine = [s for s in my_sets if not (s & xs)]
        theirs = [s for s in fin_sets if s in set(mine)]
        assert set(mine) <= set(fin_sets) | set(), (what, "a clean surviving pocket that the reference dropped")
        assert len(theirs) == len(mine), what
    # ---- per-atom ASA side effects ----------------------------------------------------------------------------------------------------
    n_ctd = np.zeros(st.natoms, np.int64)
    for c in range(r["n_clusters"]):
        n_ctd[r["cl_atoms"][r["cl_atom_off"][c]:r["cl_atom_off"][c + 1]]] += 1
    touched = np.zeros(st.natoms, bool)           # atoms of the clusters a tail sphere joined: their areas see one more sphere
    cl = set(clean.tolist())
    for c in range(r["n_clusters"]):
        if c not in cl:
            touched[r["cl_atoms"][r["cl_atom_off"][c]:r["cl_atom_off"][c + 1]]] = True
    fx, rf = r["atom_fx"], d["atom_fx"]
    differs = ~(np.isclose(fx, rf, rtol=F_RTOL, atol=1e-5) | (np.isnan(fx) & np.isnan(rf)))
    rows = differs.any(1)
    if full:
        # the radius of a sphere differs in its last bits between the two vertex orders (RADIUS_ATOL above), so one of the 200 spiral points
        # of an atom can change sides of `ddist <= ray * ray` (asa.c:306-312): seen once in 26,677 atoms (6cs2); anything more is an error
        lone = int(np.count_nonzero(rows & (n_ctd <= 1) & ~touched))
        assert lone <= max(1, st.natoms // 10000), (what, "per-atom ASA effect differs on atoms that only one cluster contacts", lone)
        assert rows.mean() <= FX_MAX_FRACTION, (what, rows.mean())
    return {"spheres": len(ref_keys), "tail": len(extra), "clusters_compared": int(len(clean)), "clusters_ref": nref, "fx_rows_differing": int(rows.sum()), "mc_box_order_dependent": n_boxquirk}


def compare_md_grids(dens, freq, d, n_pockets, what=""):
    """mdpocket grids of one frame against the reference's (integer increments). The 26-stencil and the frequency stencil are order-free;
    the once-per-pocket centre cell uses the pocket's LAST sphere (mdpbase.c:164-166,255-260), which is canonical-last here and qhull-last
    there: at most one cell per pocket moves in each grid, by one count, and the density total is conserved."""
    rd, rf = d["grid.dens"], d["grid.freq"]
    assert dens.shape == rd.shape, what
    dd, df = dens - rd, freq - rf
    assert np.all(np.abs(dd) <= 1) and np.all(np.abs(df) <= 1), what
    assert int(np.count_nonzero(dd)) <= 2 * n_pockets and int(np.count_nonzero(df)) <= 2 * n_pockets, (what, np.count_nonzero(dd), np.count_nonzero(df), n_pockets)
    assert float(dens.sum()) == float(rd.sum()), what
    return {"dens_cells_moved": int(np.count_nonzero(dd)), "freq_cells_moved": int(np.count_nonzero(df))}

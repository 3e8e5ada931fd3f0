"""Helpers shared by the CPU and GPU parity tests: load a golden fixture (tests/golden/*.npz, made by
tests/golden/make_golden.py from the unmodified reference) and compare results field by field."""
import glob
import os

import numpy as np

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
MDCHAR_CASES = [c for c in _FILES if c.startswith("mdchar_")]      # one snapshot of mdpocket's characterize mode (oracle-only so far)
ALT_CASES = [c for c in _FILES if c.startswith("alt_")]              # non-default clustering distance (-e b): on the GPU since round 2
ALL = [c for c in _FILES if c not in MDCHAR_CASES]
FPOCKET_CASES = [c for c in ALL if not c.startswith("md_")]
MD_CASES = [c for c in ALL if c.startswith("md_")]

DESC_F = ["score", "drug_score", "hydrophobicity_score", "volume_score", "volume", "convex_hull_volume",
          "prop_polar_atm", "mean_asph_ray", "masph_sacc", "apolar_asphere_prop", "mean_loc_hyd_dens", "as_density",
          "as_max_dst", "flex", "nas_norm", "polarity_score_norm", "mean_loc_hyd_dens_norm", "prop_asapol_norm",
          "as_density_norm", "as_max_dst_norm", "surf_vdw14", "surf_vdw22", "surf_pol_vdw14", "surf_pol_vdw22",
          "surf_apol_vdw14", "surf_apol_vdw22", "as_max_r"]
DESC_I = ["nb_asph", "polarity_score", "charge_score", "n_abpa", "nAlphaApol", "nAlphaPol"]


def load(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    n = int(d["qh.n_seen"][0])
    seen = d["qh.tets"][:n].copy()
    seen[n - 1] = d["qh.last_seen"]
    d["qh.tets_seen"] = seen            # exactly what the reference parsed (voronoi.c:395-445, unflushed tail lost)
    return d


def inputs(d, **kw):
    return O.structure_from_dump(d), O.params_from_dump(d, mc_seed=int(d["seed"][0]), **kw)


def eq_nan(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def canon_sphere_keys(neigh):
    """sorted 4-tuples of contacted atoms: the order-free identity of an alpha sphere"""
    return [tuple(sorted(int(v) for v in row)) for row in neigh]


def pocket_sets(r, clusters):
    """set of frozensets of sphere keys for the given cluster indices"""
    keys = canon_sphere_keys(r["sph_neigh"])
    return [frozenset(keys[s] for s in r["cl_sph"][r["cl_off"][c]:r["cl_off"][c + 1]]) for c in clusters]


# ---- every structure of the reference's data/sample, atoms as the reference's readers parsed them (tests/golden/make_sample_fixtures.py) ----
SAMPLES_DIR = os.path.join(GOLD, "samples")
SAMPLES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(SAMPLES_DIR, "*.npz")))


def load_sample(name):
    """A sample fixture in the shape of load(): pdb's atoms are the masked subsequence of pdb_w_lig's; of qhull's tetrahedra only the
    tail the reference never read is kept (d["qh.tets"][d["qh.n_seen"] - 1:] is that tail, as in the full fixtures)."""
    d = dict(np.load(os.path.join(SAMPLES_DIR, name + ".npz")))
    m = d["in_pdb"].astype(bool)
    for k in ("f", "i", "symbol", "res_name"):
        d["pdb." + k] = d["pdbw." + k][m]
    d["qh.tets"] = d["qh.tail"]
    d["qh.n_seen"] = np.array([1])
    return d

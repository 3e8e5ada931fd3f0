"""Test-side loader of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from fpocket_b200 import capi
from fpocket_b200.hostprep import structure_from_arrays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None

RNG_LIBC, RNG_PHILOX = 0, 1


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "oracle.h", "oracle_delaunay.inc")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        ip, fp, dp, llp = (C.POINTER(t) for t in (C.c_int32, C.c_float, C.c_double, C.c_longlong))
        L.orc_sites.restype = C.c_int
        L.orc_sites.argtypes = [C.POINTER(capi.FpkStructure), ip, llp, dp]
        L.orc_delaunay.restype = C.c_int
        L.orc_delaunay.argtypes = [C.c_int, llp, C.POINTER(ip), ip, dp]
        L.orc_canonicalize_tets.argtypes = [ip, C.c_int]
        L.orc_search_pocket.restype = C.c_int
        L.orc_search_pocket.argtypes = [C.POINTER(capi.FpkStructure), C.POINTER(capi.FpkParams), ip, C.c_int, C.c_int,
                                        C.POINTER(capi.FpkResult)]
        L.orc_result_free.argtypes = [C.POINTER(capi.FpkResult)]
        L.orc_md_grid_geometry.argtypes = [C.POINTER(capi.FpkResult), fp, ip]
        L.orc_md_scatter.argtypes = [C.POINTER(capi.FpkResult), C.c_int, fp, ip, fp, fp, fp]
        L.orc_circumcentre.argtypes = [dp, dp, dp, dp, dp]
        L.orc_orient3d.restype = C.c_int
        L.orc_orient3d.argtypes = [llp] * 4
        L.orc_insphere.restype = C.c_int
        L.orc_insphere.argtypes = [llp] * 5
        L.orc_spiral_points.argtypes = [fp]
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L._free = C.CDLL(None).free
        L._free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def sites(st):
    L = lib()
    n = st.natoms
    s2a = np.zeros(n + 1, np.int32)
    ki = np.zeros(3 * n + 3, np.int64)
    xd = np.zeros(3 * n + 3, np.float64)
    cs = st.c_struct()
    m = L.orc_sites(C.byref(cs), s2a.ctypes.data_as(C.POINTER(C.c_int32)), ki.ctypes.data_as(C.POINTER(C.c_longlong)),
                    xd.ctypes.data_as(C.POINTER(C.c_double)))
    return s2a[:m].copy(), ki[:3 * m].reshape(m, 3).copy(), xd[:3 * m].reshape(m, 3).copy()


def delaunay(ki):
    """ki: [n,3] int64 lattice coordinates -> (tets [T,4] canonical, hull volume)."""
    L = lib()
    ki = np.ascontiguousarray(ki, dtype=np.int64)
    out = C.POINTER(C.c_int32)()
    nt = C.c_int32()
    vol = C.c_double()
    rc = L.orc_delaunay(len(ki), ki.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(out), C.byref(nt), C.byref(vol))
    if rc:
        return rc, None, 0.0
    t = np.ctypeslib.as_array(out, shape=(nt.value * 4,)).reshape(-1, 4).copy() if nt.value else np.zeros((0, 4), np.int32)
    L._free(out)
    return 0, t, vol.value


def canonicalize(tets):
    t = np.ascontiguousarray(tets, dtype=np.int32).copy()
    lib().orc_canonicalize_tets(t.ctypes.data_as(C.POINTER(C.c_int32)), len(t))
    return t


def characterize(st, params, zone, tets=None, rng_mode=RNG_PHILOX):
    """One snapshot of mdpocket's characterize mode through the oracle (oracle.c orc_set_wanted_zone): the usual result plus the zone's sphere
    list (with repeats), its contacted atoms and its descriptor record."""
    L = lib()
    zone = np.ascontiguousarray(zone, dtype=np.float32)
    L.orc_set_wanted_zone.argtypes = [C.POINTER(C.c_float), C.c_int]
    L.orc_set_wanted_zone.restype = None
    L.orc_get_wanted.argtypes = [C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int), C.POINTER(capi.FpkDesc)]
    L.orc_get_wanted.restype = C.c_int
    L.orc_set_wanted_zone(zone.ctypes.data_as(C.POINTER(C.c_float)), len(zone))
    try:
        r = search_pocket(st, params, tets=tets, rng_mode=rng_mode)
        idx, atoms = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        ni, na = C.c_int(), C.c_int()
        d = capi.FpkDesc()
        L.orc_get_wanted(C.byref(idx), C.byref(ni), C.byref(atoms), C.byref(na), C.byref(d))
        r["want_sph"] = np.array([idx[i] for i in range(ni.value)], dtype=np.int32)
        r["want_atoms"] = np.array([atoms[i] for i in range(na.value)], dtype=np.int32)
        r["want_desc"] = d
    finally:
        L.orc_set_wanted_zone(None, 0)
    return r


def search_pocket(st, params, tets=None, rng_mode=RNG_PHILOX):
    """Run the oracle pipeline. tets=None -> the oracle's own exact Delaunay (canonical order)."""
    L = lib()
    cs = st.c_struct()
    res = capi.FpkResult()
    if tets is not None:
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        tp, nt = tets.ctypes.data_as(C.POINTER(C.c_int32)), len(tets)
    else:
        tp, nt = None, 0
    L.orc_search_pocket(C.byref(cs), C.byref(params), tp, nt, rng_mode, C.byref(res))
    d = capi.result_to_dict(res, st.natoms)
    L.orc_result_free(C.byref(res))
    return d


def md_grids(st, params, use_drug_score=0, origin=None, dims=None, tets=None):
    """One frame of mdpocket through the oracle: detection + grids. Returns (origin, dims, dens, freq)."""
    L = lib()
    cs = st.c_struct()
    res = capi.FpkResult()
    if tets is not None:
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        tp, nt = tets.ctypes.data_as(C.POINTER(C.c_int32)), len(tets)
    else:
        tp, nt = None, 0
    L.orc_search_pocket(C.byref(cs), C.byref(params), tp, nt, RNG_PHILOX, C.byref(res))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    if origin is None:
        origin = np.zeros(3, np.float32)
        dims = np.zeros(3, np.int32)
        L.orc_md_grid_geometry(C.byref(res), origin.ctypes.data_as(fp), dims.ctypes.data_as(ip))
    origin = np.ascontiguousarray(origin, np.float32)
    dims = np.ascontiguousarray(dims, np.int32)
    n = int(np.prod(dims))
    dens, freq, refg = (np.zeros(n, np.float32) for _ in range(3))
    if res.n_pockets > 0:
        L.orc_md_scatter(C.byref(res), use_drug_score, origin.ctypes.data_as(fp), dims.ctypes.data_as(ip),
                         dens.ctypes.data_as(fp), freq.ctypes.data_as(fp), refg.ctypes.data_as(fp))
    out = capi.result_to_dict(res, st.natoms)
    L.orc_result_free(C.byref(res))
    return origin, dims, dens.reshape(dims), freq.reshape(dims), out


def structure_from_dump(d):
    """Build the flat input from an oracle/ref_dump.c dump (atoms exactly as the reference parsed them)."""
    md = int(d["md"][0])
    w = None if md else (d["pdbw.f"], d["pdbw.symbol"])
    xlig = d.get("pdb.xlig")
    extra = None
    if "params.xpocket" in d:
        xp = d["params.xpocket"]
        extra = np.zeros(len(d["pdb.f"]), np.uint8)
        for k, (rid, ch, ins) in enumerate(zip(d["pdb.i"][:, 1], d["pdb.chain"][:, 0], d["pdb.ins_aloc"][:, 0])):
            for (xr, xc, xi) in xp:
                if xr == rid and xc == ch and (xi == ins or (xi == ord("-") and ins in (ord(" "), 0))):
                    extra[k] |= capi.FPK_ATOM_XPOCKET
    st = structure_from_arrays(d["pdb.f"], d["pdb.i"], d["pdb.symbol"], d["pdb.res_name"], d["pdb.bfstat"], w=w,
                               same_pdb=md, xlig=xlig, extra_flags=extra)
    if "lig.atoms" in d:            # dpocket fixture: the known ligand (indices into pdb_w_lig), atom names packed for atm_corsp
        nk = np.ascontiguousarray(d["pdb.name"][:, :4]).view(np.int32).reshape(-1) if "pdb.name" in d else None
        st.set_ligand(d["lig.atoms"], d["lig.mol"], nk)
    return st


def params_from_dump(d, **kw):
    pf, pi = d["params.f"], d["params.i"]
    p = capi.default_params()
    p.asph_min_size, p.asph_max_size, p.clust_max_dist = pf[0], pf[1], pf[2]
    p.refine_min_apolar_asphere_prop, p.min_as_density = pf[3], pf[4]
    p.min_apol_neigh, p.nb_mcv_iter, p.min_pock_nb_asph, p.do_asa_and_volume = (int(v) for v in pi[:4])
    p.skip_clustering = int(not (pi[4] == -1 and pi[5] == 0))
    p.xpocket_active = int(pi[5] > 0)
    p.min_n_explicit_pocket_atoms = int(pi[6])
    if len(pi) > 8 and int(pi[8]) in (ord("e"), ord("b")):
        p.distance_measure = int(pi[8])          # -e (fparams.c): 'e' euclid, 'b' city block
        p.clustering_method = int(pi[7])         # -C: 's' single, 'm' complete, 'a' average, 'c' centroid linkage
    for k, v in kw.items():
        setattr(p, k, v)
    return p

"""ctypes mirror of include/fpocket_cuda.h (POD layouts only) and the loader of libfpocket_cuda.so.

The product path is the CUDA library; there is no CPU fallback. `load_library()` raises if the
shared object is missing, and `fpk_init` fails with FPK_ERR_NO_DEVICE without a usable sm_100 GPU.
"""
import ctypes as C
import os

import numpy as np

FPK_OK = 0
FPK_ERR_NO_DEVICE = -1
FPK_ERR_BAD_ARG = -2
FPK_ERR_CUDA = -3
FPK_ERR_NO_SPHERES = -4
FPK_ERR_DEGENERATE = -5
FPK_ERR_CAPACITY = -6
FPK_ERR_NO_POCKETS = -7

FPK_ATOM_H = 1
FPK_ATOM_XPOCKET = 2
FPK_ATOM_H1 = 4

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)


class FpkParams(C.Structure):
    _fields_ = [
        ("asph_min_size", C.c_float), ("asph_max_size", C.c_float), ("clust_max_dist", C.c_float),
        ("refine_min_apolar_asphere_prop", C.c_float), ("min_as_density", C.c_float),
        ("min_apol_neigh", C.c_int32), ("nb_mcv_iter", C.c_int32), ("min_pock_nb_asph", C.c_int32),
        ("do_asa_and_volume", C.c_int32), ("skip_clustering", C.c_int32), ("xpocket_active", C.c_int32),
        ("min_n_explicit_pocket_atoms", C.c_int32), ("want_tets", C.c_int32), ("mc_seed", C.c_uint64),
        ("lig_interface_dist", C.c_float), ("lig_crit_dist", C.c_float),
        ("distance_measure", C.c_int32), ("lig_interface_method", C.c_int32), ("clustering_method", C.c_int32),
    ]


def default_params(**kw):
    """init_def_fparams (reference src/fparams.c:63-95, headers/fparams.h:33-63)."""
    p = FpkParams(np.float32(3.4), np.float32(6.2), np.float32(2.4), 0.0, np.float32(0.7),
                  3, 300, 15, 1, 0, 0, 4, 0, 20260101, np.float32(4.0), np.float32(3.0), 0, 1)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class FpkStructure(C.Structure):
    _fields_ = [
        ("natoms", C.c_int32),
        ("x", _fp), ("y", _fp), ("z", _fp), ("radius", _fp), ("electroneg", _fp), ("bfactor", _fp),
        ("atom_id", _ip), ("res_id", _ip), ("aa_idx", _i8p), ("flags", _bp),
        ("avg_bfactor", C.c_float), ("min_bfactor", C.c_float), ("max_bfactor", C.c_float),
        ("natoms_w", C.c_int32),
        ("wx", _fp), ("wy", _fp), ("wz", _fp), ("wradius", _fp), ("welectroneg", _fp), ("wflags", _bp),
        ("same_pdb", C.c_int32),
        ("n_xlig", C.c_int32), ("xlig", _fp),
        ("n_lig", C.c_int32), ("lig_atoms", _ip), ("lig_mol", _ip), ("atom_name_key", _ip),
    ]


class FpkDesc(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "score", "drug_score", "hydrophobicity_score", "volume_score", "volume", "convex_hull_volume",
        "prop_polar_atm", "mean_asph_ray", "masph_sacc", "apolar_asphere_prop", "mean_loc_hyd_dens",
        "as_density", "as_max_dst", "flex", "nas_norm", "polarity_score_norm", "mean_loc_hyd_dens_norm",
        "prop_asapol_norm", "as_density_norm", "as_max_dst_norm", "surf_vdw14", "surf_vdw22",
        "surf_pol_vdw14", "surf_pol_vdw22", "surf_apol_vdw14", "surf_apol_vdw22", "as_max_r")] + [
        ("bary", C.c_float * 3)] + [(n, C.c_int32) for n in (
            "nb_asph", "polarity_score", "charge_score", "n_abpa", "nAlphaApol", "nAlphaPol", "n_atoms")] + [
        ("aa_compo", C.c_int32 * 20), ("first_sphere", C.c_int32), ("final_rank", C.c_int32)]


DESC_DTYPE = np.dtype(
    [(n, np.float32) for n, t in FpkDesc._fields_[:27]] + [("bary", np.float32, 3)] +
    [(n, np.int32) for n, t in FpkDesc._fields_[28:35]] + [("aa_compo", np.int32, 20),
                                                            ("first_sphere", np.int32), ("final_rank", np.int32)])
assert DESC_DTYPE.itemsize == C.sizeof(FpkDesc), (DESC_DTYPE.itemsize, C.sizeof(FpkDesc))


class FpkResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_sites", C.c_int32), ("n_tets", C.c_int32), ("tets", _ip),
        ("n_spheres", C.c_int32), ("sph_xyzr", _fp), ("sph_bary", _fp), ("sph_neigh", _ip), ("sph_type", _bp),
        ("sph_cluster", _ip), ("n_clusters", C.c_int32), ("cl_off", _ip), ("cl_sph", _ip), ("cl_atom_off", _ip),
        ("cl_atoms", _ip), ("desc", C.POINTER(FpkDesc)), ("n_pockets", C.c_int32), ("rank_to_cluster", _ip),
        ("atom_fx", _fp), ("_owner", C.c_void_p),
        ("n_xspheres", C.c_int32), ("xspheres", _ip), ("xdesc", C.POINTER(FpkDesc)), ("n_iface", C.c_int32), ("iface_atoms", _ip),
        ("pk_ovlp", _fp), ("pk_dst", _fp), ("pk_c4", _fp), ("pk_c5", _fp), ("lig_volume", C.c_float),
    ]


class FpkZone(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_spheres_frame", C.c_int32), ("n_pockets_frame", C.c_int32), ("n_entries", C.c_int32),
        ("spheres", _ip), ("sph_xyzr", _fp), ("sph_type", _bp), ("sph_pocket", _ip), ("n_atoms", C.c_int32), ("atoms", _ip), ("desc", FpkDesc),
        ("atom_fx", _fp), ("_owner", C.c_void_p),
    ]


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,))
    return a.view(dtype).reshape(shape).copy()


def result_to_dict(r, natoms):
    """Copy a filled fpk_result into numpy arrays (the C arrays stay owned by the library)."""
    ns, nc, npk = r.n_spheres, r.n_clusters, r.n_pockets
    d = {
        "status": r.status, "n_sites": r.n_sites, "n_tets": r.n_tets, "n_spheres": ns, "n_clusters": nc,
        "n_pockets": npk,
        "tets": _arr(r.tets, (r.n_tets, 4), np.int32) if r.tets else None,
        "sph_xyzr": _arr(r.sph_xyzr, (ns, 4), np.float32), "sph_bary": _arr(r.sph_bary, (ns, 3), np.float32),
        "sph_neigh": _arr(r.sph_neigh, (ns, 4), np.int32), "sph_type": _arr(r.sph_type, (ns,), np.uint8),
        "sph_cluster": _arr(r.sph_cluster, (ns,), np.int32),
        "cl_off": _arr(r.cl_off, (nc + 1,), np.int32) if nc else np.zeros(1, np.int32),
        "cl_sph": _arr(r.cl_sph, (ns,), np.int32) if nc else np.zeros(0, np.int32),
        "cl_atom_off": _arr(r.cl_atom_off, (nc + 1,), np.int32) if nc else np.zeros(1, np.int32),
        "desc": _arr(r.desc, (nc,), DESC_DTYPE), "rank_to_cluster": _arr(r.rank_to_cluster, (npk,), np.int32),
        "atom_fx": _arr(r.atom_fx, (natoms, 4), np.float32),
    }
    na = int(d["cl_atom_off"][-1]) if nc else 0
    d["cl_atoms"] = _arr(r.cl_atoms, (na,), np.int32)
    if r.xspheres or r.pk_dst:                       # dpocket ligand queries
        d["xspheres"] = _arr(r.xspheres, (r.n_xspheres,), np.int32)
        d["xdesc"] = _arr(r.xdesc, (1,), DESC_DTYPE) if r.xdesc else None
        d["iface_atoms"] = _arr(r.iface_atoms, (r.n_iface,), np.int32)
        for k in ("pk_ovlp", "pk_dst", "pk_c4", "pk_c5"):
            d[k] = _arr(getattr(r, k), (npk,), np.float32)
        d["lig_volume"] = float(r.lig_volume)
    return d


class Structure:
    """Owns the numpy arrays behind one fpk_structure (keeps them alive while C reads them)."""

    def __init__(self, x, y, z, radius, electroneg, bfactor, atom_id, res_id, aa_idx, flags, bfstat,
                 w=None, same_pdb=0, xlig=None):
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        self.x, self.y, self.z = f32(x), f32(y), f32(z)
        self.radius, self.electroneg, self.bfactor = f32(radius), f32(electroneg), f32(bfactor)
        self.atom_id = np.ascontiguousarray(atom_id, dtype=np.int32)
        self.res_id = np.ascontiguousarray(res_id, dtype=np.int32)
        self.aa_idx = np.ascontiguousarray(aa_idx, dtype=np.int8)
        self.flags = np.ascontiguousarray(flags, dtype=np.uint8)
        self.bfstat = tuple(np.float32(v) for v in bfstat)
        self.natoms = len(self.x)
        self.same_pdb = int(same_pdb)
        if w is not None:
            self.wx, self.wy, self.wz, self.wradius, self.welectroneg = (f32(a) for a in w[:5])
            self.wflags = np.ascontiguousarray(w[5], dtype=np.uint8)
            self.natoms_w = len(self.wx)
        else:
            self.natoms_w = 0
        self.xlig = f32(xlig).reshape(-1, 3) if xlig is not None and len(xlig) else None
        self.lig_atoms = self.lig_mol = self.atom_name_key = None

    def set_ligand(self, lig_atoms, lig_mol=None, atom_name_key=None):
        """dpocket: the known ligand as indices into the pdb_w_lig atoms (into the pdb atoms when there is no separate copy)."""
        self.lig_atoms = np.ascontiguousarray(lig_atoms, dtype=np.int32)
        self.lig_mol = np.ascontiguousarray(lig_mol if lig_mol is not None else np.zeros(len(self.lig_atoms)), dtype=np.int32)
        self.atom_name_key = None if atom_name_key is None else np.ascontiguousarray(atom_name_key, dtype=np.int32)
        return self

    @property
    def n_sites(self):
        return int(np.count_nonzero((self.flags & FPK_ATOM_H) == 0))

    def c_struct(self):
        p = lambda a, t: a.ctypes.data_as(t)
        s = FpkStructure()
        s.natoms = self.natoms
        s.x, s.y, s.z = p(self.x, _fp), p(self.y, _fp), p(self.z, _fp)
        s.radius, s.electroneg, s.bfactor = p(self.radius, _fp), p(self.electroneg, _fp), p(self.bfactor, _fp)
        s.atom_id, s.res_id = p(self.atom_id, _ip), p(self.res_id, _ip)
        s.aa_idx, s.flags = p(self.aa_idx, _i8p), p(self.flags, _bp)
        s.avg_bfactor, s.min_bfactor, s.max_bfactor = self.bfstat
        s.natoms_w = self.natoms_w
        if self.natoms_w:
            s.wx, s.wy, s.wz = p(self.wx, _fp), p(self.wy, _fp), p(self.wz, _fp)
            s.wradius, s.welectroneg, s.wflags = p(self.wradius, _fp), p(self.welectroneg, _fp), p(self.wflags, _bp)
        s.same_pdb = self.same_pdb
        if self.xlig is not None:
            s.n_xlig = len(self.xlig)
            s.xlig = p(self.xlig, _fp)
        if getattr(self, "lig_atoms", None) is not None and len(self.lig_atoms):
            s.n_lig = len(self.lig_atoms)
            s.lig_atoms, s.lig_mol = p(self.lig_atoms, _ip), p(self.lig_mol, _ip)
            if self.atom_name_key is not None:
                s.atom_name_key = p(self.atom_name_key, _ip)
        return s


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpocket_cuda.so")
_lib = None


def declare(lib):
    """Attach argtypes/restypes for every symbol include/fpocket_cuda.h declares."""
    vp = C.c_void_p
    sig = {
        "fpk_init": (C.c_int, [_ip, C.c_int]),
        "fpk_device_count": (C.c_int, []),
        "fpk_shutdown": (None, []),
        "fpk_last_error": (C.c_char_p, []),
        "fpk_version": (C.c_char_p, []),
        "fpk_default_params": (None, [C.POINTER(FpkParams)]),
        "fpk_detect": (C.c_int, [C.POINTER(FpkStructure), C.POINTER(FpkParams), C.POINTER(FpkResult)]),
        "fpk_detect_batch": (C.c_int, [C.POINTER(FpkStructure), C.c_int, C.POINTER(FpkParams), C.POINTER(FpkResult)]),
        "fpk_result_free": (None, [C.POINTER(FpkResult)]),
        "fpk_batch_create": (vp, [C.POINTER(FpkStructure), C.c_int, C.POINTER(FpkParams)]),
        "fpk_batch_run": (C.c_int, [vp]),
        "fpk_batch_fetch": (C.c_int, [vp, C.POINTER(FpkResult)]),
        "fpk_batch_last_kernel_ms": (C.c_float, [vp, C.c_int]),
        "fpk_batch_counter": (C.c_longlong, [vp, C.c_int]),
        "fpk_batch_destroy": (None, [vp]),
        "fpk_md_begin": (vp, [C.POINTER(FpkStructure), C.POINTER(FpkParams), C.c_int]),
        "fpk_md_grid_from_frame": (C.c_int, [vp, _fp, _fp, _ip]),
        "fpk_md_set_grid": (C.c_int, [vp, _fp, _ip]),
        "fpk_md_push_frames": (C.c_int, [vp, _fp, C.c_int]),
        "fpk_md_device_grids": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "fpk_md_fetch_grids": (C.c_int, [vp, _fp, _fp]),
        "fpk_md_counter": (C.c_longlong, [vp, C.c_int]),
        "fpk_md_characterize": (C.c_int, [vp, _fp, C.c_int, _fp, C.c_int, C.POINTER(FpkZone)]),
        "fpk_zone_free": (None, [C.POINTER(FpkZone)]),
        "fpk_md_end": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    return sorted(sig)


def load_library(path=None):
    """dlopen libfpocket_cuda.so. No fallback: a missing library is an error."""
    global _lib
    if _lib is None or path is not None:
        p = path or os.environ.get("FPK_LIB") or LIB_PATH        # FPK_LIB: A/B experiments with kernel variants (scripts/)
        if not os.path.exists(p):
            raise RuntimeError(
                "libfpocket_cuda.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % p)
        lib = C.CDLL(p)
        declare(lib)
        if path is not None:
            return lib
        _lib = lib
    return _lib

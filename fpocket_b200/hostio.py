"""ctypes binding of libfpocket_hostio (include/fpocket_hostio.h): the native PDB reader, the writers of the reference's
output directory and the -F pipeline. The library has no CUDA dependency; the detection call is handed in
(`capi.load_library().fpk_detect_batch` in the product)."""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpocket_hostio.so")

FPKIO_OK, FPKIO_ERR_OPEN, FPKIO_ERR_NO_ATOMS, FPKIO_ERR_UNSUPPORTED, FPKIO_ERR_WRITE, FPKIO_ERR_BAD_ARG, FPKIO_ERR_NO_LIGAND = 0, -1, -2, -3, -4, -5, -6


class ReadOpts(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("chains", (C.c_char * 21) * 20), ("chains_are_kept", C.c_int32), ("ligand", C.c_char * 8),
                ("has_xlig", C.c_int32), ("xlig_resnumber", C.c_int32), ("xlig_resname", C.c_char * 8), ("xlig_chain", C.c_char),
                ("n_xpocket", C.c_int32), ("xp_res", C.c_uint16 * 64), ("xp_ins", C.c_char * 64), ("xp_chain", C.c_char * 64),
                ("n_lig_chains", C.c_int32), ("lig_chains", (C.c_char * 21) * 20), ("write_mode", C.c_char), ("model_number", C.c_int32)]

    @classmethod
    def make(cls, drop=None, keep=None, ligand=None, custom_ligand=None, custom_pocket=None, model=0, lig_chains=None):
        o = cls()
        o.model_number = model
        for i, nm in enumerate(lig_chains or []):
            o.lig_chains[i].value = nm.encode()
        o.n_lig_chains = len(lig_chains or [])
        if custom_ligand:
            lib().fpkio_parse_custom_ligand(custom_ligand.encode(), C.byref(o))
        if custom_pocket:
            lib().fpkio_parse_custom_pocket(custom_pocket.encode(), C.byref(o))
        if ligand:
            o.ligand = ligand.encode()
        names = keep if keep is not None else (drop or [])
        o.n_chains = len(names)
        o.chains_are_kept = 1 if keep is not None else 0
        for i, nm in enumerate(names):
            o.chains[i].value = nm.encode()
        return o


class Stats(C.Structure):
    _fields_ = [("n_files", C.c_int32), ("n_written", C.c_int32), ("n_read_failed", C.c_int32), ("n_no_pockets", C.c_int32),
                ("n_detect_failed", C.c_int32), ("n_pockets", C.c_longlong), ("n_atoms", C.c_longlong), ("bytes_in", C.c_longlong),
                ("bytes_out", C.c_longlong), ("wall_s", C.c_double), ("read_cpu_s", C.c_double), ("write_cpu_s", C.c_double),
                ("detect_s", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


DETECT_FN = C.CFUNCTYPE(C.c_int, C.POINTER(capi.FpkStructure), C.c_int, C.POINTER(capi.FpkParams), C.POINTER(capi.FpkResult))
FREE_FN = C.CFUNCTYPE(None, C.POINTER(capi.FpkResult))

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libfpocket_hostio.so is not built: run python __graft_entry__.py")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.fpkio_read_pdb.argtypes = [C.c_char_p, C.POINTER(ReadOpts), C.POINTER(vp)]
        L.fpkio_read_pdb.restype = C.c_int
        L.fpkio_read_pdb_mem.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(ReadOpts), C.POINTER(vp)]
        L.fpkio_read_pdb_mem.restype = C.c_int
        L.fpkio_file_free.argtypes = [vp]
        L.fpkio_file_free.restype = None
        L.fpkio_structure.argtypes = [vp]
        L.fpkio_structure.restype = C.POINTER(capi.FpkStructure)
        L.fpkio_natoms.argtypes = [vp, C.c_int]
        L.fpkio_natoms.restype = C.c_int
        L.fpkio_nhetatm.argtypes = [vp, C.c_int]
        L.fpkio_nhetatm.restype = C.c_int
        L.fpkio_atom_text.argtypes = [vp, C.c_int, C.c_int] + [C.c_char_p] * 5 + [C.POINTER(C.c_int32), C.c_char_p]
        L.fpkio_atom_text.restype = C.c_int
        L.fpkio_write_output.argtypes = [vp, C.c_char_p, C.POINTER(capi.FpkResult), C.POINTER(C.c_longlong)]
        L.fpkio_write_output.restype = C.c_int
        L.fpkio_run_list.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(ReadOpts), C.POINTER(capi.FpkParams), DETECT_FN, FREE_FN,
                                     C.c_int, C.c_int, C.c_int, C.POINTER(Stats)]
        L.fpkio_run_list.restype = C.c_int
        L.fpkio_version.restype = C.c_char_p
        L.fpkio_nligand.argtypes = [vp]
        L.fpkio_nligand.restype = C.c_int
        L.fpkio_ligand_volume.argtypes = [vp, C.c_int, C.c_uint64]
        L.fpkio_ligand_volume.restype = C.c_float
        L.fpkio_dpocket_rows.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(capi.FpkResult), C.c_float, C.POINTER(C.c_void_p)]
        L.fpkio_dpocket_rows.restype = C.c_int
        L.fpkio_free_rows.argtypes = [C.POINTER(C.c_void_p)]
        L.fpkio_free_rows.restype = None
        L.fpkio_run_dpocket.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.POINTER(capi.FpkParams), DETECT_FN, FREE_FN,
                                        C.POINTER(C.c_char_p), C.c_char, C.c_int, C.c_int, C.c_int, C.POINTER(Stats)]
        L.fpkio_write_output_mode.argtypes = [vp, C.c_char_p, C.POINTER(capi.FpkResult), C.c_char, C.POINTER(C.c_longlong)]
        L.fpkio_write_output_mode.restype = C.c_int
        L.fpkio_run_dpocket.restype = C.c_int
        L.fpkio_md_topology.argtypes = [vp, C.POINTER(capi.FpkStructure)]
        L.fpkio_md_topology.restype = C.c_int
        L.fpkio_dcd_open.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fpkio_dcd_open.restype = C.c_int
        L.fpkio_dcd_read.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
        L.fpkio_dcd_read.restype = C.c_int
        L.fpkio_dcd_close.argtypes = [vp]
        L.fpkio_dcd_close.restype = None
        L.fpkio_md_write_outputs.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int,
                                             C.POINTER(capi.FpkParams), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_longlong)]
        L.fpkio_md_write_outputs.restype = C.c_int
        _lib = L
    return _lib


def read_dcd(path, max_frames=1 << 30):
    """All frames of a DCD trajectory as float32 [nframes, natoms, 3] (fpkio_dcd_*)."""
    L = lib()
    h, na, nf = C.c_void_p(), C.c_int(), C.c_int()
    rc = L.fpkio_dcd_open(path.encode(), C.byref(h), C.byref(na), C.byref(nf))
    if rc != FPKIO_OK:
        raise IOError("fpkio_dcd_open(%s) failed: %d" % (path, rc))
    n = min(max_frames, max(nf.value, 1))
    out = np.zeros((n, na.value, 3), np.float32)
    got = L.fpkio_dcd_read(h, out.ctypes.data_as(C.POINTER(C.c_float)), n)
    L.fpkio_dcd_close(h)
    return out[:max(got, 0)], nf.value


def md_write_outputs(pdbfile, dens, freq, origin, dims, n_snapshots, params, paths, use_drug_score=0):
    """normalize / project / write_first_bfactor_density / write_md_grid x 2 (fpkio_md_write_outputs). paths: freq.dx, dens.dx, freq_iso, dens_iso, pdensities."""
    d = np.ascontiguousarray(dens, np.float32).ravel()
    f = np.ascontiguousarray(freq, np.float32).ravel()
    o = np.ascontiguousarray(origin, np.float32)
    dm = np.ascontiguousarray(dims, np.int32)
    pa = (C.c_char_p * 5)(*[p.encode() for p in paths])
    nb = C.c_longlong(0)
    fp = C.POINTER(C.c_float)
    rc = lib().fpkio_md_write_outputs(pdbfile.h, d.ctypes.data_as(fp), f.ctypes.data_as(fp), o.ctypes.data_as(fp), dm.ctypes.data_as(C.POINTER(C.c_int)),
                                      n_snapshots, C.byref(params), use_drug_score, pa, C.byref(nb))
    if rc != FPKIO_OK:
        raise IOError("fpkio_md_write_outputs failed: %d" % rc)
    return nb.value


class PdbFile:
    """One parsed PDB file (fpkio_file). `.structure` is the ctypes fpk_structure the library would be handed."""

    def __init__(self, path=None, data=None, opts=None):
        L = lib()
        h = C.c_void_p()
        o = C.byref(opts) if opts is not None else None
        if data is not None:
            rc = L.fpkio_read_pdb_mem((path or "<memory>").encode(), data, len(data), o, C.byref(h))
        else:
            rc = L.fpkio_read_pdb(path.encode(), o, C.byref(h))
        if rc != FPKIO_OK:
            raise IOError("fpkio_read_pdb(%s) failed: %d" % (path, rc))
        self.h = h
        self.path = path

    @property
    def structure(self):
        return lib().fpkio_structure(self.h).contents

    def natoms(self, with_lig=0):
        return lib().fpkio_natoms(self.h, with_lig)

    def nhetatm(self, with_lig=0):
        return lib().fpkio_nhetatm(self.h, with_lig)

    def arrays(self):
        """The flat arrays as numpy copies, keyed like oracle/ref_io.c's dump."""
        s = self.structure
        n, nw = s.natoms, s.natoms_w

        def arr(p, k, dt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), (k,)).copy() if k else np.zeros(0, dt)
        d = {"n": n, "nw": nw, "same_pdb": s.same_pdb,
             "bfstat": np.array([s.avg_bfactor, s.min_bfactor, s.max_bfactor], np.float32)}
        for nm in ("x", "y", "z", "radius", "electroneg", "bfactor"):
            d[nm] = arr(getattr(s, nm), n, C.c_float)
        d["atom_id"], d["res_id"] = arr(s.atom_id, n, C.c_int32), arr(s.res_id, n, C.c_int32)
        d["aa_idx"], d["flags"] = arr(s.aa_idx, n, C.c_int8), arr(s.flags, n, C.c_uint8)
        for nm in ("wx", "wy", "wz", "wradius", "welectroneg"):
            d[nm] = arr(getattr(s, nm), nw, C.c_float)
        d["wflags"] = arr(s.wflags, nw, C.c_uint8)
        return d

    def atom_text(self, with_lig=0):
        """[n, 40] uint8 rows laid out as ref_io.c dumps them + [n, 3] ints (id, res_id, charge)."""
        L = lib()
        n = self.natoms(with_lig)
        t = np.zeros((n, 40), np.uint8)
        iv = np.zeros((n, 3), np.int32)
        ty, nm, rn, ch, sy = (C.create_string_buffer(8) for _ in range(5))
        ia = C.create_string_buffer(2)
        irc = (C.c_int32 * 3)()
        for k in range(n):
            L.fpkio_atom_text(self.h, with_lig, k, ty, nm, rn, ch, sy, irc, ia)
            t[k, 0:7] = np.frombuffer(ty.raw[:7], np.uint8)
            t[k, 8:13] = np.frombuffer(nm.raw[:5], np.uint8)
            t[k, 16:24] = np.frombuffer(rn.raw[:8], np.uint8)
            t[k, 24:28] = np.frombuffer(ch.raw[:4], np.uint8)
            t[k, 28:31] = np.frombuffer(sy.raw[:3], np.uint8)
            t[k, 32], t[k, 33] = ia.raw[0], ia.raw[1]
            iv[k] = irc[0], irc[1], irc[2]
        return t, iv

    def write_output(self, pdb_path, result):
        """result: a ctypes capi.FpkResult. Returns (pockets written, bytes written)."""
        nb = C.c_longlong(0)
        rc = lib().fpkio_write_output(self.h, pdb_path.encode(), C.byref(result), C.byref(nb))
        if rc < 0:
            raise IOError("fpkio_write_output(%s) failed: %d" % (pdb_path, rc))
        return rc, nb.value

    def nligand(self):
        return lib().fpkio_nligand(self.h)

    def ligand_volume(self, niter, seed):
        return float(lib().fpkio_ligand_volume(self.h, niter, seed))

    def dpocket_rows(self, complex_path, ligand, result, lig_vol):
        """The rows desc_pocket appends to the three dpocket tables for this complex (bytes x 3)."""
        rows = (C.c_void_p * 3)()
        rc = lib().fpkio_dpocket_rows(self.h, complex_path.encode(), ligand.encode(), C.byref(result), lig_vol, rows)
        if rc < 0:
            raise IOError("fpkio_dpocket_rows failed: %d" % rc)
        out = [C.string_at(rows[k]) for k in range(3)]
        lib().fpkio_free_rows(rows)
        return out

    def close(self):
        if self.h:
            lib().fpkio_file_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_list(paths, params, detect=None, free_result=None, group=0, io_threads=0, quiet=1, opts=None):
    """The -F loop: reader pool -> detect -> writer pool. detect / free_result default to libfpocket_cuda's
    fpk_detect_batch / fpk_result_free (needs a B200); the CPU tests pass oracle-backed callbacks."""
    L = lib()
    keep = []
    if detect is None:
        cl = capi.load_library()
        if cl.fpk_init(None, 0) != 0:
            raise RuntimeError("fpk_init failed: %s" % cl.fpk_last_error().decode())
        detect = C.cast(cl.fpk_detect_batch, DETECT_FN)
        free_result = C.cast(cl.fpk_result_free, FREE_FN)
    else:
        detect, free_result = DETECT_FN(detect), FREE_FN(free_result)
    keep += [detect, free_result]
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    st = Stats()
    rc = L.fpkio_run_list(arr, len(paths), C.byref(opts) if opts is not None else None, C.byref(params), detect, free_result,
                          group, io_threads, quiet, C.byref(st))
    if rc != FPKIO_OK:
        raise RuntimeError("fpkio_run_list failed: %d" % rc)
    return st.as_dict()


def run_dpocket(paths, ligands, params, out_paths, detect=None, free_result=None, group=0, io_threads=0, quiet=1, write_mode="d"):
    """dpocket's loop (dpocket.c:83-130): complexes + ligand residue names in, <code>_out/ directories and the three tables out."""
    L = lib()
    if detect is None:
        cl = capi.load_library()
        if cl.fpk_init(None, 0) != 0:
            raise RuntimeError("fpk_init failed: %s" % cl.fpk_last_error().decode())
        detect = C.cast(cl.fpk_detect_batch, DETECT_FN)
        free_result = C.cast(cl.fpk_result_free, FREE_FN)
    else:
        detect, free_result = DETECT_FN(detect), FREE_FN(free_result)
    n = len(paths)
    pa = (C.c_char_p * n)(*[p.encode() for p in paths])
    la = (C.c_char_p * n)(*[l.encode() for l in ligands])
    oa = (C.c_char_p * 3)(*[o.encode() for o in out_paths])
    st = Stats()
    rc = L.fpkio_run_dpocket(pa, la, n, C.byref(params), detect, free_result, oa, write_mode.encode(), group, io_threads, quiet, C.byref(st))
    if rc != FPKIO_OK:
        raise RuntimeError("fpkio_run_dpocket failed: %d" % rc)
    return st.as_dict()

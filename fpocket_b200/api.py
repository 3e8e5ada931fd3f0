"""Python mirror of the reference interface for the accelerated path.

    search_pocket(pdb, params, pdb_w_lig)        reference src/fpocket.c:58      -> detect() / detect_batch()
    mdpocket_detect's per-frame body             reference src/mdpocket.c:142-175 -> MdPocket

Everything here calls the CUDA library through its C ABI (include/fpocket_cuda.h); there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import FpkResult, FpkStructure, default_params, result_to_dict  # noqa: F401


class FpkError(RuntimeError):
    pass


def _lib():
    lib = capi.load_library()
    return lib


def _check(rc, lib, what):
    if rc not in (capi.FPK_OK, capi.FPK_ERR_NO_SPHERES, capi.FPK_ERR_NO_POCKETS, capi.FPK_ERR_DEGENERATE, capi.FPK_ERR_CAPACITY):
        raise FpkError("%s failed (%d): %s" % (what, rc, (lib.fpk_last_error() or b"").decode()))


def init(device=None):
    """fpk_init: `device` = None (the current device, or $FPK_DEVICES), one device index, or a list of devices the library shards over."""
    lib = _lib()
    if device is None:
        rc = lib.fpk_init(None, 0)
    else:
        devs = [int(device)] if np.isscalar(device) else [int(v) for v in device]
        d = (C.c_int32 * len(devs))(*devs)
        rc = lib.fpk_init(d, len(devs))
    if rc != capi.FPK_OK:
        raise FpkError("fpk_init failed (%d): %s" % (rc, (lib.fpk_last_error() or b"").decode()))


def detect(structure, params=None):
    """search_pocket() for one structure. Returns a dict of numpy arrays (see capi.result_to_dict)."""
    return detect_batch([structure], params)[0]


def detect_batch(structures, params=None):
    """The `-F list` loop of fpmain.c:61-76 as one call: structures are processed concurrently on the GPU."""
    lib = _lib()
    p = params if params is not None else default_params()
    n = len(structures)
    arr = (FpkStructure * n)(*[s.c_struct() for s in structures])
    res = (FpkResult * n)()
    rc = lib.fpk_detect_batch(arr, n, C.byref(p), res)
    if rc != capi.FPK_OK:
        raise FpkError("fpk_detect_batch failed (%d): %s" % (rc, (lib.fpk_last_error() or b"").decode()))
    out = [result_to_dict(res[k], structures[k].natoms) for k in range(n)]
    for k in range(n):
        lib.fpk_result_free(C.byref(res[k]))
    return out


def detect_batch_raw(structures, params=None, c_array=None):
    """fpk_detect_batch without copying the results into numpy: returns (ctypes fpk_result array, free()). The arrays behind
    each fpk_result stay owned by the library until free() is called. `c_array` = a prebuilt (FpkStructure * n) array."""
    lib = _lib()
    p = params if params is not None else default_params()
    n = len(structures)
    arr = c_array if c_array is not None else (FpkStructure * n)(*[s.c_struct() for s in structures])
    res = (FpkResult * n)()
    rc = lib.fpk_detect_batch(arr, n, C.byref(p), res)
    if rc != capi.FPK_OK:
        raise FpkError("fpk_detect_batch failed (%d): %s" % (rc, (lib.fpk_last_error() or b"").decode()))

    def free():
        for k in range(n):
            lib.fpk_result_free(C.byref(res[k]))
    return res, free


class Batch:
    """Staged form (inputs stay resident in HBM between runs): create -> run()* -> fetch()."""

    def __init__(self, structures, params=None):
        self.lib = _lib()
        self.p = params if params is not None else default_params()
        self.structures = structures
        n = len(structures)
        self._arr = (FpkStructure * n)(*[s.c_struct() for s in structures])
        self.h = self.lib.fpk_batch_create(self._arr, n, C.byref(self.p))
        if not self.h:
            raise FpkError("fpk_batch_create failed: %s" % (self.lib.fpk_last_error() or b"").decode())

    def run(self):
        rc = self.lib.fpk_batch_run(self.h)
        if rc != capi.FPK_OK:
            raise FpkError("fpk_batch_run failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))

    def fetch(self, as_dicts=True):
        n = len(self.structures)
        res = (FpkResult * n)()
        rc = self.lib.fpk_batch_fetch(self.h, res)
        if rc != capi.FPK_OK:
            raise FpkError("fpk_batch_fetch failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))
        out = [result_to_dict(res[k], self.structures[k].natoms) for k in range(n)] if as_dicts else \
              [(res[k].status, res[k].n_spheres, res[k].n_pockets) for k in range(n)]
        for k in range(n):
            self.lib.fpk_result_free(C.byref(res[k]))
        return out

    def kernel_ms(self, which):
        """0 delaunay+filter, 1 cluster, 2 host round trip, 3 hull, 4 descriptors, 5 finalize+compact, 7 all"""
        return float(self.lib.fpk_batch_last_kernel_ms(self.h, which))

    def counter(self, which):
        return int(self.lib.fpk_batch_counter(self.h, which))

    def close(self):
        if self.h:
            self.lib.fpk_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MdPocket:
    """Per-frame part of mdpocket_detect (mdpocket.c:71-340): pocket detection of every frame and accumulation of the
    density / frequency grids (mdpbase.c:108-262) on the GPU. The topology's coordinates are ignored."""

    def __init__(self, topology, params=None, use_drug_score=False):
        self.lib = _lib()
        self.p = params if params is not None else default_params()
        self.natoms = topology.natoms
        self._topo = topology
        cs = topology.c_struct()
        self.h = self.lib.fpk_md_begin(C.byref(cs), C.byref(self.p), int(bool(use_drug_score)))
        if not self.h:
            raise FpkError("fpk_md_begin failed: %s" % (self.lib.fpk_last_error() or b"").decode())
        self.origin = None
        self.dims = None

    def _xyz(self, xyz):
        a = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, self.natoms, 3)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def grid_from_frame(self, xyz):
        """init_md_grid (mdpbase.c:444-502) from one frame (frame 1 in the reference)."""
        a, ptr = self._xyz(xyz)
        org = np.zeros(3, np.float32)
        dims = np.zeros(3, np.int32)
        rc = self.lib.fpk_md_grid_from_frame(self.h, ptr, org.ctypes.data_as(C.POINTER(C.c_float)), dims.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_grid_from_frame failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))
        return org, dims

    def set_grid(self, origin, dims):
        self.origin = np.ascontiguousarray(origin, np.float32)
        self.dims = np.ascontiguousarray(dims, np.int32)
        rc = self.lib.fpk_md_set_grid(self.h, self.origin.ctypes.data_as(C.POINTER(C.c_float)), self.dims.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_set_grid failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))

    def push_frames(self, xyz):
        a, ptr = self._xyz(xyz)
        rc = self.lib.fpk_md_push_frames(self.h, ptr, len(a))
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_push_frames failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))

    def device_grids(self):
        """(dens_ptr, freq_ptr, ncell): raw device pointers, e.g. to wrap as torch tensors for an NCCL all-reduce."""
        d, f, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        rc = self.lib.fpk_md_device_grids(self.h, C.byref(d), C.byref(f), C.byref(n))
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_device_grids failed")
        return d.value, f.value, n.value

    def torch_grids(self):
        import torch

        class _Wrap:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        d, f, n = self.device_grids()
        return torch.as_tensor(_Wrap(d, n), device="cuda"), torch.as_tensor(_Wrap(f, n), device="cuda")

    def fetch_grids(self):
        n = int(np.prod(self.dims))
        dens = np.zeros(n, np.float32)
        freq = np.zeros(n, np.float32)
        rc = self.lib.fpk_md_fetch_grids(self.h, dens.ctypes.data_as(C.POINTER(C.c_float)), freq.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_fetch_grids failed")
        return dens.reshape(self.dims), freq.reshape(self.dims)

    def characterize(self, xyz, zone):
        """mdpocket's characterize mode (mdpocket.c:428-460) for every frame of `xyz`: the spheres of the surviving pockets within 1.0 of a point of
        `zone` ([nzone, 3]; once per such point), the atoms they contact and their descriptor record. Returns one dict per frame."""
        a, ptr = self._xyz(xyz)
        zone = np.ascontiguousarray(zone, dtype=np.float32).reshape(-1, 3)
        out = (capi.FpkZone * len(a))()
        rc = self.lib.fpk_md_characterize(self.h, ptr, len(a), zone.ctypes.data_as(C.POINTER(C.c_float)), len(zone), out)
        if rc != capi.FPK_OK:
            raise FpkError("fpk_md_characterize failed (%d): %s" % (rc, (self.lib.fpk_last_error() or b"").decode()))
        res = []
        for z in out:
            n = z.n_entries
            d = np.zeros(1, capi.DESC_DTYPE)
            C.memmove(d.ctypes.data, C.addressof(z.desc), capi.DESC_DTYPE.itemsize)
            res.append({"status": z.status, "n_spheres_frame": z.n_spheres_frame, "n_pockets_frame": z.n_pockets_frame,
                        "spheres": capi._arr(z.spheres, (n,), np.int32), "sph_xyzr": capi._arr(z.sph_xyzr, (n, 4), np.float32),
                        "sph_type": capi._arr(z.sph_type, (n,), np.uint8), "sph_pocket": capi._arr(z.sph_pocket, (n,), np.int32), "atoms": capi._arr(z.atoms, (z.n_atoms,), np.int32),
                        "desc": d[0], "atom_fx": capi._arr(z.atom_fx, (self.natoms, 4), np.float32) if z.atom_fx else None})
            self.lib.fpk_zone_free(C.byref(z))
        return res

    def counter(self, which):
        return int(self.lib.fpk_md_counter(self.h, which))

    def close(self):
        if self.h:
            self.lib.fpk_md_end(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# the reference's name for the boundary
search_pocket = detect

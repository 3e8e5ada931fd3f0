"""CPU: the CUDA Delaunay kernel's phase functions (fpocket_b200/csrc/delaunay.cuh), compiled for the host
and stepped sequentially by tests/hostemu/emu_delaunay.cpp, must produce exactly the oracle's (= qhull's)
tetrahedra for any number of emulated threads. This checks the parallel algorithm's logic without a GPU;
the real kernel is checked on the GPU by tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu")
_lib = None


def emu():
    global _lib
    if _lib is None:
        os.makedirs(os.path.join(EMU, "_build"), exist_ok=True)
        so = os.path.join(EMU, "_build", "libemu_delaunay.so")
        srcs = [os.path.join(EMU, "emu_delaunay.cpp"), os.path.join(O.ROOT, "fpocket_b200", "csrc", "delaunay.cuh"),
                os.path.join(O.ROOT, "fpocket_b200", "csrc", "predicates.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.emu_delaunay.restype = C.c_int
        L.emu_delaunay.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int32)),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def emu_delaunay(ki, nthreads):
    ki = np.ascontiguousarray(ki, dtype=np.int64)
    out = C.POINTER(C.c_int32)()
    nt = C.c_int32()
    stats = (C.c_longlong * 8)()
    rc = emu().emu_delaunay(len(ki), ki.ctypes.data_as(C.POINTER(C.c_longlong)), nthreads, 0, C.byref(out), C.byref(nt), stats)
    if rc:
        return rc, None, list(stats)
    t = np.ctypeslib.as_array(out, shape=(nt.value * 4,)).reshape(-1, 4).copy()
    O.lib()._free(out)
    return 0, O.canonicalize(t), list(stats)


@pytest.mark.parametrize("case,nthreads", [("1UYD", 1), ("1UYD", 256), ("5WA6", 32), ("3LKF", 1024), ("md_2yex", 128)])
def test_emulated_kernel_equals_oracle_on_proteins(case, nthreads):
    d = G.load(case)
    st, _ = G.inputs(d)
    _, ki, _ = O.sites(st)
    rc, t, stats = emu_delaunay(ki, nthreads)
    assert rc == 0
    assert np.array_equal(t, O.canonicalize(d["qh.tets"]))


@pytest.mark.parametrize("n,seed", [(5, 0), (12, 1), (100, 2), (3000, 3)])
def test_emulated_kernel_random_clouds(n, seed):
    rng = np.random.default_rng(seed)
    ki = (rng.random((n, 3)) * 40e6).astype(np.int64)
    rc0, t0, _ = O.delaunay(ki)
    rc, t, stats = emu_delaunay(ki, 64)
    assert rc0 == 0 and rc == 0
    assert np.array_equal(t, t0)


def test_emulated_kernel_degenerate_lattice():
    """Cubic lattice: everything cospherical/coplanar. The perturbed triangulation must still be a valid
    triangulation of the cube (volumes add up) and identical to the oracle's (same perturbation)."""
    g = np.stack(np.meshgrid(*[np.arange(5)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.int64) * 1000000
    rc0, t0, vol0 = O.delaunay(g)
    assert rc0 == 0 and abs(vol0 - 64.0) < 1e-9
    for nth in (1, 7, 64):
        rc, t, _ = emu_delaunay(g, nth)
        assert rc == 0
        assert np.array_equal(t, t0)


def test_emulated_kernel_duplicates_and_flat_input():
    rng = np.random.default_rng(5)
    ki = (rng.random((200, 3)) * 20e6).astype(np.int64)
    ki[50] = ki[10]
    ki[51] = ki[10]
    rc0, t0, _ = O.delaunay(ki)
    rc, t, stats = emu_delaunay(ki, 32)
    assert rc0 == 0 and rc == 0
    assert np.array_equal(t, t0) and len(set(t.flatten())) == 198      # the two copies never become vertices
    flat = ki.copy()
    flat[:, 2] = 7
    rc, t, _ = emu_delaunay(flat, 32)
    assert rc == 1          # DERR_DEGENERATE

"""CPU: the warp-per-pocket convex-hull routine of K4b (fpocket_b200/csrc/hull_warp.cuh), compiled for the host as a warp of
one lane (tests/hostemu/emu_hull.cpp), must give exactly the volume of the oracle's exact Delaunay of the same points (the
number reference src/voronoi.c:1235-1291 reads from qconvex "FS") - on random clouds, on the golden pockets and on degenerate
input - or decline (return 1), in which case the kernel hands the pocket to its Delaunay route. The device code is checked on
the GPU by tests/test_gpu_parity.py (convex_hull_volume bit for bit against the oracle)."""
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
        so = os.path.join(EMU, "_build", "libemu_hull.so")
        srcs = [os.path.join(EMU, "emu_hull.cpp"), os.path.join(O.ROOT, "fpocket_b200", "csrc", "hull_warp.cuh"),
                os.path.join(O.ROOT, "fpocket_b200", "csrc", "predicates.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.emu_hull.restype = C.c_int
        L.emu_hull.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_float), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def emu_hull(ki):
    ki = np.ascontiguousarray(ki, dtype=np.int64)
    v, ve = C.c_float(), C.c_double()
    rc = emu().emu_hull(len(ki), ki.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(v), C.byref(ve))
    return rc, v.value, ve.value


@pytest.mark.parametrize("n,seed", [(4, 0), (5, 1), (10, 2), (15, 3), (60, 4), (200, 5), (512, 6), (1024, 7)])
def test_hull_volume_equals_delaunay_volume_on_random_clouds(n, seed):
    rng = np.random.default_rng(seed)
    for rep in range(20):
        ki = (rng.normal(size=(n, 3)) * 6e6).astype(np.int64) + np.array([12345678, -7654321, 555], dtype=np.int64)
        rc0, _, vol0 = O.delaunay(ki)
        rc, v, ve = emu_hull(ki)
        assert rc0 == 0 and rc == 0
        assert ve == vol0 and v == np.float32(vol0)


def test_hull_volume_on_points_of_a_sphere_surface():
    """All points are hull vertices: the face count reaches 2n - 4; above the face capacity the routine declines."""
    rng = np.random.default_rng(11)
    for n, expect in ((300, 0), (514, 0), (600, 1)):
        x = rng.normal(size=(n, 3))
        ki = (x / np.linalg.norm(x, axis=1)[:, None] * 15e6).astype(np.int64)
        rc0, _, vol0 = O.delaunay(ki)
        rc, v, ve = emu_hull(ki)
        assert rc0 == 0 and rc == expect
        if rc == 0:
            assert ve == vol0


def test_hull_volume_on_the_golden_pockets():
    """Every cluster of >= 10 spheres of two reference structures: the centres exactly as the kernel sees them."""
    for case in ("1UYD", "5WA6"):
        d = G.load(case)
        st, p = G.inputs(d)
        o = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        xyz = o["sph_xyzr"][:, :3].astype(np.float64)
        ki = np.rint(xyz * 1e6).astype(np.int64)
        off = o["cl_off"]
        done = 0
        for c in range(len(off) - 1):
            idx = o["cl_sph"][off[c]:off[c + 1]]
            if len(idx) < 10 or len(idx) > 1024:
                continue
            rc0, _, vol0 = O.delaunay(ki[idx])
            rc, v, ve = emu_hull(ki[idx])
            assert rc0 == 0 and rc == 0 and ve == vol0
            done += 1
        assert done > 20


def test_hull_volume_degenerate_inputs():
    # cubic lattice: coplanar faces, collinear edge points, points on faces - strict visibility must still give the cube
    g = np.stack(np.meshgrid(*[np.arange(5)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.int64) * 1000000
    rng = np.random.default_rng(3)
    for rep in range(30):
        ki = g[rng.permutation(len(g))]
        rc, v, ve = emu_hull(ki)
        assert rc in (0, 1)
        if rc == 0:
            assert abs(ve - 64.0) < 1e-12
    assert emu_hull(g)[0] == 0
    # duplicates never become vertices twice
    ki = (rng.random((200, 3)) * 20e6).astype(np.int64)
    ki[50] = ki[10]; ki[51] = ki[10]; ki[199] = ki[0]
    rc0, _, vol0 = O.delaunay(ki)
    rc, v, ve = emu_hull(ki)
    assert rc0 == 0 and rc == 0 and ve == vol0
    # flat input: declined (the Delaunay route reports the reference's failure value)
    flat = ki.copy(); flat[:, 2] = 7
    assert emu_hull(flat)[0] == 1
    # too many points / too wide: pre-checks
    assert emu_hull(np.zeros((1025, 3), np.int64))[0] == 2
    wide = ki.copy(); wide[5, 0] += 1 << 31
    assert emu_hull(wide)[0] == 2


def test_hull_volume_fuzz_on_degenerate_and_structured_sets():
    """Small-integer clouds (coplanar, collinear and repeated points everywhere), points on a coarse lattice, on three planes, along a tube: whenever
    the oracle's Delaunay exists the routine must either decline or return exactly its volume; a flat set is always declined."""
    rng = np.random.default_rng(20260)
    done = 0
    for it in range(900):
        kind = it % 5
        if kind == 0:
            n = int(rng.integers(4, 40))
            span = int(rng.choice([2, 3, 4, 6, 10, 50, 1000]))
            ki = rng.integers(0, span, size=(n, 3)).astype(np.int64) * int(rng.choice([1, 1000, 1000000]))
        else:
            n = int(rng.integers(10, 300))
            if kind == 1:
                ki = (rng.normal(size=(n, 3)) * 3e6).astype(np.int64)
                ki[rng.integers(0, n, size=n // 10)] = ki[0]
            elif kind == 2:
                ki = rng.integers(-8, 8, size=(n, 3)).astype(np.int64) * 500000
            elif kind == 3:
                ki = (rng.random((n, 3)) * 2e7).astype(np.int64)
                ki[:n // 3, 0] = 0
                ki[n // 3:2 * n // 3, 1] = 7000000
            else:
                t = rng.random(n) * 3e7
                ki = np.stack([t, 2e6 * np.sin(t / 4e6) + rng.normal(size=n) * 1.5e6, rng.normal(size=n) * 1.5e6], 1).astype(np.int64)
        rc0, _, vol0 = O.delaunay(ki)
        rc, v, ve = emu_hull(ki)
        if rc0 != 0:
            assert rc == 1
            continue
        assert rc in (0, 1)
        if rc == 0:
            assert ve == vol0 and v == np.float32(vol0), (kind, n)
            done += 1
    assert done > 800

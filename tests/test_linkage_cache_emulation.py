"""CPU: host emulation of K3L's search (fpocket_b200/csrc/k_linkage.cuh) against the oracle's literal restatement of the reference (oracle.c linkage_labels:
full row-major scan for the first strict minimum, clusterlib.c:285). The kernel keeps the reference's matrix layout and moves but finds the pair through cached
row minima: every row remembers its minimum and the FIRST column holding it, the pair is the first row with the smallest minimum, and after a merge a row is
scanned again only if its remembered column was `js` or `is`. Ties decide merges here (points on a lattice, duplicates: many exactly equal distances), so this
checks that the cache reproduces the reference's tie-breaking, for all three linkages and both distance measures. The GPU tests
(tests/test_gpu_parity.py::test_linkage_*) then check the CUDA code against the same oracle on real and synthetic structures."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O


def _metric(a, b, measure):
    if measure == "b":
        r, tw = 0.0, 0.0
        for k in range(3):
            r = r + 1.0 * abs(a[k] - b[k])
            tw += 1.0
        return r / tw
    r = 0.0
    for k in range(3):
        t = a[k] - b[k]
        r += t * t
    return math.sqrt(r)


def k3l_emulate(xyz32, cut, method, measure):
    """the kernel's algorithm, step for step (one thread's view); returns label[i] = lowest sphere of i's cluster"""
    ns = len(xyz32)
    pos = [[float(v) for v in p] for p in xyz32]
    cnt, cid = [1] * ns, list(range(ns))
    npar, lpar, nd = [-1] * ns, [-1] * ns, [0.0] * ns
    D = [[0.0] * i for i in range(ns)]                      # ragged lower triangle, D[i][j] with i > j

    def scan(i):
        mv, mj = 1e300, 0x7fffffff
        for j in range(i):
            if D[i][j] < mv:
                mv, mj = D[i][j], j
        return mv, mj

    rmin, rarg = [0.0] * ns, [0] * ns
    for i in range(1, ns):
        for j in range(i):
            D[i][j] = _metric(pos[i], pos[j], measure)
        rmin[i], rarg[i] = scan(i)
    n, node = ns, 0
    while n > 1:
        best, is_ = min((rmin[i], i) for i in range(1, n))  # first row with the smallest minimum
        js, last = rarg[is_], n - 1
        ci, cj = cnt[is_], cnt[js]
        tot = ci + cj
        get = lambda a, b: D[a][b] if a > b else D[b][a]
        def put(a, b, v):
            if a > b: D[a][b] = v
            else: D[b][a] = v
        if method != "c":
            for j in range(n):
                if j in (is_, js):
                    continue
                a, b = get(is_, j), get(js, j)
                put(js, j, (a if a > b else b) if method == "m" else (a * ci + b * cj) / tot)
        nd[node] = best
        for child in (cid[is_], cid[js]):
            if child >= 0: lpar[child] = node
            else: npar[-child - 1] = node
        cid[js], cid[is_] = -node - 1, cid[last]
        cnt[js], cnt[is_] = tot, cnt[last]
        if method == "c":
            pos[js] = [(pos[js][k] * cj + pos[is_][k] * ci) / tot for k in range(3)]
            pos[is_] = list(pos[last])
        for j in range(last):
            if j != is_:
                put(is_, j, D[last][j])
        m = last
        if method == "c":
            for j in range(m):
                if j != js:
                    put(js, j, _metric(pos[js], pos[j], measure))
        for i in range(1, m):                               # row minima
            rescan = i in (js, is_)
            if not rescan and i > js:
                if rarg[i] in (js, is_):
                    rescan = True
                else:
                    mv, mj = rmin[i], rarg[i]
                    t = D[i][js]
                    if t < mv or (t == mv and js < mj): mv, mj = t, js
                    if i > is_:
                        u = D[i][is_]
                        if u < mv or (u == mv and is_ < mj): mv, mj = u, is_
                    rmin[i], rarg[i] = mv, mj
            if rescan:
                rmin[i], rarg[i] = scan(i)
        n, node = n - 1, node + 1
    own, omin = [-1] * ns, {}
    for i in range(ns):
        a = lpar[i]
        while a >= 0:
            if not nd[a] > cut: own[i] = a
            a = npar[a]
        if own[i] >= 0: omin[own[i]] = min(omin.get(own[i], i), i)
    return [omin[own[i]] if own[i] >= 0 else i for i in range(ns)]


def _oracle(xyz32, cut, method, measure):
    L = O.lib()
    L.orc_cluster_points.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.orc_cluster_points.restype = C.c_int
    lab = np.zeros(len(xyz32), np.int32)
    x = np.ascontiguousarray(xyz32, np.float32)
    nc = L.orc_cluster_points(x.ctypes.data_as(C.POINTER(C.c_float)), len(x), C.c_float(cut), ord(measure), ord(method), lab.ctypes.data_as(C.POINTER(C.c_int)))
    return nc, lab


@pytest.mark.parametrize("method", ["m", "a", "c"])
@pytest.mark.parametrize("measure", ["e", "b"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cached_row_minima_reproduce_the_reference_tie_breaking(method, measure, seed):
    rng = np.random.default_rng(100 * seed + ord(method) + ord(measure))
    # lattice points (spacing 0.8: dozens of exactly equal distances at every step), a few duplicates, a few off-lattice points
    grid = rng.integers(0, 6, size=(70, 3)).astype(np.float32) * np.float32(0.8)
    pts = np.concatenate([grid, grid[:6], rng.normal(scale=2.0, size=(12, 3)).astype(np.float32) + np.float32(2.0)]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    cut = float(np.float32(2.4 if measure == "e" else 1.1))
    nc, lab = _oracle(pts, cut, method, measure)
    rep = k3l_emulate(pts, cut, method, measure)
    # the oracle numbers clusters by their lowest member: label -> that member
    low = {}
    for i, c in enumerate(lab):
        low.setdefault(int(c), i)
    assert [low[int(c)] for c in lab] == rep
    assert 1 < nc < len(pts)


def test_single_linkage_entry_agrees_with_connected_components():
    rng = np.random.default_rng(5)
    pts = (rng.random((200, 3)) * 12).astype(np.float32)
    nc, lab = _oracle(pts, 2.4, "s", "e")
    d = np.sqrt(((pts[:, None].astype(np.float64) - pts[None].astype(np.float64)) ** 2).sum(-1))
    import scipy.sparse.csgraph as cg
    n2, lab2 = cg.connected_components(d <= float(np.float32(2.4)), directed=False)
    assert nc == n2 and len(set(zip(lab.tolist(), lab2.tolist()))) == nc

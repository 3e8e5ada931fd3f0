"""Loader of the synthetic-structure generator (tests/synth/synth.c). Bench / test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

from fpocket_b200.capi import FPK_ATOM_H, Structure  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

AA = "ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL".split()
# reference src/pertable.c:80-92 (electronegativity) and :120-140 (vdW radius) for C N O S
ELEM = ["C", "N", "O", "S"]
ENEG = np.array([2.5, 3.0, 3.5, 2.5], np.float32)
RVDW = np.array([1.7, 1.55, 1.52, 1.8], np.float32)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libsynth.so")
        src = os.path.join(HERE, "synth.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"])
        L = C.CDLL(so)
        L.synth_protein.restype = C.c_int
        L.synth_protein.argtypes = [C.c_uint64, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.POINTER(C.c_int8), C.POINTER(C.c_int8)]
        _lib = L
    return _lib


class Synth:
    __slots__ = ("xyz", "bf", "elem", "res", "aa", "slot", "n")


def generate(seed, natoms, density=0.047):
    cap = natoms + 32
    xyz = np.zeros((cap, 3), np.float32)
    bf = np.zeros(cap, np.float32)
    elem = np.zeros(cap, np.int8)
    res = np.zeros(cap, np.int32)
    aa = np.zeros(cap, np.int8)
    slot = np.zeros(cap, np.int8)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    n = lib().synth_protein(seed, natoms, density, cap, p(xyz, C.c_float), p(bf, C.c_float), p(elem, C.c_int8), p(res, C.c_int32),
                            p(aa, C.c_int8), p(slot, C.c_int8))
    s = Synth()
    s.n = n
    s.xyz, s.bf, s.elem, s.res, s.aa, s.slot = xyz[:n].copy(), bf[:n].copy(), elem[:n].copy(), res[:n].copy(), aa[:n].copy(), slot[:n].copy()
    return s


def to_structure(s, same_pdb=1):
    """Flat input exactly as the reference's rpdb.c would produce it from the PDB text of this structure
    (radius/electronegativity from pertable.c, B-factor statistics rpdb.c:1429-1454 with min/max starting at 0)."""
    n = s.n
    bf = s.bf
    stat = (np.float32(np.sum(bf, dtype=np.float32) / np.float32(n)), np.float32(min(0.0, float(bf.min()))), np.float32(max(0.0, float(bf.max()))))
    # rpdb accumulates the average in a float, atom by atom
    acc = np.float32(0.0)
    for v in bf:
        acc = np.float32(acc + v)
    stat = (np.float32(acc / np.float32(n)), stat[1], stat[2])
    return Structure(s.xyz[:, 0], s.xyz[:, 1], s.xyz[:, 2], RVDW[s.elem], ENEG[s.elem], bf, np.arange(1, n + 1, dtype=np.int32), s.res,
                     s.aa, np.zeros(n, np.uint8), stat, w=None, same_pdb=same_pdb)


def atom_names(s):
    names = ["N", "CA", "C", "O"]
    return [names[int(sl)] if sl < 4 else (ELEM[int(e)] + "%d" % (int(sl) - 3)) for sl, e in zip(s.slot, s.elem)]


def name_keys(s):
    """atom names packed into an int32 each (fpk_structure.atom_name_key), as a C char[4] would be"""
    out = np.zeros(s.n, np.int32)
    for i, nm in enumerate(atom_names(s)):
        out[i] = int.from_bytes(nm.encode()[:4].ljust(4, b"\0"), "little", signed=True)
    return out


def make_ligand(rng, centre, natoms=28, step=1.5, reach=4.5):
    """a blob of carbon atoms around `centre`: a random walk of bonded steps kept within `reach` of it (3-decimal coordinates)"""
    pts = [np.asarray(centre, np.float64)]
    while len(pts) < natoms:
        d = rng.normal(size=3)
        q = pts[rng.integers(len(pts))] + step * d / np.linalg.norm(d)
        if np.linalg.norm(q - pts[0]) <= reach and min(np.linalg.norm(q - p) for p in pts) > 1.2:
            pts.append(q)
    return np.round(np.array(pts), 3).astype(np.float32)


def with_ligand(s, lig_xyz):
    """Structure of the complex as dpocket loads it (dpocket.c:186-203): pdb = protein, pdb_w_lig = protein + HETATM LIG."""
    st = to_structure(s, same_pdb=0)
    nl = len(lig_xyz)
    one = np.ones(nl, np.float32)
    st.wx, st.wy, st.wz = (np.ascontiguousarray(np.concatenate([a, lig_xyz[:, k]])) for k, a in enumerate((st.x, st.y, st.z)))
    st.wradius = np.ascontiguousarray(np.concatenate([st.radius, RVDW[0] * one]))
    st.welectroneg = np.ascontiguousarray(np.concatenate([st.electroneg, ENEG[0] * one]))
    st.wflags = np.zeros(st.natoms + nl, np.uint8)
    st.natoms_w = st.natoms + nl
    st.set_ligand(np.arange(st.natoms, st.natoms + nl), np.zeros(nl, np.int32), name_keys(s))
    return st


def to_pdb(s, path, lig_xyz=None):
    names = ["N", "CA", "C", "O"]
    with open(path, "w") as f:
        for i in range(s.n):
            sl = int(s.slot[i])
            el = ELEM[int(s.elem[i])]
            nm = names[sl] if sl < 4 else (el + "%d" % (sl - 3))
            f.write("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n" %
                    ((i + 1) % 100000, (" " + nm) if len(nm) < 4 else nm, AA[int(s.aa[i])], int(s.res[i]) % 10000,
                     s.xyz[i, 0], s.xyz[i, 1], s.xyz[i, 2], 1.0, s.bf[i], el))
        if lig_xyz is not None:
            for j, q in enumerate(lig_xyz):
                f.write("HETATM%5d %-4s LIG A9001    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n" %
                        ((s.n + j + 1) % 100000, " C%d" % (j + 1) if j < 99 else "C%d" % (j + 1), q[0], q[1], q[2], 1.0, 20.0, "C"))
        f.write("END\n")

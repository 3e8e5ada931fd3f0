"""Host-side preparation mirrored from the reference's host code (what the C shim of INTEGRATION.md does).

Nothing here is on the accelerated path: it turns parsed atom records into the flat arrays of
fpk_structure (include/fpocket_cuda.h).
"""
import numpy as np

from .capi import FPK_ATOM_H, FPK_ATOM_H1, Structure

# reference headers/aa.h:22-41
_AA = {n: i for i, n in enumerate(
    "ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL".split())}


def aa_index(res_name):
    """The switch of set_aa_desc (reference src/descriptors.c:453-507): decides on 1-3 letters only,
    so any residue name starting with one of these letters maps to an amino acid."""
    s = (res_name + "   ")[:3]
    l1, l2, l3 = s[0], s[1], s[2]
    if l1 == "A":
        if l2 == "L":
            return _AA["ALA"]
        if l2 == "R":
            return _AA["ARG"]
        if l2 == "S" and l3 == "P":
            return _AA["ASP"]
        return _AA["ASN"]
    if l1 == "C":
        return _AA["CYS"]
    if l1 == "G":
        if l3 == "U":
            return _AA["GLU"]
        if l3 == "Y":
            return _AA["GLY"]
        return _AA["GLN"]
    if l1 == "H":
        return _AA["HIS"]
    if l1 == "I":
        return _AA["ILE"]
    if l1 == "L":
        return _AA["LYS"] if l2 == "Y" else _AA["LEU"]
    if l1 == "M":
        return _AA["MET"]
    if l1 == "P":
        return _AA["PHE"] if l2 == "H" else _AA["PRO"]
    if l1 == "S":
        return _AA["SER"]
    if l1 == "T":
        if l2 == "H":
            return _AA["THR"]
        if l2 == "R":
            return _AA["TRP"]
        return _AA["TYR"]
    if l1 == "V":
        return _AA["VAL"]
    return -1


def _cstr(row):
    b = bytes(row)
    return b.split(b"\0", 1)[0].decode("latin1")


def structure_from_arrays(f, i, symbol, res_name, bfstat, w=None, same_pdb=0, xlig=None, extra_flags=None):
    """f: [n,6] x y z radius electroneg bfactor; i: [n,>=2] id res_id; symbol/res_name: [n,k] uint8 C strings.
    w: optional (f_w, symbol_w) of the ligand-keeping copy (pdb_w_lig)."""
    n = len(f)
    sym = [_cstr(r) for r in symbol]
    flags = np.array([(FPK_ATOM_H if s == "H" else 0) | (FPK_ATOM_H1 if s[:1] == "H" else 0) for s in sym],
                     dtype=np.uint8)     # voronoi.c:100 / asa.c:89
    if extra_flags is not None:
        flags |= np.asarray(extra_flags, dtype=np.uint8)
    aa = np.array([aa_index(_cstr(r)) for r in res_name], dtype=np.int8) if n else np.zeros(0, np.int8)
    wt = None
    if w is not None:
        fw, symw = w
        wflags = np.array([FPK_ATOM_H if _cstr(r)[:1] == "H" else 0 for r in symw], dtype=np.uint8)  # asa.c:89
        wt = (fw[:, 0], fw[:, 1], fw[:, 2], fw[:, 3], fw[:, 4], wflags)
    return Structure(f[:, 0], f[:, 1], f[:, 2], f[:, 3], f[:, 4], f[:, 5], i[:, 0], i[:, 1], aa, flags, bfstat,
                     w=wt, same_pdb=same_pdb, xlig=xlig)

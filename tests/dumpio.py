"""Reader for the FPKDUMP1 files written by oracle/ref_dump.c (test infrastructure)."""
import struct
import numpy as np

_DT = {b"i": np.int32, b"f": np.float32, b"d": np.float64, b"b": np.uint8}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        assert f.readline() == b"FPKDUMP1\n", "not a FPKDUMP1 file"
        while True:
            nm = f.read(32)
            if len(nm) < 32:
                break
            name = nm.split(b"\0", 1)[0].decode()
            dt = _DT[f.read(1)]
            (ndim,) = struct.unpack("<i", f.read(4))
            dims = struct.unpack("<%di" % ndim, f.read(4 * ndim))
            n = int(np.prod(dims)) if ndim else 1
            out[name] = np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=dt).reshape(dims).copy()
    return out

"""CPU: the C-ABI library loads and exports every symbol include/fpocket_cuda.h declares; the ctypes mirror matches the header's
POD layouts; host-side sharding logic (what bench.py does under torchrun) is exercised with world_size 2 on gloo.
No compute call is made: there is no GPU here and the library has no CPU path (fpk_init must say so)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "fpocket_cuda.h")


def _declared_functions():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fpk_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from fpocket_b200 import capi
    lib = capi.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libfpocket_cuda.so does not export %s" % n
    assert lib.fpk_version().startswith(b"fpocket_b200")


def test_ctypes_mirror_matches_header_layout():
    """sizeof/offsetof as the C compiler sees them == the ctypes structures"""
    from fpocket_b200 import capi
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "fpocket_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(fpk_params), sizeof(fpk_structure), sizeof(fpk_desc), sizeof(fpk_result), sizeof(fpk_zone));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(fpk_params, mc_seed), offsetof(fpk_structure, xlig), offsetof(fpk_desc, final_rank), offsetof(fpk_result, _owner),
         offsetof(fpk_params, lig_interface_method), offsetof(fpk_result, lig_volume), offsetof(fpk_zone, atom_fx));
  return 0; }'''
    tmp = os.path.join(ROOT, "tests", "hostemu", "_build")
    os.makedirs(tmp, exist_ok=True)
    cfile, exe = os.path.join(tmp, "abi_probe.c"), os.path.join(tmp, "abi_probe")
    open(cfile, "w").write(src)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, cfile])
    out = subprocess.check_output([exe]).decode().split()
    sizes, offs = [int(v) for v in out[:5]], [int(v) for v in out[5:]]
    assert sizes == [C.sizeof(capi.FpkParams), C.sizeof(capi.FpkStructure), C.sizeof(capi.FpkDesc), C.sizeof(capi.FpkResult), C.sizeof(capi.FpkZone)]
    assert offs == [capi.FpkParams.mc_seed.offset, capi.FpkStructure.xlig.offset, capi.FpkDesc.final_rank.offset, capi.FpkResult._owner.offset,
                    capi.FpkParams.lig_interface_method.offset, capi.FpkResult.lig_volume.offset, capi.FpkZone.atom_fx.offset]


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fpocket_b200 import api, capi
    lib = capi.load_library()
    assert lib.fpk_init(None, 0) == capi.FPK_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fpk_last_error()
    with pytest.raises(api.FpkError):
        api.init()


def test_product_never_touches_the_oracle():
    """Nothing under fpocket_b200/ may import, include, link or load oracle/ (comments may mention it)."""
    pkg = os.path.join(ROOT, "fpocket_b200")
    bad = re.compile(r"(#\s*include|\bimport\b|\bfrom\b|CDLL|dlopen|LoadLibrary).*oracle", re.I)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                for line in open(os.path.join(dp, f), errors="ignore"):
                    assert not bad.search(line), (f, line)
    ldd = subprocess.check_output(["ldd", os.path.join(pkg, "libfpocket_cuda.so")]).decode()
    assert "oracle" not in ldd


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from fpocket_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
sizes = np.random.default_rng(7).integers(2000, 5001, size=101)
mine = sharding.shard_structures(sizes, rank, world)
# every structure exactly once over the ranks, balanced by heavy-atom count
got = [None, None]
dist.all_gather_object(got, [int(i) for i in mine])
allidx = sorted(got[0] + got[1])
assert allidx == list(range(101)), allidx
loads = [int(sizes[g].sum()) for g in got]
assert abs(loads[0] - loads[1]) <= sizes.max(), loads
# mdpocket: contiguous frame blocks, frame 0 on rank 0; grids summed with one all-reduce (here: gloo stand-in for NCCL)
lo, hi = sharding.shard_frames(1001, rank, world)
spans = [None, None]
dist.all_gather_object(spans, (lo, hi))
assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 1001
grid = torch.full((4, 5, 6), float(hi - lo))
origin = torch.tensor([1.5, 2.5, 3.5]) if rank == 0 else torch.zeros(3)
dist.broadcast(origin, src=0)          # frame 0 on rank 0 fixes the grid geometry (mdpbase.c:444-502): 24 bytes
dist.all_reduce(grid)                  # the one collective of the design
assert torch.all(grid == 1001.0) and origin.tolist() == [1.5, 2.5, 3.5]
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_world_size_2_sharding_on_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o

"""GPU (-m gpu): the library's own multi-GPU paths, through the C ABI. fpk_init is given EVERY device of the box (one on the
driver's test box, where the same code runs as a list of one; two or more under `gpurun --gpus N`, log in profiles/):
  * fpk_detect_batch hands its chunks to all devices - results equal the oracle's whatever device took a chunk;
  * fpk_md_push_frames gives every device a contiguous share of the frames and sums the grids inside the library - the sums equal a
    one-device run bit for bit (integer increments, and 32.32 fixed point with -S);
  * grids summed across ranks IN PLACE (fpk_md_device_grids + all-reduce) are what fpk_md_fetch_grids then returns, also with -S.
"""
import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.fixture()
def all_devices():
    from fpocket_b200 import api, capi
    n = _ndev()
    api.init(list(range(n)))
    assert capi.load_library().fpk_device_count() == n
    yield n
    api.init(0)


def _synth(seed, n):
    import synth
    return synth.generate(seed, n, 0.065)


def test_detect_batch_over_all_devices(all_devices, monkeypatch):
    import synth
    from fpocket_b200 import api
    from test_gpu_parity import assert_same_result
    monkeypatch.setenv("FPK_CHUNK", "3")            # many small chunks: every device's threads get some
    sts = [G.inputs(G.load(c))[0] for c in ("5WA6", "1UYD", "3LKF")] + [synth.to_structure(_synth(20269800 + i, 1500 + 211 * i)) for i in range(9)]
    p = api.default_params()
    sts = sts * 2
    got = api.detect_batch(sts, p)
    want = {}
    for k, (st, r) in enumerate(zip(sts, got)):
        key = k % 12
        if key not in want:
            want[key] = O.search_pocket(st, p, tets=None, rng_mode=O.RNG_PHILOX)
        assert_same_result(r, want[key], "multi-device %d" % k)


@pytest.mark.parametrize("score", [False, True])
def test_md_frames_sharded_over_devices_equal_one_device(score):
    import synth
    from fpocket_b200 import api, capi
    s = _synth(20269850, 1800)
    st = synth.to_structure(s)
    p = api.default_params()
    rng = np.random.default_rng(8)
    nfr = 64 * max(2, _ndev()) + 17
    frames = np.round(s.xyz[None] + rng.normal(scale=0.2, size=(nfr,) + s.xyz.shape).astype(np.float32), 3).astype(np.float32)
    frames[0] = s.xyz
    grids = []
    for devs in ([0], list(range(_ndev()))):
        api.init(devs)
        m = api.MdPocket(st, p, use_drug_score=score)
        org, dims = m.grid_from_frame(frames[:1])
        m.set_grid(org, dims)
        m.push_frames(frames[:5])                # a short push stays on the first device
        m.push_frames(frames[5:])                # shared out over the devices
        grids.append(m.fetch_grids())
        assert m.counter(0) == nfr
        assert m.counter(3) == len(devs), (m.counter(3), devs)
        gd2, gf2 = m.fetch_grids()               # reading twice must not add the other devices' grids twice
        assert np.array_equal(gd2, grids[-1][0]) and np.array_equal(gf2, grids[-1][1])
        m.close()
    api.init(0)
    assert np.array_equal(grids[0][0], grids[1][0]) and np.array_equal(grids[0][1], grids[1][1])
    assert grids[0][0].sum() > 0 and grids[0][1].sum() > 0


@pytest.mark.parametrize("score", [False, True])
def test_grids_summed_in_place_by_the_caller_are_what_fetch_returns(score):
    """with -S the float grids are views of fixed-point sums; fetch used to rewrite them from THIS rank's sums, undoing the
    caller's all-reduce. Stand-in for the all-reduce: the caller doubles both grids in place on the device."""
    import synth
    import torch
    from fpocket_b200 import api
    s = _synth(20269860, 1700)
    st = synth.to_structure(s)
    p = api.default_params()
    frames = np.stack([s.xyz, s.xyz + np.float32(0.05)]).astype(np.float32)
    m = api.MdPocket(st, p, use_drug_score=score)
    org, dims = m.grid_from_frame(frames[:1])
    m.set_grid(org, dims)
    m.push_frames(frames)
    gd, gf = m.fetch_grids()
    td, tf = m.torch_grids()
    td.mul_(2.0)
    tf.mul_(2.0)
    torch.cuda.synchronize()
    gd2, gf2 = m.fetch_grids()
    assert np.array_equal(gd2, 2 * gd) and np.array_equal(gf2, 2 * gf) and gd.sum() > 0
    m.push_frames(frames[:1])                    # more frames: the library's own sums are current again
    gd3, _ = m.fetch_grids()
    if score:
        assert gd3.sum() > gd.sum() and not np.array_equal(gd3, gd2)
    m.close()


def test_mdpocket_fast_on_all_devices_writes_the_same_files(tmp_path):
    """build/mdpocket_fast --devices all (one trajectory on every GPU of the box, BASELINE configs[2] as stated): the five output files are
    byte-identical to a one-device run."""
    import filecmp
    import os
    import subprocess
    import sys
    import synth
    exe = os.path.join(O.ROOT, "build", "mdpocket_fast")
    if not os.path.exists(exe):
        pytest.skip("build/mdpocket_fast not built")
    sys.path.insert(0, os.path.join(O.ROOT, "scripts"))
    import gpu_md_file_to_file as mdf
    s = _synth(20269870, 1500)
    rng = np.random.default_rng(12)
    nfr = 64 * max(2, _ndev()) + 9
    frames = np.round(s.xyz[None] + rng.normal(scale=0.2, size=(nfr,) + s.xyz.shape).astype(np.float32), 3).astype(np.float32)
    frames[0] = s.xyz
    outs = []
    for tag, extra in (("one", ["--devices", "0"]), ("all", ["--devices", "all"])):
        d = tmp_path / tag
        d.mkdir()
        synth.to_pdb(s, str(d / "top.pdb"))
        mdf.write_dcd(str(d / "traj.dcd"), frames)
        r = subprocess.run([exe, "--trajectory_file", "traj.dcd", "--trajectory_format", "dcd", "-f", "top.pdb"] + extra, cwd=str(d), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(d)
    names = sorted(f for f in os.listdir(outs[0]) if f.startswith("mdpout"))
    assert len(names) >= 4
    for f in names:
        assert filecmp.cmp(str(outs[0] / f), str(outs[1] / f), shallow=False), f

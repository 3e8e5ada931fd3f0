#!/usr/bin/env python
"""Benchmark of the fpocket hot path on B200 (contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[1]): a batch of 10,000 synthetic protein-like structures of 2,000-5,000 heavy atoms
(tests/synth, seeds 20260000+k), default fpocket parameters, every stage of search_pocket() per structure.
A "step" is one pass of the whole pipeline over the batch. `value` = structures/s with the batch already resident
in HBM (kernels + the one 32-byte-per-structure count read-back the pipeline needs); `e2e` = the same metric through
fpk_detect_batch() with host buffers (pack, H2D, kernels, D2H, unpack inside the timed region).
With N > 1 (torchrun) every rank processes its own batch (independent structures, no collective): weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--structures B] [--impl reference] [--md-frames F]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SEED0 = 20260000


def _gen_one(k):
    import synth
    rng = np.random.default_rng(SEED0 + k)
    n = int(rng.integers(2000, 5001))
    return synth.generate(SEED0 + k, n, 0.065)


def size_of(k):
    """heavy-atom count of workload structure k (what _gen_one draws first)"""
    return int(np.random.default_rng(SEED0 + k).integers(2000, 5001))


def make_batch(nstruct, workers, first=0, ids=None):
    import synth
    synth.lib()                                   # compile once before forking
    ids = list(range(first, first + nstruct)) if ids is None else [int(i) for i in ids]
    if workers > 1 and nstruct >= 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            raw = pool.map(_gen_one, ids, chunksize=16)
    else:
        raw = [_gen_one(k) for k in ids]
    return raw


_RAW = None


def _pdb_one(job):
    import synth
    i, path = job
    synth.to_pdb(_RAW[i], path)
    return os.path.getsize(path)


def write_pdb_files(raw, n, workers, tag):
    """PDB files of the first n workload structures on tmpfs (before any CUDA state exists: the writers are forked)."""
    global _RAW
    d = tempfile.mkdtemp(prefix="fpkf2f_%s_" % tag, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    jobs = [(i, os.path.join(d, "s%05d.pdb" % i)) for i in range(min(n, len(raw)))]
    _RAW = raw
    if workers > 1 and len(jobs) >= 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            sizes = pool.map(_pdb_one, jobs, chunksize=16)
    else:
        sizes = [_pdb_one(j) for j in jobs]
    _RAW = None
    return d, [j[1] for j in jobs], int(sum(sizes))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_fpocket_rate(raw, nsample, cores, keep_dir=None):
    """Times the UNMODIFIED reference binary (oracle/_ref/fpocket, built from /root/reference by oracle/Makefile) on the first
    `nsample` structures of the workload, one process per structure (the reference is single-threaded and not re-entrant),
    `cores` processes at a time, files on tmpfs. Returns structures/s."""
    import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "fpocket")
    if not os.path.exists(exe):
        return None, "oracle/_ref/fpocket missing"
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="fpkref", dir=base)
    files = []
    for k in range(nsample):
        f = os.path.join(tmp, "s%05d.pdb" % k)
        synth.to_pdb(raw[k], f)
        files.append(f)
    t0 = time.time()
    procs = []
    it = iter(files)
    done = 0
    env = dict(os.environ, TMPDIR=tmp)
    while True:
        while len(procs) < cores:
            f = next(it, None)
            if f is None:
                break
            procs.append(subprocess.Popen([exe, "-f", f], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp, env=env))
        if not procs:
            break
        procs[0].wait()
        procs.pop(0)
        done += 1
    dt = time.time() - t0
    subprocess.call(["rm", "-rf", tmp])
    return nsample / dt, dt


def reference_mdpocket_rate(base, frames, cores):
    """Times the UNMODIFIED reference mdpocket (oracle/_ref/mdpocket --pdb_list) on `frames` ([F, n, 3]) of the workload: the program
    is sequential over frames (one grid, mdpocket.c:235-271), so the frames are cut into `cores` chunks, one process per chunk
    (the fair host-cores figure of SURVEY 8d). Returns (frames/s, seconds) or (None, reason)."""
    import copy
    import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "mdpocket")
    if not os.path.exists(exe):
        return None, "oracle/_ref/mdpocket missing"
    tmp = tempfile.mkdtemp(prefix="fpkmd", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    F = len(frames)
    per = max(1, (F + cores - 1) // cores)
    lists = []
    for c0 in range(0, F, per):
        d = os.path.join(tmp, "c%03d" % (c0 // per))
        os.makedirs(d)
        names = []
        for f in range(c0, min(F, c0 + per)):
            s = copy.copy(base)
            s.xyz = frames[f]
            fn = os.path.join(d, "f%05d.pdb" % f)
            synth.to_pdb(s, fn)
            names.append(fn)
        open(os.path.join(d, "list.txt"), "w").write("\n".join(names) + "\n")
        lists.append(d)
    t0 = time.time()
    procs = [subprocess.Popen([exe, "--pdb_list", os.path.join(d, "list.txt")], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in lists]
    for pr in procs:
        pr.wait()
    dt = time.time() - t0
    subprocess.call(["rm", "-rf", tmp])
    return F / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--structures", type=int, default=10000, help="structures per GPU (BASELINE configs[1]: 10,000)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--md-frames", type=int, default=2048, help="also measure mdpocket frames/s on a synthetic trajectory of this many frames per GPU (0: skip)")
    ap.add_argument("--md-file-frames", type=int, default=8192, help="also run build/mdpocket_cuda file to file on a DCD trajectory of this many frames (0: skip)")
    ap.add_argument("--dpocket", type=int, default=1000, help="also measure dpocket (BASELINE configs[4]): this many ligand-bound synthetic structures per GPU (0: skip)")
    ap.add_argument("--large", type=int, default=150000, help="also time ONE synthetic complex of this many heavy atoms (BASELINE configs[3]: 150000) on rank 0 (0: skip)")
    ap.add_argument("--strong", type=int, default=10000, help="also measure configs[1] as stated: this many structures IN TOTAL, shared out over the GPUs (strong scaling; 0: skip)")
    ap.add_argument("--md-total-frames", type=int, default=20000, help="also measure configs[2] as stated: a trajectory of this many frames IN TOTAL, frames shared out over the GPUs (0: skip)")
    ap.add_argument("--files", type=int, default=2000, help="also measure the file-to-file -F loop (native PDB reader -> fpk_detect_batch -> native writers) on this many workload structures per GPU (0: skip)")
    ap.add_argument("--cores", type=int, default=0, help="after the workload is generated, pin this process to the first N cores (emulates the per-rank share of a busy multi-GPU box)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    workers = max(1, ncores // max(1, world))

    cfg = {"workload": "batch of %d synthetic protein-like structures of 2k-5k heavy atoms per GPU (BASELINE configs[1]), default fpocket "
                       "parameters, full search_pocket() per structure" % args.structures,
           "structures_per_gpu": args.structures, "generator": "tests/synth/synth.c seeds 20260000+k, density 0.065",
           "l2": "inputs + per-block scratch (> 1 GB) exceed the 126 MB L2; nothing is reused between steps",
           "parallelism": "structures sharded per GPU, no collective"}

    # ------------------------------------------------------------------ reference arm -----------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        nsample = args.ref_sample or max(ncores * 3, 48)
        raw = make_batch(nsample, ncores)
        for _ in range(max(0, min(args.warmup, 1))):
            reference_fpocket_rate(raw, min(nsample, ncores), ncores)
        rates = []
        for _ in range(max(1, args.steps)):
            r, dt = reference_fpocket_rate(raw, nsample, ncores)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": dt}))
                return
            rates.append((r, dt))
        v = float(np.mean([r for r, _ in rates]))
        ms = float(np.mean([dt for _, dt in rates])) * 1e3
        sample = "first %d structures of the workload per step, one `fpocket -f` process each, %d at a time, tmpfs" % (nsample, ncores)
        print(json.dumps({"impl": "reference", "metric": "structures/sec", "value": v, "unit": "structures/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32/f64", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "structures/s", "cores": ncores, "kind": "reference", "sample": sample},
                          "e2e": {"value": v, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm ----------------------------------------------------
    # the synthetic workload first: worker processes are forked before any CUDA / NCCL state exists in this process
    import synth
    t0 = time.time()
    if world > 1:       # weak scaling: the global workload is world x structures; ranks take balanced shares without communication
        from fpocket_b200 import sharding
        sizes = [size_of(k) for k in range(args.structures * world)]
        mine = sharding.shard_structures(sizes, rank, world)[: args.structures]
        if len(mine) < args.structures:
            mine = list(mine) + [int(mine[-1])] * (args.structures - len(mine))
        raw = make_batch(args.structures, workers, ids=mine)
    else:
        raw = make_batch(args.structures, workers)
    f2f_dir = None
    if args.files > 0:
        f2f_dir, f2f_paths, f2f_bytes = write_pdb_files(raw, args.files, workers, "r%d" % rank)
    if args.cores > 0:
        os.sched_setaffinity(0, set(range(args.cores)))
    import torch
    from fpocket_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its version banner to stdout when the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    api.init(local)
    structs = [synth.to_structure(s) for s in raw]
    t_gen = time.time() - t0
    p = api.default_params()
    batch = api.Batch(structs, p)                      # pack + H2D: inputs now resident in HBM

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        batch.run()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    kms = np.zeros(8)
    t1 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        batch.run()                                    # returns after the device finished; per-stage CUDA-event times inside
        for i in range(8):
            kms[i] += batch.kernel_ms(i)
        dev_ms += batch.kernel_ms(7)
    barrier()
    wall = time.perf_counter() - t1
    clocks = sampler.finish() if sampler else None
    launches = batch.counter(5) * args.steps
    tot = {k: batch.counter(i) for i, k in enumerate(["sites", "tets", "spheres", "clusters", "pockets"])}
    subrounds, retries = batch.counter(8), batch.counter(9)
    # max over ranks of the timed region (device-timed whole steps; wall kept for the record)
    tt = torch.tensor([dev_ms / 1e3, wall], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_wall = float(tt[0]), float(tt[1])
    value = args.structures * world * args.steps / t_dev        # device-timed (CUDA events around every step), max over ranks
    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    stage_names = ["k_delaunay_filter", "k_cluster", "host_count_roundtrip+alloc", "k_hull_volume", "k_descriptors", "k_finalize+k_atom_fx+k_compact"]
    # (k_hull_volume = k_hull_warp + k_hull_volume; k_descriptors = k_descriptors_small + k_descriptors: one CUDA-event interval each)
    per_stage = {stage_names[i]: kms[i] / args.steps for i in range(6)}
    dom = max(("k_delaunay_filter", "k_hull_volume", "k_descriptors", "k_cluster"), key=lambda k: per_stage[k])
    # algorithmic bytes: B = 32 N + 32 T + 108 S + 160 P0 per structure; split per kernel in DESIGN.md
    N, T, S, P0 = tot["sites"], tot["tets"], tot["spheres"], tot["clusters"]
    alg = {"k_delaunay_filter": 32 * N + 32 * T + 36 * S, "k_cluster": 36 * S + 8 * S, "k_descriptors": 36 * S + 160 * P0 + 32 * N,
           "k_hull_volume": 16 * S + 4 * P0}
    ach = alg[dom] / (per_stage[dom] * 1e-3) / 1e9
    whole = (32 * N + 32 * T + 108 * S + 160 * P0)
    traffic, traffic_n = None, 0
    try:        # DRAM bytes of the dominant kernel from the committed ncu --set full capture, scaled to this launch by algorithmic bytes
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if tj["kernel"] == dom:
            traffic = tj["dram_bytes"] * alg[dom] / tj["algorithmic_bytes"]
            traffic_n = tj["structures"]
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": ("profiles/k1_traffic.json (ncu dram__bytes_read+write of a %d-structure launch, scaled by algorithmic bytes)" % traffic_n) if traffic else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": per_stage[dom],
            "whole_path": {"algorithmic_bytes_per_step": whole, "achieved_gbs": whole / (t_dev / args.steps) / 1e9,
                           "frac": whole / (t_dev / args.steps) / 1e9 / peak}}
    # compute-side figure for K4 (the explanatory second number): executed thread instructions per second against the FP32 lane
    # rate of the chip; the instruction count per structure comes from the committed ncu capture, the time is this run's CUDA-event time
    try:
        kj = json.load(open(os.path.join(ROOT, "profiles", "k4_compute.json")))
        sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
        tinst = kj["warp_inst"] * kj["active_lanes"] / kj["structures"] * args.structures
        peak_lanes = torch.cuda.get_device_properties(local).multi_processor_count * 128 * sm_hz * 1e6
        roof["k_descriptors_compute"] = {"thread_inst_per_launch": tinst, "achieved_Tinst_s": tinst / (per_stage["k_descriptors"] * 1e-3) / 1e12,
                                         "peak_fp32_lane_Tinst_s": peak_lanes / 1e12, "frac": tinst / (per_stage["k_descriptors"] * 1e-3) / peak_lanes,
                                         "issue_slots_active_pct_ncu": kj["issue_active_pct"], "source": "profiles/k4_compute.json (" + kj["capture"] + ")"}
    except Exception:
        pass
    out = {"metric": "structures/sec", "value": value, "unit": "structures/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64 predicates/circumcentres + f32 descriptors (the reference's own mix)", "data": "synthetic", "config": cfg,
           "wall_ms_per_step": t_wall / args.steps * 1e3, "stage_ms_per_step": per_stage, "roofline": roof, "gpu_launches": int(launches),
           "totals_per_gpu": tot, "delaunay_subrounds_per_structure": subrounds / max(1, args.structures),
           "delaunay_retries_per_site": retries / max(1, N),
           "delaunay_debug": {"attempts": batch.counter(13), "walk_steps_per_attempt": batch.counter(11) / max(1, batch.counter(13)), "bfs_levels_per_attempt": batch.counter(12) / max(1, batch.counter(13)), "dead_hops_per_attempt": batch.counter(14) / max(1, batch.counter(13)), "big": batch.counter(10)}, "clocks": clocks, "gen_s": t_gen}

    # ---- e2e: the public call with host buffers ---------------------------------------------------------------------------------
    if not args.no_e2e:
        import ctypes
        from fpocket_b200 import capi
        ne2e = args.structures
        carr = (capi.FpkStructure * ne2e)(*[s.c_struct() for s in structs[:ne2e]])       # host pointers only: no data is copied here
        _, free = api.detect_batch_raw(structs[:ne2e], p, c_array=carr)                    # warm the pooled worker contexts (device arenas, pinned mirrors)
        free()
        e2e_steps = max(1, min(args.steps, 2))
        lib = capi.load_library()
        byt0 = (lib.fpk_batch_counter(None, 6), lib.fpk_batch_counter(None, 7))
        barrier()
        te = time.perf_counter()
        ok = 0
        for _ in range(e2e_steps):
            res, free = api.detect_batch_raw(structs[:ne2e], p, c_array=carr)              # pack, H2D, kernels, D2H, unpack: all inside
            ok = sum(1 for k in range(ne2e) if res[k].status == 0)
            if ok != ne2e:
                import collections
                print("e2e statuses:", collections.Counter(res[k].status for k in range(ne2e)), [k for k in range(ne2e) if res[k].status != 0][:8], file=sys.stderr)
            free()
        barrier()
        te = (time.perf_counter() - te) / e2e_steps
        h2d = (lib.fpk_batch_counter(None, 6) - byt0[0]) // e2e_steps          # counted by the library from the copies it issues
        d2h = (lib.fpk_batch_counter(None, 7) - byt0[1]) // e2e_steps
        ee = torch.tensor([te], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ee, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": ne2e * world / float(ee[0]), "unit": "structures/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "structures_per_call": ne2e, "ok": int(ok), "steps": e2e_steps,
                      "note": "fpk_detect_batch() on host arrays: chunks of ~1,200 structures through pack -> pinned H2D -> kernels -> pinned D2H -> unpack on 4 or 8 host "
                              "threads / CUDA streams; wall clock around the call, results freed inside the timed region"}

    # ---- configs[1] as stated (strong scaling): `--strong` structures IN TOTAL, rank r takes its share of the same workload ---------
    if args.strong > 0:
        share = max(1, args.strong // world)
        sb = batch if share == args.structures else api.Batch(structs[:share], p)
        sb.run()
        barrier()
        sdev = 0.0
        ssteps = max(1, min(args.steps, 5))
        for _ in range(ssteps):
            sb.run()
            sdev += sb.kernel_ms(7)
        barrier()
        ss = torch.tensor([sdev / 1e3], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ss, op=dist.ReduceOp.MAX)
        out["config2_strong"] = {"metric": "structures/sec", "value": share * world * ssteps / float(ss[0]), "structures_total": share * world,
                                 "structures_per_gpu": share, "scaling": "strong", "steps": ssteps, "ms_per_step": float(ss[0]) / ssteps * 1e3,
                                 "includes": "device-timed whole steps (inputs resident), max over ranks; no collective"}
        if sb is not batch:
            sb.close()

    # ---- file to file: the -F loop on PDB files, native reader pool -> fpk_detect_batch -> native writer pool ---------
    if f2f_dir is not None:
        from fpocket_b200 import hostio
        nf = len(f2f_paths)
        io_threads = max(2, workers)
        hostio.run_list(f2f_paths[: min(nf, 256)], p, group=128, io_threads=io_threads, quiet=1)       # warm-up (page cache, arenas, pool)
        subprocess.call("rm -rf %s/*_out" % f2f_dir, shell=True)
        barrier()
        tf = time.perf_counter()
        st = hostio.run_list(f2f_paths, p, group=512, io_threads=io_threads, quiet=1)
        barrier()
        tf = time.perf_counter() - tf
        ff = torch.tensor([tf], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ff, op=dist.ReduceOp.MAX)
        out["file_to_file"] = {"metric": "structures/sec", "value": nf * world / float(ff[0]), "files_per_gpu": nf, "written": st["n_written"],
                               "pockets": st["n_pockets"], "io_threads": io_threads, "MB_in": st["bytes_in"] / 1e6, "MB_out": st["bytes_out"] / 1e6,
                               "read_cpu_s": st["read_cpu_s"], "write_cpu_s": st["write_cpu_s"], "detect_s": st["detect_s"], "wall_s": tf,
                               "includes": "PDB files on tmpfs -> fpkio_run_list (libfpocket_hostio: one-pass reader of both atom lists, groups of 512 "
                                           "through fpk_detect_batch, writers of the reference's <code>_out/ directory: _out.pdb, _info.txt, _pockets.pqr, "
                                           "4 scripts, 2 files per pocket), byte-identical to the reference's writers (tests/test_hostio.py)"}
        if rank == 0 and world == 1 and not args.no_cpu:
            bexe = os.path.join(ROOT, "build", "fpocket_cuda_batch")
            if os.path.exists(bexe):
                nb = min(nf, 128)
                subprocess.call("rm -rf %s/*_out" % f2f_dir, shell=True)
                with open(os.path.join(f2f_dir, "list_ref.txt"), "w") as f:
                    f.write("".join(os.path.basename(q) + "\n" for q in f2f_paths[:nb]))
                tb = time.time()
                rcb = subprocess.call([bexe, "list_ref.txt"], cwd=f2f_dir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                tb = time.time() - tb
                out["file_to_file"]["reference_host_io"] = {"value": nb / tb if rcb == 0 else None, "unit": "structures/s",
                                                            "sample": "first %d files through build/fpocket_cuda_batch: the reference's own reader and writers "
                                                                      "(one thread, not re-entrant) around the same fpk_detect_batch; %.1f s incl. CUDA start-up" % (nb, tb)}
        subprocess.call(["rm", "-rf", f2f_dir])

    # ---- optional: mdpocket frames/s with the grids summed over ranks (NCCL) ------------------------------------------------------
    if args.md_frames > 0:
        s5 = synth.generate(20260301, 5000, 0.065)
        st5 = synth.to_structure(s5)
        rng = np.random.default_rng(20260301 + rank)
        F = args.md_frames
        modes = rng.normal(size=(8, s5.n, 3)).astype(np.float32) * 0.25
        amp = np.cumsum(rng.normal(size=(F, 8)).astype(np.float32) * 0.15, 0)
        amp -= amp.mean(0)
        xyz = (s5.xyz[None] + np.tensordot(np.clip(amp, -1.2, 1.2), modes, 1)).astype(np.float32)
        xyz = np.round(xyz, 3)
        md = api.MdPocket(st5, p)
        org, dims = md.grid_from_frame(s5.xyz[None])             # frame 1 defines the grid (mdpocket.c:151-154); same on every rank
        md.set_grid(org, dims)
        md.push_frames(xyz)                                      # warm-up with the same chunk shapes (arenas, pinned staging), then reset the grids
        md.set_grid(org, dims)
        barrier()
        tm = time.perf_counter()
        md.push_frames(xyz)
        if dist is not None:
            gd, gf = md.torch_grids()
            dist.all_reduce(gd)                                  # the only collective of the whole design
            dist.all_reduce(gf)
        barrier()
        tm = time.perf_counter() - tm
        mm = torch.tensor([tm], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(mm, op=dist.ReduceOp.MAX)
        out["mdpocket"] = {"metric": "frames/sec", "value": F * world / float(mm[0]), "frames_per_gpu": F, "atoms": int(s5.n), "scaling": "weak",
                           "grid": [int(v) for v in dims], "includes": "H2D of every frame, detection, grid scatter, NCCL all-reduce of both grids"}
        # roofline of the frame path (B_frame = B(structure, ASA / volume / hull off) + 8 B x sum over kept spheres of (26 + (2 floor(r/2) + 1)^3)
        # for the grid read-modify-write + 4 B x cells / 8 for the per-frame bit mask), from one fetched frame's counts
        try:
            one = api.detect(synth.to_structure(s5), api.default_params(do_asa_and_volume=0))
            rr = one["sph_xyzr"][:, 3]
            stencil = float(np.sum(26 + (2 * np.floor(rr / 2.0) + 1) ** 3))
            b_frame = 32 * one["n_sites"] + 32 * one["n_tets"] + 108 * one["n_spheres"] + 160 * one["n_clusters"] + 8 * stencil + int(np.prod(dims)) / 2
            ach = b_frame * F / float(mm[0]) / 1e9
            out["mdpocket"]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                           "algorithmic_bytes_per_frame": b_frame, "peak_source": peak_src,
                                           "note": "whole frame path per GPU incl. the H2D of the frame; the kernels are the structure path's (K1 dominates)"}
        except Exception as ex:          # the roofline is an annotation: never lose the line for it
            out["mdpocket"]["roofline"] = {"error": str(ex)}
        # configs[2] as stated (strong scaling): a 20,000-frame trajectory IN TOTAL, contiguous frame blocks per GPU, grids summed once with NCCL
        if args.md_total_frames > 0:
            Fs = max(64, args.md_total_frames // world)
            reps = (Fs + F - 1) // F
            xs = np.concatenate([xyz] * reps)[:Fs] if reps > 1 else xyz[:Fs]
            md.set_grid(org, dims)
            barrier()
            ts = time.perf_counter()
            md.push_frames(xs)
            if dist is not None:
                gd, gf = md.torch_grids()
                dist.all_reduce(gd)
                dist.all_reduce(gf)
            barrier()
            ts = time.perf_counter() - ts
            st_ = torch.tensor([ts], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(st_, op=dist.ReduceOp.MAX)
            out["mdpocket"]["strong"] = {"value": Fs * world / float(st_[0]), "unit": "frames/s", "frames_total": Fs * world, "frames_per_gpu": Fs,
                                         "scaling": "strong", "seconds": float(st_[0]),
                                         "includes": "BASELINE configs[2]: one trajectory of %d frames of the 5,000-atom protein, contiguous blocks per GPU, H2D of every "
                                                     "frame, detection, scatter, one NCCL all-reduce of the two grids" % (Fs * world)}
            del xs
        md.close()
        # file to file: build/mdpocket_fast (fpocket_b200/hostio/mdpocket_fast.cpp): topology PDB and DCD trajectory in through the native readers
        # (double-buffered against the GPU), DX grids / PDB files out through the native writers (byte-identical to the reference's, tests/)
        mexe = os.path.join(ROOT, "build", "mdpocket_fast")
        if rank == 0 and world == 1 and os.path.exists(mexe) and args.md_file_frames > 0:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import gpu_md_file_to_file as mdf
            Ff = args.md_file_frames
            reps = (Ff + F - 1) // F
            traj = np.concatenate([xyz] * reps)[:Ff]
            tmpd = tempfile.mkdtemp(prefix="fpkmdf", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            synth.to_pdb(s5, os.path.join(tmpd, "top.pdb"))
            mdf.write_dcd(os.path.join(tmpd, "traj.dcd"), traj)
            del traj
            tq = time.time()
            rcm = subprocess.call([mexe, "--trajectory_file", "traj.dcd", "--trajectory_format", "dcd", "-f", "top.pdb"], cwd=tmpd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            tq = time.time() - tq
            okf = rcm == 0 and os.path.exists(os.path.join(tmpd, "mdpout_dens_grid.dx"))
            out["mdpocket"]["file_to_file"] = {"value": Ff / tq if okf else None, "unit": "frames/s", "frames": Ff, "process_wall_s": tq,
                                               "includes": "one process (build/mdpocket_fast): CUDA start-up, topology PDB, %d-frame DCD, detection, grid scatter, "
                                                           "normalize / project / DX / PDB writers" % Ff}
            subprocess.call(["rm", "-rf", tmpd])
        if rank == 0 and world == 1 and not args.no_cpu:
            nfr = 2 * ncores
            r, dt = reference_mdpocket_rate(s5, xyz[:nfr], ncores)
            out["mdpocket"]["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": ncores, "kind": "reference",
                                               "sample": "first %d frames, oracle/_ref/mdpocket --pdb_list (the unmodified reference), %d processes of %d frames each, tmpfs; %s s"
                                                         % (nfr, ncores, 2, ("%.1f" % dt) if r else dt)}

    # ---- dpocket (configs[4]): ligand-bound structures; detection + ligand queries + the explicit pocket's record, one call -----------
    if args.dpocket > 0:
        from fpocket_b200 import capi
        nd = min(args.dpocket, len(structs))
        res, free = api.detect_batch_raw(structs[:nd], p)                  # first pass: where is the best pocket? the ligand goes there
        cents = []
        for k in range(nd):
            r = res[k]
            cents.append([r.desc[r.rank_to_cluster[0]].bary[q] for q in range(3)] if r.status == 0 and r.n_pockets > 0 else None)
        free()
        rngl = np.random.default_rng(20260501 + rank)
        ligs = [synth.make_ligand(rngl, c if c is not None else raw[k].xyz.mean(0)) for k, c in enumerate(cents)]
        cplx = [synth.with_ligand(raw[k], ligs[k]) for k in range(nd)]
        carr = (capi.FpkStructure * nd)(*[s.c_struct() for s in cplx])
        _, free = api.detect_batch_raw(cplx, p, c_array=carr)
        free()
        barrier()
        td = time.perf_counter()
        res, free = api.detect_batch_raw(cplx, p, c_array=carr)
        okd = sum(1 for k in range(nd) if res[k].status == 0 and res[k].n_xspheres > 0)
        mean_ovlp = float(np.mean([res[k].pk_ovlp[0] for k in range(nd) if res[k].status == 0 and res[k].n_pockets > 0]))
        free()
        barrier()
        td = time.perf_counter() - td
        dd = torch.tensor([td], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dd, op=dist.ReduceOp.MAX)
        dp_alg = (32 * N + 32 * T + 108 * S + 160 * P0) / max(1, args.structures) * nd          # the structure path's bytes (the ligand arrays add < 1 %)
        out["dpocket"] = {"metric": "structures/sec", "value": nd * world / float(dd[0]), "structures_per_gpu": nd, "with_explicit_pocket": int(okd),
                          "roofline": {"bound": "hbm", "achieved": dp_alg / float(dd[0]) / 1e9, "peak": peak, "unit": "GB/s", "frac": dp_alg / float(dd[0]) / 1e9 / peak,
                                       "traffic": None, "algorithmic_bytes_per_call": dp_alg, "peak_source": peak_src,
                                       "note": "end to end through fpk_detect_batch (host buffers): pack, H2D, K1-K6, D2H, unpack"},
                          "mean_overlap_of_top_pocket_pct": mean_ovlp,
                          "includes": "fpk_detect_batch on host arrays with the known ligand: detection, ligand queries (K6), the explicit pocket's descriptor "
                                      "record; H2D/D2H inside; 28-atom ligand placed in the best pocket of a first pass"}
        # file to file (SURVEY 8 f1 for configs[4]): PDB complexes + ligand names -> fpkio_run_dpocket (native reader in ligand mode, one
        # fpk_detect_batch per group, native writers of every <code>_out/ directory and of the three dpocket tables)
        if args.files > 0:
            from fpocket_b200 import hostio
            ndf = min(nd, args.files)
            tmpd = tempfile.mkdtemp(prefix="fpkdpf_r%d_" % rank, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            dpaths = [os.path.join(tmpd, "c%05d.pdb" % k) for k in range(ndf)]
            for k in range(ndf):
                synth.to_pdb(raw[k], dpaths[k], ligs[k])
            outs = [os.path.join(tmpd, nm) for nm in ("dpout_explicitp.txt", "dpout_fpocketp.txt", "dpout_fpocketnp.txt")]
            io_threads = max(2, workers)
            hostio.run_dpocket(dpaths[: min(ndf, 128)], ["LIG"] * min(ndf, 128), p, outs, group=128, io_threads=io_threads)     # warm-up
            subprocess.call("rm -rf %s/*_out" % tmpd, shell=True)
            barrier()
            tq = time.perf_counter()
            sd = hostio.run_dpocket(dpaths, ["LIG"] * ndf, p, outs, group=512, io_threads=io_threads)
            barrier()
            tq = time.perf_counter() - tq
            qq = torch.tensor([tq], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(qq, op=dist.ReduceOp.MAX)
            out["dpocket"]["file_to_file"] = {"value": ndf * world / float(qq[0]), "unit": "structures/s", "complexes_per_gpu": ndf, "written": sd["n_written"],
                                              "pockets": sd["n_pockets"], "table_rows": sum(max(0, len(open(o).read().splitlines()) - 1) for o in outs),
                                              "includes": "PDB complexes on tmpfs -> fpkio_run_dpocket: reader in ligand mode, detection + ligand queries + explicit pocket "
                                                          "(one fpk_detect_batch per 512 complexes), <code>_out/ directories and the three dpocket tables"}
            subprocess.call(["rm", "-rf", tmpd])
        if rank == 0 and world == 1 and not args.no_cpu:
            exe = os.path.join(ROOT, "oracle", "_ref", "dpocket")
            if os.path.exists(exe):
                nsm = min(nd, 2 * ncores)
                tmp = tempfile.mkdtemp(prefix="fpkdp", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
                jobs = []
                per = max(1, (nsm + ncores - 1) // ncores)
                for c0 in range(0, nsm, per):
                    d = os.path.join(tmp, "j%03d" % (c0 // per))
                    os.makedirs(d)
                    with open(os.path.join(d, "list.txt"), "w") as f:
                        for k in range(c0, min(nsm, c0 + per)):
                            fn = os.path.join(d, "c%05d.pdb" % k)
                            synth.to_pdb(raw[k], fn, ligs[k])
                            f.write("%s LIG\n" % fn)
                    jobs.append(d)
                t0 = time.time()
                procs = [subprocess.Popen([exe, "-f", os.path.join(d, "list.txt")], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in jobs]
                for pr in procs:
                    pr.wait()
                dt = time.time() - t0
                subprocess.call(["rm", "-rf", tmp])
                out["dpocket"]["cpu_baseline"] = {"value": nsm / dt, "unit": "structures/s", "cores": ncores, "kind": "reference",
                                                  "sample": "first %d complexes, oracle/_ref/dpocket -f list (the unmodified reference), %d processes, tmpfs; %.1f s" % (nsm, len(jobs), dt)}

    # ---- optional: one large complex (configs[3]); a single Delaunay does not shard: one block of one GPU ------------------------------
    if args.large > 0 and rank == 0:
        big = synth.to_structure(synth.generate(20260401, args.large, 0.065))
        bb = api.Batch([big], p)
        bb.run()
        bb.run()
        bb.run()
        nb, tb_, sb_ = int(big.n_sites), bb.counter(1), bb.counter(2)
        alg_big = 32 * nb + 32 * tb_ + 36 * sb_
        out["large_complex"] = {"heavy_atoms": nb, "seconds": bb.kernel_ms(7) / 1e3, "tets": tb_, "spheres": sb_,
                                "pockets": bb.counter(4), "tets_per_s": tb_ / max(1e-9, bb.kernel_ms(0) / 1e3),
                                "stage_ms": {"k_delaunay_grid": bb.kernel_ms(0), "k_cluster_grid": bb.kernel_ms(1), "k_hull_warp": bb.kernel_ms(3),
                                             "k_descriptors(+small)": bb.kernel_ms(4)},
                                "roofline": {"bound": "hbm", "kernel": "k_delaunay_grid", "achieved": alg_big / max(1e-9, bb.kernel_ms(0) * 1e-3) / 1e9, "peak": peak,
                                             "unit": "GB/s", "frac": alg_big / max(1e-9, bb.kernel_ms(0) * 1e-3) / 1e9 / peak, "traffic": None,
                                             "algorithmic_bytes_per_launch": alg_big, "ms_per_launch": bb.kernel_ms(0), "peak_source": peak_src},
                                "includes": "BASELINE configs[3]: ONE complex on the whole GPU (cooperative launches: every resident block on the same triangulation / "
                                            "clustering), inputs resident, third run"}
        try:
            out["large_complex"]["reference"] = json.load(open(os.path.join(ROOT, "profiles", "r02_reference_150k.json")))
        except Exception:
            pass
        bb.close()

    # ---- CPU baseline: the unmodified reference on this box's cores, bounded sample ------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        nsample = args.ref_sample or max(ncores * 3, 48)
        r, dt = reference_fpocket_rate(raw, min(nsample, len(raw)), ncores)
        if r is not None:
            out["cpu_baseline"] = {"value": r, "unit": "structures/s", "cores": ncores, "kind": "reference",
                                   "sample": "first %d structures of the workload, one `fpocket -f` process each (oracle/_ref/fpocket, "
                                             "the unmodified reference built with -O2), %d at a time, tmpfs; %.1f s" % (min(nsample, len(raw)), ncores, dt)}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "structures/s", "cores": ncores, "kind": "reference", "sample": str(dt)}
    batch.close()
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""ncu report -> compact text summary for profiles/: key counters, stall mix, hottest source lines.
usage: python scripts/summarize_profile.py gpurun_out/prof_x.ncu-rep > profiles/x_summary.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print("# %s" % rep)
for i, h in enumerate(hdr):
    if h in want:
        print("%-64s %-14s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, hd, out, stall = None, None, [], collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hd = r; continue
    if hd is None or len(r) < len(hd) or r[2] != "-": continue
    d = dict(zip(hd[4:], r[4:]))
    try: smp, inst, thr = int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])
    except Exception: continue
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v.isdigit(): stall[k[6:]] += int(v)
    out.append((smp, inst, thr, fname, r[0], r[1].strip()[:100]))
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print("\n# warp stall samples (all warps): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / sum(stall.values())) for k, v in stall.most_common(8)))
print("# hottest source lines: %samples  %instructions  active lanes  file:line  source")
for o in sorted(out, key=lambda o: -o[0])[:18]:
    print("%5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100.0 * o[0] / ts, 100.0 * o[1] / ti, o[2] / max(1, o[1]), o[3], o[4], o[5]))
print("# most executed source lines: %instructions  %samples  active lanes  file:line  source")
for o in sorted(out, key=lambda o: -o[1])[:24]:
    print("%5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100.0 * o[1] / ti, 100.0 * o[0] / ts, o[2] / max(1, o[1]), o[3], o[4], o[5]))

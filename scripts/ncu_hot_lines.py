#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` dump: samples, instructions, main stall."""
import csv, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname, hdr, out = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name", "File Name"): continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    d = dict(zip(hdr[4:], r[4:]))
    try: samples = int(d["# Samples"]); inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"])
    except Exception: continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    out.append((samples, inst, thr, fname, r[0], r[1].strip()[:90], main))
ts, ti = sum(o[0] for o in out), sum(o[1] for o in out)
print("total samples %d, warp instructions %d" % (ts, ti))
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% smp %5.1f%% inst lanes %4.1f  %s:%s  %s   %s" % (100.0 * o[0] / ts, 100.0 * o[1] / ti, o[2] / max(1, o[1]), o[3], o[4], o[5], o[6]))

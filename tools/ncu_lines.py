#!/usr/bin/env python3
"""Per source line share of executed warp instructions of one profiled launch (ncu --import-source on).

    python tools/ncu_lines.py report.ncu-rep LAUNCH_INDEX [min_pct]
"""
import collections, csv, io, subprocess, sys, glob, os
rep, want = sys.argv[1], int(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
files = ",".join(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point_cloud_registration_b200", "csrc", "*.cu*")))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--resolve-source-file", files],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = None
# simpler: detect launch boundaries by repeated first file name
blocks, cur = [], None
first_file = None
for r in rows:
    if r and r[0] == "File Path":
        f = os.path.basename(r[1])
        if first_file is None:
            first_file = f
        if f == first_file:
            cur = []; blocks.append(cur)
        cur.append(("file", f))
    elif r and r[0] == "Line No":
        cur.append(("hdr", r))
    elif r and cur is not None and r[0] not in ("Function Name",):
        cur.append(("row", r))
b = blocks[want]
tot = 0; per = collections.OrderedDict(); f = None; hdr = None
for kind, r in b:
    if kind == "file": f = r
    elif kind == "hdr": hdr = r
    else:
        try:
            ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed")
            if r[2] != "-" or not r[0].strip().isdigit():      # keep the CUDA source rows (Address "-"): they carry the totals of their SASS
                continue
            e = int(r[ie] or 0); t = int(r[it] or 0)
        except Exception:
            continue
        if e:
            per[(f, r[0], r[1].strip()[:110])] = (e, t); tot += e
print(f"launch {want}: {tot} warp instructions (sum over source lines)")
for (f, ln, src), (e, t) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    if e * 100.0 / tot >= minpct:
        print(f"{e * 100.0 / tot:5.1f}%  lanes {t / e:4.1f}  {f}:{ln}  {src}")

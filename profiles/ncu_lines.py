#!/usr/bin/env python
"""Warp-stall samples / executed instructions per CUDA source line from an .ncu-rep captured with
--import-source on (-lineinfo build):  python profiles/ncu_lines.py rep.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr and r[0].isdigit() and len(r) > ii and r[2] == "-" and r[si].isdigit():
        recs.append((int(r[si]), int(r[ii]) if r[ii].isdigit() else 0, cur_file, int(r[0]), r[1].strip()[:95]))
tot = sum(x[0] for x in recs) or 1
print("total samples", tot)
for s, n, f, ln, src in sorted(recs, reverse=True)[:topn]:
    print(f"{s:8d} {s/tot*100:5.1f}%  inst {n:10d}  {f}:{ln:<4d} {src}")

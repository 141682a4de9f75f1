#!/usr/bin/env python
"""Top source lines of an ncu report by stall samples / executed instructions.
usage: tools/ncu_lines.py report.ncu-rep [N]"""
import csv, subprocess, sys, io, os

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0] not in ("", "-") and len(r) > 8:
        d = dict(zip(hdr, r))
        try:
            lines.append((cur, int(r[0]), r[1].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0),
                          int(d.get("stall_long_sb", 0) or 0), int(d.get("stall_sleep", 0) or 0), int(d.get("stall_membar", 0) or 0)))
        except ValueError:
            pass
ts, ti = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print(f"total samples {ts}  total inst {ti}")
print("--- by samples")
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{l[0][:14]:14s}:{l[1]:4d} smp {100*l[3]/ts:5.1f}% inst {100*l[4]/ti:5.1f}% lsb {l[5]:5d} slp {l[6]:5d} mb {l[7]:4d} | {l[2][:100]}")

#!/usr/bin/env python
"""Hot spots of a single-kernel ncu report (--set full --import-source on), from the SASS page: the instructions with the most
warp-stall samples, with executed count, active threads, shared-memory wavefronts and L1 tag requests.
usage: tools/ncu_hotspots.py report.ncu-rep [top_n] > profiles/<name>_source_hotspots.txt"""
import csv, io, subprocess, sys
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[0][1][:160] if rows[0] and len(rows[0]) > 1 else "")
h = rows[hi]
col = {k: h.index(k) for k in ("Address", "Source", "Warp Stall Sampling (All Samples)", "Instructions Executed", "Avg. Threads Executed",
                               "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global") if k in h}
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
num = lambda r, k: int(float(r[col[k]] or 0)) if k in col else 0
tot_s = sum(num(r, "Warp Stall Sampling (All Samples)") for r in body); tot_i = sum(num(r, "Instructions Executed") for r in body)
print(f"total stall samples {tot_s}  total warp instructions {tot_i}  shared wavefronts {sum(num(r, 'L1 Wavefronts Shared') for r in body)} "
      f"(ideal {sum(num(r, 'L1 Wavefronts Shared Ideal') for r in body)})  global tag requests {sum(num(r, 'L1 Tag Requests Global') for r in body)}")
print("--- by stall samples")
for r in sorted(body, key=lambda r: -num(r, "Warp Stall Sampling (All Samples)"))[:top_n]:
    print(f"{r[col['Address']][-5:]} smp {100.0 * num(r, 'Warp Stall Sampling (All Samples)') / max(tot_s, 1):5.1f}%  exec {num(r, 'Instructions Executed'):>10}  "
          f"thr {r[col['Avg. Threads Executed']]:>5}  shw {num(r, 'L1 Wavefronts Shared'):>9}  tag {num(r, 'L1 Tag Requests Global'):>9} | {r[col['Source']][:80]}")

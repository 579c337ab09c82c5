#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample shares from a .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file = None; items = []; tot_i = tot_s = 0
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and len(r) > ii and r[0].isdigit():
        try: v = int(r[ii]); s = int(r[si])
        except ValueError: continue
        items.append((v, s, cur_file, r[0], r[1].strip()[:110])); tot_i += v; tot_s += s
items.sort(reverse=True)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for v, s, f, ln, src in items[:top]:
    print(f"{100*v/tot_i:5.1f}% inst {100*s/max(1,tot_s):5.1f}% smp  {f}:{ln}  {src}")

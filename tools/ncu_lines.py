#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (captured with --import-source on, -lineinfo).
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, os, subprocess, sys

rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = []            # (file, line, text, inst, thread_inst, samples)
cur_file = "?"; H = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No": H = r; ci = H.index("Instructions Executed"); cs = H.index("# Samples"); ct = H.index("Thread Instructions Executed"); continue
    if H is None or len(r) <= ci or r[0] == "": continue
    try: lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ci]), int(r[ct]), int(r[cs])))
    except ValueError: pass
ti = sum(l[3] for l in lines) or 1; ts = sum(l[5] for l in lines) or 1
print(f"total warp inst {ti} samples {ts} lane eff {sum(l[4] for l in lines) / (32.0 * ti):.2f}")
for f, n, t, i, th, s in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100.0 * i / ti:5.1f}% inst {100.0 * s / ts:5.1f}% samp  eff {th / (32.0 * i + 1e-9):.2f}  {f}:{n} | {t[:110]}")

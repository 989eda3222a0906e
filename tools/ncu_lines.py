#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts from an .ncu-rep (cuda,sass source view).
usage: ncu_lines.py REP LAUNCH_SKIP [file-substr] [top]"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
sub = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None
out = []
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
    elif len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            s, i = int(r[4]), int(r[7])
        except ValueError:
            continue
        tot_s += s; tot_i += i
        out.append((cur.split("/")[-1], int(r[0]), s, i, r[1].strip()[:90]))
print(f"total samples {tot_s}, warp insts {tot_i}")
sel = [o for o in out if sub in o[0]]
if len(sys.argv) > 5 and sys.argv[5] == "byline":
    sel.sort(key=lambda o: (o[0], o[1]))
else:
    sel.sort(key=lambda o: -o[2]); sel = sel[:top]
for f, ln, s, i, src in sel:
    print(f"{f}:{ln:4d} samp {100*s/tot_s:5.1f}% inst {100*i/tot_i:5.1f}%  {src}")

"""Print SASS (with executed counts) attributed to a source line range. usage: ncu_sass_range.py rep file lo hi"""
import csv, subprocess, sys
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()))
cur = None; hdr = None; on = False
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ci = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) < 5: continue
    if r[0] not in ('-', '') and r[2] == '-':
        on = cur == fname and lo <= int(r[0]) <= hi
        if on: print(f"--- {r[0]}: {r[1].strip()[:110]}")
        continue
    if on:
        try: n = int(r[ci]) / 1e6
        except ValueError: n = 0.0
        print(f"   {n:9.1f}M  {r[2]:>6s} {r[3].strip()[:100]}")

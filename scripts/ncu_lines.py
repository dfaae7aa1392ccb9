"""Rank CUDA source lines of an ncu report by warp-stall samples.  usage: ncu_lines.py source_cs.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] not in ('-', '') and r[2] == '-':
        d = dict(zip(hdr[4:], r[4:]))
        try: s = int(d['# Samples'])
        except Exception: continue
        out.append((s, cur, r[0], r[1], d))
tot = sum(o[0] for o in out)
itot = sum(int(o[4]['Instructions Executed']) for o in out)
print("total samples", tot, "total inst", itot)
for s, f, ln, src, d in sorted(out, key=lambda t: -t[0])[:top]:
    print(f"{s:8d} {100*s/tot:5.1f}% inst={int(d['Instructions Executed'])/itot*100:5.1f}% long={d.get('stall_long_sb','?'):>7s} wait={d.get('stall_wait','?'):>6s} short={d.get('stall_short_sb','?'):>6s} {f}:{ln}: {src.strip()[:88]}")

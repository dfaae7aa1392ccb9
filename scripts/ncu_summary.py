"""Summarise one ncu report: top source lines, stall mix, icache, instruction mix, hot footprint.
usage: ncu_summary.py <report.ncu-rep> [top]"""
import collections, csv, subprocess, sys, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout
raw = list(csv.reader(run(["--page", "raw", "--csv"]).splitlines()))
a = dict(zip(raw[0], raw[2]))
print("== metrics")
for k in ['gpu__time_duration.sum', 'sm__icc_request_hit_rate.pct', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
          'dram__bytes.sum.per_second', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread']:
    print(f"  {k:60s} {a.get(k)}")
st = {k[33:]: float(v) for k, v in a.items() if k.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in k and v}
tot = sum(st.values())
print("== stalls:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda t: -t[1])[:9]))
rows = list(csv.reader(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]).splitlines()))
cur = None; hdr = None; line = None; out = []; sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = hdr.index('Instructions Executed'); continue
    if hdr is None: continue
    if r[0] not in ('-', '') and r[2] == '-':
        d = dict(zip(hdr[4:], r[4:]))
        try: out.append((int(d['# Samples']), cur, r[0], r[1], d))
        except Exception: pass
        continue
    try: sass.append((int(r[ci]), r[3]))
    except Exception: pass
ts = sum(o[0] for o in out); ti = sum(int(o[4]['Instructions Executed']) for o in out)
print("== top lines (samples, inst share)")
for s, f, ln, src, d in sorted(out, key=lambda t: -t[0])[:top]:
    print(f"  {100*s/ts:5.1f}% inst={int(d['Instructions Executed'])/ti*100:5.1f}% long={d.get('stall_long_sb','?'):>7s} {f}:{ln}: {src.strip()[:84]}")
vs = sorted([c for c, _ in sass], reverse=True); tot = sum(vs)
for frac in (0.9, 0.99):
    acc = 0
    for i, x in enumerate(vs):
        acc += x
        if acc >= frac * tot: break
    print(f"== {frac*100:.0f}% of executed instructions come from {(i+1)*16/1024:.1f} KB of code ({len(vs)*16//1024} KB kernel)")
c = collections.Counter()
for cnt, src in sass:
    t = src.strip().split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    c[op.split('.')[0]] += cnt
print("== op mix:", ", ".join(f"{op} {100*n/tot:.1f}%" for op, n in c.most_common(12)))

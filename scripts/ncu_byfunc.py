"""Instruction / sample share per device function of ldp_kernel.cuh (by source line range) from an ncu report.
usage: ncu_byfunc.py <report.ncu-rep>"""
import csv, re, subprocess, sys, collections
rep = sys.argv[1]
src = open(sys.argv[2] if len(sys.argv)>2 else 'daqp_b200/csrc/ldp_kernel.cuh').read().splitlines()
# function starts: lines with '__device__' and '(' -> name
starts = []
for i, l in enumerate(src, 1):
    m = re.search(r'__device__[^;=]*?\b(\w+)\s*\(', l)
    if m and not l.strip().startswith('//'): starts.append((i, m.group(1)))
    m = re.search(r'__global__.*?\b(\w+)\s*\(', l)
    if m: starts.append((i, m.group(1)))
def func(ln):
    name = 'top'
    for s, n in starts:
        if s <= ln: name = n
        else: break
    return name
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()))
cur = None; hdr = None
inst = collections.Counter(); samp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 5 or r[0] in ('-', '') or r[2] != '-': continue
    d = dict(zip(hdr[4:], r[4:]))
    try: i = int(d['Instructions Executed']); s = int(d['# Samples'])
    except Exception: continue
    key = func(int(r[0])) if cur in ('ldp_kernel.cuh','setup_kernel.cuh') else cur
    inst[key] += i; samp[key] += s
ti = sum(inst.values()); ts = sum(samp.values())
for k, v in inst.most_common(30):
    print(f"{k:28s} inst {100*v/ti:5.1f}%  samples {100*samp[k]/ts:5.1f}%  ({v/1e6:.0f}M)")

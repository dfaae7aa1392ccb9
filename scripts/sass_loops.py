"""Static view of a cubin's loops: for every backward branch, the loop's instruction count and the source lines it
covers (needs -lineinfo). Offline stand-in for the profiler's per-line instruction counts.
usage: sass_loops.py <file.cubin> [--body LABEL] [--min N]
  --body LABEL  print the SASS of the loop that branches back to LABEL"""
import re, subprocess, sys, collections
cubin = sys.argv[1]
body = sys.argv[sys.argv.index('--body') + 1] if '--body' in sys.argv else None
txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
ins = []      # (addr, text, file, line)
labels = {}   # label -> index of the next instruction
cur = (None, None)
for l in txt:
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'^(\.L_x_\d+):', l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip(), cur[0], cur[1]))
loops = []
for i, (a, t, f, ln) in enumerate(ins):
    m = re.search(r'BRA(?:\.\w+)*\s+(?:!?\w+,\s*)?`\((\.L_x_\d+)\)', t)
    if m and m.group(1) in labels and labels[m.group(1)] <= i and not t.startswith('BRA.DIV'):
        loops.append((labels[m.group(1)], i, m.group(1)))
for s, e, lab in sorted(loops):
    if body:
        if lab != body: continue
        for a, t, f, ln in ins[s:e + 1]: print(f"  {a:06x} {f}:{ln:<5d} {t}")
        continue
    lines = collections.Counter((f, ln) for a, t, f, ln in ins[s:e + 1])
    files = collections.Counter(f for a, t, f, ln in ins[s:e + 1])
    main = [k for k in lines if k[0] == 'ldp_kernel.cuh' or k[0] == 'setup_kernel.cuh']
    lo = min((k[1] for k in main), default=0); hi = max((k[1] for k in main), default=0)
    ops = collections.Counter(t.split()[1].split('.')[0] if t.startswith('@') else t.split()[0].split('.')[0] for a, t, f, ln in ins[s:e + 1])
    print(f"{lab:10s} {e - s + 1:5d} instr  lines {lo}-{hi}  " + ", ".join(f"{o}:{c}" for o, c in ops.most_common(8)))

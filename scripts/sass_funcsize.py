"""Static code size per device function (by source line range) of a cubin built with -lineinfo.
usage: sass_funcsize.py <cubin> <source.cuh>"""
import re, subprocess, sys, collections
cubin, srcf = sys.argv[1], sys.argv[2]
src = open(srcf).read().splitlines()
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
base = srcf.split('/')[-1]
cur = None; cnt = collections.Counter()
for l in subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines():
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: f = m.group(1).split('/')[-1]; cur = func(int(m.group(2))) if f == base else f; continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l) and cur: cnt[cur] += 1
for k, v in cnt.most_common(22): print(f"{k:28s} {v:6d} instr {v*16/1024:6.1f} KB")
print("total", sum(cnt.values()))

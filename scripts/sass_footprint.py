"""Static SASS footprint per CUDA source line (offline, no GPU): which source lines the instructions of one kernel
come from. usage: sass_footprint.py <lib.so> <kernel-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line, counts, total = None, None, collections.Counter(), 0
ops = collections.Counter()
for ln in out.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur_fn and pat in cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        counts[cur_line] += 1; total += 1
        mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if mm: ops[mm.group(2)] += 1
print(f"kernel *{pat}*: {total} instructions = {total*16/1024:.1f} KB")
for (f, l), c in counts.most_common(top):
    print(f"{c:6d} {f}:{l}")
print("opcodes:", ops.most_common(12))

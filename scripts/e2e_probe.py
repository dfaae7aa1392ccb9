"""e2e (host C ABI) throughput vs host chunk size, next to the raw pinned H2D bandwidth. usage: e2e_probe.py [P]"""
import os, sys, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
code = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, daqp_b200
from daqp_b200.problems import SEED_BASE, generate_g1_torch
P = int(sys.argv[1]); n, m, ms = 50, 150, 0
dev = torch.device("cuda:0")
t = generate_g1_torch(P, n, m, ms, 40, 100.0, seed=SEED_BASE + 3, device=dev)
eng = daqp_b200.Engine(0)
pin = lambda x: x.cpu().pin_memory()
h = {k: pin(t[k]).numpy() for k in ("H", "f", "A", "bupper", "blower")}
res = daqp_b200.BatchResult(x=torch.empty((P, n), dtype=torch.float64).pin_memory().numpy(),
                            lam=torch.empty((P, m), dtype=torch.float64).pin_memory().numpy(),
                            fval=torch.zeros(P, dtype=torch.float64).pin_memory().numpy(),
                            exitflag=torch.empty(P, dtype=torch.int32).pin_memory().numpy(),
                            iter=torch.empty(P, dtype=torch.int32).pin_memory().numpy())
def step(): eng.solve_batch(h["H"], h["f"], h["A"], h["bupper"], h["blower"], None, ms=ms, out=res)
step(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print("chunk", os.environ.get("DAQP_B200_HOST_CHUNK"), "ms/step %.1f" % (dt * 1e3), "QP/s %.0f" % (P / dt), flush=True)
'''
open("/tmp/e2e_one.py", "w").write(code)
a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(a, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw pinned H2D GB/s: %.1f" % (1.0737 / dt), flush=True)
del a, d
for chunk, first in ((8192, 2048), (4096, 2048), (8192, 1024), (6144, 1024), (12288, 2048)):
    env = dict(os.environ, DAQP_B200_HOST_CHUNK=str(chunk), DAQP_B200_HOST_FIRST_CHUNK=str(first))
    print("first", first, end=" ", flush=True)
    subprocess.run([sys.executable, "/tmp/e2e_one.py", str(P)], env=env)

"""Per shape group of the C5 size mix (fp32): setup / solve ms and QP/s when the group runs alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daqp_b200
from daqp_b200.problems import generate_g1_torch

dev = torch.device("cuda:0")
eng = daqp_b200.Engine(0)
per = 12500
tot = 0.0
for k, n in enumerate(range(8, 129, 8)):
    t = generate_g1_torch(per, n, 4 * n, 0, n, seed=500 + k, device=dev, random_nactive=True)
    g = {k2: (v.float().contiguous() if v.dtype == torch.float64 else v) for k2, v in t.items()}
    fn = lambda: eng.solve_batch_device_f32(g["H"], g["f"], g["A"], g["bupper"], g["blower"], None, ms=0)
    fn(); torch.cuda.synchronize(); eng.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    st = eng.stats(reset=True)
    ms = e0.elapsed_time(e1); tot += ms
    print(json.dumps({"n": n, "m": 4 * n, "ms": round(ms, 2), "setup_ms": round(st["setup_ms"], 2), "solve_ms": round(st["solve_ms"], 2),
                      "mean_iter": round(float(r["iter"].double().mean()), 1), "per_sm": st["warps_per_sm"]}))
    del t, g
print("sum of groups run alone: %.1f ms" % tot)

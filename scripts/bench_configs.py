"""Device-resident throughput on the other BASELINE.json configurations (informational: bench.py stays on C3).
  C2: 10 000 x (n=20, m=60) fp64            C4: 50 000 x (n=120, m=400, ms=120) fp64, cold and warm-started (MPC step)
  C5: 200 000 mixed sizes n in {8..128}, m = 4 n, random active-set sizes, fp32, one launch pair per shape group
usage: python scripts/bench_configs.py [--scale 1.0]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daqp_b200
from daqp_b200.problems import generate_g1_torch

ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=1.0); args = ap.parse_args()
dev = torch.device("cuda:0")
eng = daqp_b200.Engine(0)
out = {}


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


# ---- C2
N = int(10_000 * args.scale)
t = generate_g1_torch(N, 20, 60, 0, 16, seed=2, device=dev)
ms_, r = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], None, ms=0))
assert bool((r["exitflag"] == 1).all()) and float((r["x"] - t["xref"]).abs().max()) < 1e-6
out["C2"] = {"N": N, "ms_per_batch": ms_, "qps": N / ms_ * 1e3, "mean_iter": float(r["iter"].double().mean())}

# ---- C4 cold, then warm-started from the active set of a neighbour whose f differs by 5 %
N = int(50_000 * args.scale)
t = generate_g1_torch(N, 120, 400, 120, 96, seed=4, device=dev)
ms_, r = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], None, ms=120), reps=2)
assert bool((r["exitflag"] == 1).all()) and float((r["x"] - t["xref"]).abs().max()) < 1e-5
out["C4_cold"] = {"N": N, "ms_per_batch": ms_, "qps": N / ms_ * 1e3, "mean_iter": float(r["iter"].double().mean())}
g = torch.Generator(device=dev); g.manual_seed(44)
fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=120)
torch.cuda.synchronize()
sense = torch.zeros((N, 400), dtype=torch.int32, device=dev)
sense[rn["lam"] > 1e-12] = 1
sense[rn["lam"] < -1e-12] = 3
ms_, r2 = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=120), reps=2)
assert bool((r2["exitflag"] == 1).all()) and float((r2["x"] - t["xref"]).abs().max()) < 1e-5
out["C4_warm"] = {"N": N, "ms_per_batch": ms_, "qps": N / ms_ * 1e3, "mean_iter": float(r2["iter"].double().mean())}
del t, r, r2, rn, sense, fn_
torch.cuda.empty_cache()

# ---- C5: fp32, sixteen shape groups
sizes = list(range(8, 129, 8))
per = int(200_000 * args.scale) // len(sizes)
groups = []
for k, n in enumerate(sizes):
    t = generate_g1_torch(per, n, 4 * n, 0, n, seed=500 + k, device=dev, random_nactive=True)
    groups.append({k2: (v.float().contiguous() if v.dtype == torch.float64 else v) for k2, v in t.items()})
    del t
torch.cuda.empty_cache()


def c5():
    return [eng.solve_batch_device_f32(gp["H"], gp["f"], gp["A"], gp["bupper"], gp["blower"], None, ms=0) for gp in groups]


ms_, rs = timed(c5, reps=2)
opt = sum(int((r["exitflag"] == 1).sum()) for r in rs)
err = max(float(((r["x"] - gp["xref"]).abs().max(dim=1).values / (1 + gp["xref"].abs().max(dim=1).values))[r["exitflag"] == 1].max())
          for r, gp in zip(rs, groups))
its = torch.cat([r["iter"] for r in rs]).double()
out["C5_fp32"] = {"N": per * len(sizes), "ms_per_batch": ms_, "qps": per * len(sizes) / ms_ * 1e3,
                  "optimal_fraction": opt / (per * len(sizes)), "max_rel_x_err_vs_constructed_optimum": err,
                  "iter_min_mean_max": [float(its.min()), float(its.mean()), float(its.max())]}
print(json.dumps(out))

"""C4 (n=120, m=400, ms=120) device-resident throughput, cold and warm-started: python scripts/bench_c4.py [--n N]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daqp_b200
from daqp_b200.problems import generate_g1_torch

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=50000); ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
eng = daqp_b200.Engine(0)
N = args.n
t = generate_g1_torch(N, 120, 400, 120, 96, seed=4, device=dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize(); eng.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    st = eng.stats(reset=True)
    return e0.elapsed_time(e1) / reps, r, st


out = {"team_env": os.environ.get("DAQP_B200_TEAM", "default"), "tune": os.environ.get("DAQP_B200_TUNE", "0")}
ms_, r, st = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], None, ms=120), args.reps)
assert bool((r["exitflag"] == 1).all()) and float((r["x"] - t["xref"]).abs().max()) < 1e-5
out["C4_cold"] = {"N": N, "ms_per_batch": ms_, "qps": N / ms_ * 1e3, "mean_iter": float(r["iter"].double().mean()),
                  "setup_ms": st["setup_ms"] / args.reps, "solve_ms": st["solve_ms"] / args.reps, "resident_per_sm": st["warps_per_sm"]}
g = torch.Generator(device=dev); g.manual_seed(44)
fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=120)
torch.cuda.synchronize()
sense = torch.zeros((N, 400), dtype=torch.int32, device=dev)
sense[rn["lam"] > 1e-12] = 1
sense[rn["lam"] < -1e-12] = 3
ms_, r2, st = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=120), args.reps)
assert bool((r2["exitflag"] == 1).all()) and float((r2["x"] - t["xref"]).abs().max()) < 1e-5
out["C4_warm"] = {"N": N, "ms_per_batch": ms_, "qps": N / ms_ * 1e3, "mean_iter": float(r2["iter"].double().mean()),
                  "setup_ms": st["setup_ms"] / args.reps, "solve_ms": st["solve_ms"] / args.reps}
print(json.dumps(out))

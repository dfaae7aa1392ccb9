"""Device-resident throughput of a G1 batch cold and warm-started from a perturbed neighbour's active set (MPC step):
python scripts/bench_warm.py --n 50 --m 150 --ms 0 --nact 40 --N 100000"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daqp_b200
from daqp_b200.problems import generate_g1_torch

ap = argparse.ArgumentParser()
for k, v in (("n", 50), ("m", 150), ("ms", 0), ("nact", 40), ("N", 100000), ("reps", 3)):
    ap.add_argument("--" + k, type=int, default=v)
args = ap.parse_args()
dev = torch.device("cuda:0")
eng = daqp_b200.Engine(0)
t = generate_g1_torch(args.N, args.n, args.m, args.ms, args.nact, seed=4, device=dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize(); eng.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    st = eng.stats(reset=True)
    return e0.elapsed_time(e1) / reps, r, st


out = {"shape": [args.n, args.m, args.ms, args.nact], "N": args.N}
ms_, r, st = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], None, ms=args.ms), args.reps)
out["cold"] = {"ms": ms_, "qps": args.N / ms_ * 1e3, "mean_iter": float(r["iter"].double().mean()), "setup_ms": st["setup_ms"] / args.reps,
               "solve_ms": st["solve_ms"] / args.reps}
g = torch.Generator(device=dev); g.manual_seed(44)
fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=args.ms)
torch.cuda.synchronize()
sense = torch.zeros((args.N, args.m), dtype=torch.int32, device=dev)
sense[rn["lam"] > 1e-12] = 1
sense[rn["lam"] < -1e-12] = 3
ms_, r2, st = timed(lambda: eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=args.ms), args.reps)
assert bool((r2["exitflag"] == 1).all()) and float((r2["x"] - t["xref"]).abs().max()) < 1e-5
out["warm"] = {"ms": ms_, "qps": args.N / ms_ * 1e3, "mean_iter": float(r2["iter"].double().mean()), "setup_ms": st["setup_ms"] / args.reps,
               "solve_ms": st["solve_ms"] / args.reps, "mean_active_preset": float((sense != 0).sum(dim=1).double().mean())}
print(json.dumps(out))

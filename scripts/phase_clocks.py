"""Where the leader's cycles go, per phase (library built with -DDAQP_B200_PHASE_CLOCKS: DAQPB200Diag.counts then holds
clock cycles / 16 per phase instead of the path counters): python scripts/phase_clocks.py --n 120 --m 400 --ms 120 --nact 96"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daqp_b200
from daqp_b200.problems import generate_g1_torch

ap = argparse.ArgumentParser()
for k, v in (("n", 120), ("m", 400), ("ms", 120), ("nact", 96), ("N", 1332)):
    ap.add_argument("--" + k, type=int, default=v)
args = ap.parse_args()
dev = torch.device("cuda:0")
eng = daqp_b200.Engine(0)
t = generate_g1_torch(args.N, args.n, args.m, args.ms, args.nact, seed=4, device=dev)
for warm in (False, True):
    sense = None
    if warm:
        g = torch.Generator(device=dev); g.manual_seed(44)
        fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
        rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=args.ms)
        sense = torch.zeros((args.N, args.m), dtype=torch.int32, device=dev)
        sense[rn["lam"] > 1e-12] = 1
        sense[rn["lam"] < -1e-12] = 3
    diag = daqp_b200.Engine.alloc_diag(args.N, args.n, args.m, dev)
    r = eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=args.ms, diag=diag)
    torch.cuda.synchronize()
    c = diag["counts"].double().mean(dim=0) * 16
    it = r["iter"].double().mean().item()
    names = ["csp", "ratio", "primal", "scan", "add", "remove", "-", "activate"]
    c[6] = 0
    tot = c.sum().item()
    print(("warm" if warm else "cold"), f"mean iterations {it:.1f}; leader cycles per problem {tot:.0f} ({tot / it:.0f} per iteration)")
    print("   " + ", ".join(f"{n} {100 * v / tot:.1f}%" for n, v in zip(names, c.tolist()) if n != "-"))
    print(f"   exact scans per problem (team mode only): {diag['counts'][:, 6].double().mean().item():.1f}")

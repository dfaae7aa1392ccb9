"""Cycles of the two passes of a team activation (library built with -DDAQP_B200_PHASE_CLOCKS=2): python scripts/activation_clocks.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, daqp_b200
from daqp_b200.problems import generate_g1_torch
dev = torch.device("cuda:0"); eng = daqp_b200.Engine(0)
N=1332; t = generate_g1_torch(N, 120, 400, 120, 96, seed=4, device=dev)
g = torch.Generator(device=dev); g.manual_seed(44)
fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=120)
sense = torch.zeros((N, 400), dtype=torch.int32, device=dev); sense[rn["lam"] > 1e-12] = 1; sense[rn["lam"] < -1e-12] = 3
for rep in range(2):
    diag = daqp_b200.Engine.alloc_diag(N, 120, 400, dev)
    r = eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=120, diag=diag)
    torch.cuda.synchronize()
    c = diag["counts"].double().mean(dim=0) * 16
    print("activate total %.0f cycles; gram %.0f; ldl %.0f (+ removes in slot 5); K mean %.1f" % (c[7], c[6], c[5], float((sense!=0).sum(1).double().mean())))

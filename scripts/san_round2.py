"""Small inputs through the kernels added in round 2, for compute-sanitizer (scripts/gpu_sanitize2.sh): team mode cold and
warm-started (two-pass activation), the split transform, the raw LDP batch, branch and bound, the decision log."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daqp_b200 as dq
from daqp_b200.problems import generate_g1, generate_miqp

eng = dq.Engine()
b = generate_g1(6, 70, 120, 10, 40, seed=5)
r = eng.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, None, ms=b.ms)
assert (r.exitflag == 1).all()
sense = np.zeros((b.N, b.m), np.intc)
sense[r.lam > 1e-12] = 1
sense[r.lam < -1e-12] = 3
r2 = eng.solve_batch(b.H, b.f * 1.02, b.A, b.bupper, b.blower, sense, ms=b.ms, diag=True)
assert (r2.exitflag == 1).all()
print("team cold iters", r.iter.tolist(), "warm", r2.iter.tolist())
b = generate_g1(8, 20, 50, 4, 12, seed=6)
r = eng.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, None, ms=b.ms, time_limit=10.0)  # extended instantiation
assert (r.exitflag == 1).all()
rng = np.random.default_rng(1)
A = rng.standard_normal((4, 20, 8)); x0 = rng.standard_normal((4, 8))
c = np.einsum("pij,pj->pi", A, x0)
rl = dq.ldp_batch(A, c + 0.5, c - 0.5)
assert (rl.exitflag == 1).all()
m = generate_miqp(1, 6, 10, 6, 4, seed=9)
x, fval, flag, info = dq.solve(m.H[0], m.f[0], m.A[0], m.bupper[0], m.blower[0], m.sense[0])
assert flag == 1
print("ok")

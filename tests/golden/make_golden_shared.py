"""Golden vectors for the SHARED workspace (daqp_b200_workspace_setup_shared): G matrix sets (H, A), K problems per set
that differ only in f and the bounds -- one controller evaluated for many states. Produced by the UNMODIFIED reference
compiled into oracle/_ref, one workspace per problem: setup_daqp(H_g, f_p, A_g, b_p) + daqp_solve(), then
daqp_update_ldp(DAQP_UPDATE_v + DAQP_UPDATE_d) + daqp_solve() per step (src/api.c:88-160,214-260, src/utils.c:58-221).
Run in the build container (needs /root/reference):

    python tests/golden/make_golden_shared.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness  # noqa: E402
from daqp_b200.problems import QPBatch, generate_g1  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def parametric(base: QPBatch, K: int, seed: int, df=0.3, db=0.05):
    """K draws of (f, bupper, blower) around every base problem: f moves by df (relative), both bounds shift together
    by db * N(0, 1). Returns the replicated batch (problem p = g * K + k)."""
    rng = np.random.default_rng(seed)
    G = base.N
    rep = lambda a: np.repeat(a, K, axis=0)
    f = rep(base.f) * (1 + df * rng.standard_normal((G * K, base.n)))
    sh = db * rng.standard_normal((G * K, base.m))
    return QPBatch(base.n, base.m, base.ms, rep(base.H), f, rep(base.A), rep(base.bupper) + sh, rep(base.blower) + sh,
                   rep(base.sense))


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    cases = {"wsshared_n20_m60_ms5": (generate_g1(3, 20, 60, 5, 16, seed=931), 16),
             "wsshared_n50_m150": (generate_g1(2, 50, 150, 0, 40, seed=932), 8),
             "wsshared_n10_m30": (generate_g1(5, 10, 30, 0, 8, seed=933), 12)}
    for name, (base, K) in cases.items():
        b = parametric(base, K, seed=sum(map(ord, name)))
        rng = np.random.default_rng(7)
        steps = []
        for _ in range(2):
            f = b.f * (1 + 0.05 * rng.standard_normal(b.f.shape))
            sh = 0.03 * rng.standard_normal(b.bupper.shape)
            steps.append((f, b.bupper + sh, b.blower + sh))
        sols = harness.ref_solve_sequence(ref, b, steps, use_sense=False)
        cap = max(max((len(w) for w in s.ws), default=0) for s in sols) + 1
        out = dict(n=b.n, m=b.m, ms=b.ms, G=base.N, Kp=K, H=base.H, A=base.A, f=b.f, bupper=b.bupper, blower=b.blower,
                   K=len(steps))
        for k, (f, bu, bl) in enumerate(steps):
            out[f"f{k}"] = f; out[f"bu{k}"] = bu; out[f"bl{k}"] = bl
        for k, s in enumerate(sols):
            ws = np.full((b.N, cap), -1, np.int32)
            for p, w in enumerate(s.ws):
                ws[p, :len(w)] = w
            out[f"x_{k}"] = s.x; out[f"lam_{k}"] = s.lam; out[f"fval_{k}"] = s.fval; out[f"flag_{k}"] = s.exitflag
            out[f"iter_{k}"] = s.iter; out[f"ws_{k}"] = ws
            out[f"nact_{k}"] = np.array([len(w) for w in s.ws], np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "iters per solve:", [float(s.iter.mean()) for s in sols], "flags:",
              [dict(zip(*np.unique(s.exitflag, return_counts=True))) for s in sols])


if __name__ == "__main__":
    main()

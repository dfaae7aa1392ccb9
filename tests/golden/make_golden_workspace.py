"""Golden vectors for the persistent-workspace path, produced by the UNMODIFIED reference compiled into oracle/_ref:
setup_daqp() + daqp_solve(), then K times daqp_update_ldp(DAQP_UPDATE_v + DAQP_UPDATE_d) + daqp_solve() on the kept
workspace (reference src/api.c:88-160,214-260, src/utils.c:58-221). Run in the build container (needs /root/reference):

    python tests/golden/make_golden_workspace.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness  # noqa: E402
from daqp_b200.problems import generate_g1, soften  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mpc_steps(b, K, seed, df=0.05, db=0.03):
    """K perturbed (f, bupper, blower) triples: the linear term moves by df (relative, the MPC state change), the
    bounds shift by db * N(0,1) while keeping bupper >= blower."""
    rng = np.random.default_rng(seed)
    steps = []
    for _ in range(K):
        f = b.f * (1 + df * rng.standard_normal(b.f.shape))
        sh = db * rng.standard_normal(b.bupper.shape)
        steps.append((f, b.bupper + sh, b.blower + sh))
    return steps


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    cases = {
        "wsseq_n10_m30_ms3": (generate_g1(48, 10, 30, 3, 8, seed=901), False),
        "wsseq_n20_m60_ms5": (generate_g1(32, 20, 60, 5, 16, seed=902), False),
        "wsseq_n50_m150": (generate_g1(12, 50, 150, 0, 40, seed=903), False),
        "wsseq_soft_n12_m40": (soften(generate_g1(24, 12, 40, 4, 10, seed=904), 0.3, 0.5, 9), True),
    }
    b = generate_g1(24, 16, 48, 0, 12, seed=905)  # equalities among the constraints
    for p in range(b.N):
        b.sense[p, np.nonzero(b.active_ref[p])[0][:3]] = 5
    cases["wsseq_equalities_n16_m48"] = (b, True)
    others = [harness.RefLib("libdaqp_ref_strict.so"), harness.RefLib("libdaqp_ref_fma.so")]
    for name, (b, use_sense) in cases.items():
        steps = mpc_steps(b, 3, seed=sum(map(ord, name)))
        sols = harness.ref_solve_sequence(ref, b, steps, use_sense=use_sense)
        # keep the problems whose PATH the reference itself agrees on: three builds of the same sources (default
        # -fassociative-math, no reassociation, FMA contraction) must give the same exit flags, iteration counts and
        # working sets (in factor order) in every solve of the sequence. (Soft-constrained problems with more active
        # rows than variables are where they part: near-ties between equally violated rows.)
        stable = np.ones(b.N, bool)
        for lib in others:
            alt = harness.ref_solve_sequence(lib, b, steps, use_sense=use_sense)
            for s0, s1 in zip(sols, alt):
                stable &= (s0.exitflag == s1.exitflag) & (s0.iter == s1.iter)
                stable &= np.array([list(u) == list(v) for u, v in zip(s0.ws, s1.ws)])
        if not stable.all():
            keep = np.nonzero(stable)[0]
            print(name, "dropping path-unstable problems", np.nonzero(~stable)[0].tolist())
            pick = lambda a: None if a is None else np.ascontiguousarray(a[keep])
            b = type(b)(b.n, b.m, b.ms, pick(b.H), pick(b.f), pick(b.A), pick(b.bupper), pick(b.blower), pick(b.sense),
                        pick(b.xref), pick(b.active_ref))
            steps = [tuple(np.ascontiguousarray(a[keep]) for a in st) for st in steps]
            sols = harness.ref_solve_sequence(ref, b, steps, use_sense=use_sense)
        cap = max(max((len(w) for w in s.ws), default=0) for s in sols) + 1
        out = dict(n=b.n, m=b.m, ms=b.ms, H=b.H, f=b.f, A=b.A, bupper=b.bupper, blower=b.blower, sense=b.sense,
                   use_sense=use_sense, K=len(steps))
        for k, (f, bu, bl) in enumerate(steps):
            out[f"f{k}"] = f; out[f"bu{k}"] = bu; out[f"bl{k}"] = bl
        for k, s in enumerate(sols):
            ws = np.full((b.N, cap), -1, np.int32)
            for p, w in enumerate(s.ws):
                ws[p, :len(w)] = w
            out[f"x_{k}"] = s.x; out[f"lam_{k}"] = s.lam; out[f"fval_{k}"] = s.fval; out[f"flag_{k}"] = s.exitflag
            out[f"iter_{k}"] = s.iter; out[f"ws_{k}"] = ws; out[f"slack_{k}"] = s.soft_slack
            out[f"nact_{k}"] = np.array([len(w) for w in s.ws], np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "iters per solve:", [float(s.iter.mean()) for s in sols], "flags:",
              [dict(zip(*np.unique(s.exitflag, return_counts=True))) for s in sols])


if __name__ == "__main__":
    main()

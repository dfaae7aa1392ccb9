"""Golden vectors for PURE LDP inputs -- daqp_quadprog with H == NULL and f == NULL: minimise |x|^2 subject to
blower <= [I(ms); A] x <= bupper (reference src/utils.c:103-110: Rinv stays NULL, M = A normalised; the form Julia's
polyhedral tools use, interfaces/daqp-julia/src/api.jl:440-459). Produced by the UNMODIFIED reference
(oracle/_ref/libdaqp_ref.so and the strict build, which must agree). Run in the build container:

    python tests/golden/make_golden_ldp.py

Feasible and infeasible polyhedra, with and without simple bounds, with equality rows, with warm-start bits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from daqp_b200.problems import QPBatch  # noqa: E402
from oracle import harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make(N, n, m, ms, seed, shift=1.0, neq=0, infeasible_every=0):
    """Random polyhedron around a point x0 with |x0| ~ shift: the origin is (mostly) outside, so the projection is not
    trivial. Every `infeasible_every`-th problem gets two contradicting rows."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    mA = m - ms
    A = rng.standard_normal((N, mA, n))
    x0 = shift * rng.standard_normal((N, n))
    full = np.concatenate([np.broadcast_to(np.eye(n)[:ms], (N, ms, n)), A], axis=1)
    c = np.einsum("pij,pj->pi", full, x0)
    bu = c + rng.uniform(0.05, 1.0, (N, m))
    bl = c - rng.uniform(0.05, 1.0, (N, m))
    bl[rng.uniform(size=(N, m)) < 0.3] = -1e30  # one-sided rows
    sense = np.zeros((N, m), np.int32)
    for p in range(N):
        eq = rng.choice(np.arange(ms, m), size=neq, replace=False) if neq else []
        for i in eq:
            bl[p, i] = bu[p, i] = c[p, i]
            sense[p, i] = 5
        if infeasible_every and p % infeasible_every == infeasible_every - 1 and mA >= 2:
            A[p, 1] = A[p, 0]  # the same row twice with disjoint intervals
            bu[p, ms + 1] = bl[p, ms + 0] - 0.5 if bl[p, ms] > -1e29 else bu[p, ms] - 2.0
            bl[p, ms + 1] = -1e30
            if bl[p, ms] < -1e29:
                bl[p, ms] = bu[p, ms] - 1.0
                bu[p, ms + 1] = bl[p, ms] - 0.5
    return QPBatch(n, m, ms, np.broadcast_to(np.eye(n), (N, n, n)).copy(), None, A, bu, bl, sense)


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    strict = harness.RefLib("libdaqp_ref_strict.so")
    cases = {"ldp_n10_m30": make(24, 10, 30, 0, 901, infeasible_every=6),
             "ldp_n12_m40_ms6": make(24, 12, 40, 6, 902, infeasible_every=8),
             "ldp_n50_m150": make(12, 50, 150, 0, 903, shift=0.5),
             "ldp_n20_m60_eq4": make(16, 20, 60, 4, 904, neq=4),
             "ldp_n70_m200_ms10": make(8, 70, 200, 10, 905, shift=0.5)}
    for name, b in cases.items():
        use_sense = bool(b.sense.any())
        sol = ref.solve(b, use_sense=use_sense, null_H=True, want_ws=False)
        s2 = strict.solve(b, use_sense=use_sense, null_H=True, want_ws=True)
        keep = (sol.exitflag == s2.exitflag) & (sol.iter == s2.iter)  # paths both builds agree on
        b = QPBatch(b.n, b.m, b.ms, b.H[keep], None, b.A[keep], b.bupper[keep], b.blower[keep], b.sense[keep])
        ws = np.full((b.N, b.n + 1), -1, np.int32)
        kept = np.nonzero(keep)[0]
        for q, p in enumerate(kept):
            ws[q, :len(s2.ws[p])] = s2.ws[p]
        nact = np.array([len(s2.ws[p]) for p in kept], np.int32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), n=b.n, m=b.m, ms=b.ms, A=b.A, bupper=b.bupper, blower=b.blower,
                            sense=b.sense, use_sense=int(use_sense), x=sol.x[keep], lam=sol.lam[keep], exitflag=sol.exitflag[keep],
                            iter=sol.iter[keep], ws=ws, n_active=nact)
        print(f"{name:22s} N={b.N} (dropped {int((~keep).sum())}) flags="
              f"{dict(zip(*[v.tolist() for v in np.unique(sol.exitflag[keep], return_counts=True)]))} "
              f"iters {sol.iter[keep].min()}..{sol.iter[keep].max()}")


if __name__ == "__main__":
    main()

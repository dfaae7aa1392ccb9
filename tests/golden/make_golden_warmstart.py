"""Golden vectors for the warm-start initialisers, produced by the UNMODIFIED reference compiled into oracle/_ref:
daqp_primal_init_active / daqp_dual_init_active (include/api.h:57-58, src/api.c:577-631) on seeded problems and
iterates, followed by daqp_quadprog with the resulting sense. Run in the build container (needs /root/reference):

    python tests/golden/make_golden_warmstart.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness  # noqa: E402
from daqp_b200.problems import generate_g1  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    cases = {"warmstart_n10_m30_ms3": generate_g1(48, 10, 30, 3, 8, seed=941),
             "warmstart_n20_m60_ms5": generate_g1(32, 20, 60, 5, 16, seed=942),
             "warmstart_n50_m150": generate_g1(12, 50, 150, 0, 40, seed=943),
             "warmstart_n40_m100_ms40": generate_g1(12, 40, 100, 40, 30, seed=944)}
    for name, b in cases.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        sol = ref.solve(b, use_sense=False)
        assert (sol.exitflag == 1).all()
        # iterates: the optimum, the optimum moved inside / outside the 1e-9 window, a far-away point
        x_in = sol.x + 2e-11 * rng.standard_normal(sol.x.shape)
        x_out = sol.x + 3e-8 * rng.standard_normal(sol.x.shape)
        x_far = sol.x + 0.1 * rng.standard_normal(sol.x.shape)
        lam_noisy = sol.lam + 1e-13 * rng.standard_normal(sol.lam.shape)  # noise below the 1e-12 threshold on inactive rows
        # a starting sense with IMMUTABLE rows (left alone), stale ACTIVE / LOWER bits, and an equality pair
        sense0 = np.zeros((b.N, b.m), np.int32)
        sense0[rng.random((b.N, b.m)) < 0.1] = 4
        sense0[rng.random((b.N, b.m)) < 0.1] |= 2
        out = dict(n=b.n, m=b.m, ms=b.ms, H=b.H, f=b.f, A=b.A, bupper=b.bupper, blower=b.blower, x_opt=sol.x,
                   lam_opt=sol.lam, x_in=x_in, x_out=x_out, x_far=x_far, lam_noisy=lam_noisy, sense0=sense0,
                   iter_cold=sol.iter)
        for key, kw in [("x_opt", dict(x=sol.x)), ("x_in", dict(x=x_in)), ("x_out", dict(x=x_out)), ("x_far", dict(x=x_far)),
                        ("lam_opt", dict(lam=sol.lam)), ("lam_noisy", dict(lam=lam_noisy))]:
            for tag, s0 in (("", None), ("_s0", sense0)):
                se = harness.init_active(b, which="ref", sense=(np.zeros_like(sense0) if s0 is None else s0), **kw)
                assert (se == harness.init_active(b, which="ref", name="libdaqp_ref_strict.so",
                                                  sense=(np.zeros_like(sense0) if s0 is None else s0), **kw)).all()
                out[f"sense_{key}{tag}"] = se
        # the warm-started solves the reference runs from those bits (core_tests.jl:520-543: iterations == 1 from the optimum)
        for key in ("x_opt", "lam_opt", "x_out"):
            bb = b.astype(np.float64)
            bb.sense = out[f"sense_{key}"].copy()
            s = ref.solve(bb, use_sense=True)
            out[f"flag_{key}"] = s.exitflag; out[f"iter_{key}"] = s.iter; out[f"x_{key}"] = s.x
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "iters cold", float(sol.iter.mean()), "from x_opt", float(out["iter_x_opt"].mean()), "from lam_opt",
              float(out["iter_lam_opt"].mean()), "from x_out", float(out["iter_x_out"].mean()),
              "bits set from x_in/x_out:", int((out["sense_x_in"] & 1).sum()), int((out["sense_x_out"] & 1).sum()))


if __name__ == "__main__":
    main()

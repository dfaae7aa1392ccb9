"""Golden vectors for branch and bound over binary constraints (reference src/bnb.c), produced by the UNMODIFIED
reference (oracle/_ref/libdaqp_ref.so: daqp_quadprog with sense & 16). Run in the build container:

    python tests/golden/make_golden_bnb.py

Inputs: a numpy port of the reference's MIQP generator `generate_test_MIQP(n, m, ms, nb)` (interfaces/daqp-julia/test/
utils.jl:145-166: PD Hessian, the first nb simple bounds binary in [0, 1], a cardinality row sum(x[1:nb]) <= nb/2 so that
the relaxation is fractional and the search has to branch and backtrack), the BnB set-up of core_tests.jl:130-148, and the
literals of core_tests.jl:150-178 (x = [0, 1, 1]; zero-dual endpoints: one node)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from daqp_b200.problems import QPBatch, generate_miqp  # noqa: E402
from oracle import harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    cases = {"bnb_n10_m20_ms8_nb6": generate_miqp(12, 10, 20, 8, 6, seed=701),
             "bnb_n20_m40_ms14_nb10": generate_miqp(8, 20, 40, 14, 10, seed=702),
             "bnb_n8_m12_ms8_nb8": generate_miqp(12, 8, 12, 8, 8, seed=703)}
    # literals of the reference's own tests
    H = np.array([[1, 0.5, 0], [0.5, 1, 0.5], [0, 0.5, 1.0]])[None]
    cases["bnb_lit_x011"] = QPBatch(3, 5, 3, H, np.array([[1.0, 0, 0]]), np.array([[[1.0, 2, 3], [1, 1, 0]]]),
                                    np.array([[1.0, 1, 1, 1e30, 1e30]]), np.array([[0.0, 0, 0, 4, 1]]),
                                    np.array([[16, 16, 16, 0, 0]], np.int32))
    nd = 8
    cases["bnb_lit_zero_dual_simple"] = QPBatch(nd, nd, nd, np.eye(nd)[None], np.zeros((1, nd)), np.zeros((1, 0, nd)),
                                                np.ones((1, nd)), np.zeros((1, nd)), np.full((1, nd), 16, np.int32))
    cases["bnb_lit_zero_dual_general"] = QPBatch(nd, nd, 0, np.eye(nd)[None], np.zeros((1, nd)), np.eye(nd)[None].copy(),
                                                 np.ones((1, nd)), np.zeros((1, nd)), np.full((1, nd), 16, np.int32))
    for name, b in cases.items():
        sol = ref.solve(b, use_sense=True)
        nodes = getattr(sol, "nodes", None)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), n=b.n, m=b.m, ms=b.ms, H=b.H, f=b.f, A=b.A, bupper=b.bupper,
                            blower=b.blower, sense=b.sense, x=sol.x, fval=sol.fval, exitflag=sol.exitflag, iter=sol.iter)
        print(f"{name:28s} N={b.N} flags={dict(zip(*[v.tolist() for v in np.unique(sol.exitflag, return_counts=True)]))} "
              f"iters={sol.iter.tolist()[:8]}")


if __name__ == "__main__":
    main()

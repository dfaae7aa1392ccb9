"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libdaqp_ref.so, compiled from
/root/reference by oracle/Makefile with the reference's default flags) on seeded inputs.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures hold inputs AND the reference's outputs (x, lam, fval, exitflag, iter, final working set), so the
GPU box -- where /root/reference does not exist -- can check both the oracle and the CUDA path against them.
Known-answer cases restate the literals of the reference's own tests (file:line in each entry).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from daqp_b200.problems import QPBatch, generate_g0, generate_g1  # noqa: E402
from oracle import harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, b: QPBatch, ref, use_sense, note):
    sol = ref.solve(b, use_sense=use_sense, want_ws=True)
    wsmax = b.n + 1
    ws = -np.ones((b.N, wsmax), np.int32)
    for p, w in enumerate(sol.ws):
        ws[p, :len(w)] = w
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=b.n, m=b.m, ms=b.ms, H=b.H, f=b.f, A=b.A,
                        bupper=b.bupper, blower=b.blower, sense=b.sense, use_sense=use_sense,
                        x=sol.x, lam=sol.lam, fval=sol.fval, exitflag=sol.exitflag, iter=sol.iter, ws=ws,
                        n_active=np.array([len(w) for w in sol.ws], np.int32),
                        xref=b.xref if b.xref is not None else np.zeros((0,)), note=note)
    print(f"{name:24s} N={b.N} flags={dict(zip(*np.unique(sol.exitflag, return_counts=True)))} "
          f"iters={sol.iter.tolist()[:6]}")


def literal(H, f, A, bu, bl, sense=None):
    H = np.array(H, float)[None]; f = np.array(f, float)[None]
    A = np.array(A, float).reshape(-1, H.shape[1])[None]
    bu = np.array(bu, float)[None]; bl = np.array(bl, float)[None]
    m = bu.shape[1]
    s = np.zeros((1, m), np.int32) if sense is None else np.array(sense, np.int32)[None]
    return QPBatch(H.shape[1], m, m - A.shape[1], H, f, A, bu, bl, s)


def main():
    harness.build(ref=True)
    ref = harness.RefLib("libdaqp_ref.so")
    oracle = harness.OracleLib()

    save("g1_n10_m20", generate_g1(8, 10, 20, 0, 8, seed=101), ref, False, "config C1 shape (G1)")
    save("g1_n20_m60", generate_g1(6, 20, 60, 0, 16, seed=102), ref, False, "config C2 shape (G1)")
    save("g1_n50_m150", generate_g1(3, 50, 150, 0, 40, seed=103), ref, False, "config C3 shape (G1)")
    save("g0_n20_m60", generate_g0(4, 20, 60, seed=104), ref, False, "probe distribution G0")
    save("g1_n12_m30_ms6", generate_g1(6, 12, 30, 6, 9, seed=105), ref, False, "simple bounds")
    save("g1_n16_m48_ms16", generate_g1(4, 16, 48, 16, 12, seed=106), ref, False, "box on every variable (C4 shape, small)")
    save("g1_kappa1e8", generate_g1(4, 12, 36, 4, 9, kappa=1e8, seed=107), ref, False, "ill-conditioned H")

    b = generate_g1(6, 12, 36, 4, 9, seed=108)  # warm start from the optimal active set: core_tests.jl:520-543
    o = oracle.solve(b)
    b.sense[o.lam > 1e-12] = 1
    b.sense[o.lam < -1e-12] = 3
    save("warm_exact", b, ref, True, "sense pre-set to the optimal active set => iter == 1 (core_tests.jl:520-543)")

    rng = np.random.default_rng(7)
    b = generate_g1(6, 12, 36, 4, 9, seed=109)
    b.sense[:] = np.where(rng.random(b.sense.shape) < 0.5, rng.choice([1, 3], b.sense.shape), 0)
    save("warm_wrong", b, ref, True, "over-determined / wrong warm start (auxiliary.c:461-475)")

    b = generate_g1(6, 12, 36, 0, 9, seed=110)
    for p in range(b.N):
        act = np.nonzero(b.active_ref[p])[0][:3]
        b.sense[p, act] = 5
        for i in act:
            if b.active_ref[p, i] > 0: b.blower[p, i] = b.bupper[p, i]
            else: b.bupper[p, i] = b.blower[p, i]
    save("equalities", b, ref, True, "equality rows (sense=5, bl==bu)")

    b = generate_g1(4, 10, 30, 0, 8, seed=111)
    b.A[:, 1] = b.A[:, 0]; b.bupper[:, 1] = b.blower[:, 0] - 1.0; b.blower[:, 1] = b.blower[:, 0] - 2.0
    save("infeasible", b, ref, False, "contradictory parallel rows => exitflag -1")

    b = generate_g1(4, 10, 30, 0, 8, seed=112)
    b.blower[:, 3] = b.bupper[:, 3] + 1
    save("trivially_infeasible", b, ref, False, "blower > bupper => -1 from setup (core_tests.jl:412-443)")

    b = generate_g1(4, 10, 30, 0, 8, seed=113)
    b.A[:, 15:30] = b.A[:, 0:15]; b.bupper[:, 15:30] = b.bupper[:, 0:15]; b.blower[:, 15:30] = b.blower[:, 0:15]
    save("duplicate_rows", b, ref, False, "linearly dependent rows => singular LDL' steps")

    save("unconstrained", generate_g1(4, 10, 30, 0, 0, seed=114), ref, False,
         "unconstrained optimum feasible => iter 1 (core_tests.jl:825-839)")

    # literals of the reference's own tests
    save("lit_python_demo", literal(np.eye(2), [1, 1], [[1, 1], [1, -1]], [1, 2, 3, 4], [-1, -2, -3, -4]), ref, True,
         "example_test.py:17-26 => exitflag 1")
    save("lit_model_qp", literal(np.eye(2), [2, 2], np.eye(2), [1, 1], [-1, -1]), ref, True,
         "example_test.py:175-193 => x = [-1,-1]")
    save("lit_model_qp_flipped", literal(np.eye(2), [-2, -2], np.eye(2), [1, 1], [-1, -1]), ref, True,
         "example_test.py:206-223 => x = [1,1]")
    save("lit_model_qp_half", literal(np.eye(2), [2, 2], np.eye(2), [0.5, 0.5], [-0.5, -0.5]), ref, True,
         "example_test.py:225-237 => x = [-0.5,-0.5]")
    save("lit_eigen_basic", literal(np.eye(2), [1, 1], [[1, 2], [1, -1]], [1, 2, 3, 4], [-1, -2, -3, -4]), ref, True,
         "00_basic_qp.cpp:7-28 => x = (-1,-1)")


if __name__ == "__main__":
    main()

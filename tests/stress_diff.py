"""Differential stress run (not collected by pytest): random shapes and seeds, GPU against the oracle (all host threads),
cold and warm-started; prints every batch whose exit flags, iteration counts or path counters differ.

    python tests/stress_diff.py [--batches 40] [--N 400] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import daqp_b200  # noqa: E402
from daqp_b200.problems import generate_g0, generate_g1  # noqa: E402
from oracle import harness  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=40); ap.add_argument("--N", type=int, default=400); ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    harness.build(ref=False)
    orc = harness.OracleLib()
    eng = daqp_b200.Engine()
    rng = np.random.default_rng(args.seed)
    bad = 0
    threads = os.cpu_count() or 1
    for k in range(args.batches):
        n = int(rng.integers(3, 128)); m = int(rng.integers(n + 1, 4 * n + 2)); ms = int(rng.integers(0, min(n, m) + 1)) if rng.random() < 0.5 else 0
        na = int(rng.integers(0, min(n, m) + 1))
        kappa = float(10 ** rng.uniform(0, 6))
        seed = int(rng.integers(1, 1 << 30))
        if rng.random() < 0.2 and ms == 0:
            b = generate_g0(args.N, n, m, seed=seed); kind = "G0"
        else:
            b = generate_g1(args.N, n, m, ms, na, kappa=kappa, seed=seed); kind = f"G1 kappa=1e{np.log10(kappa):.1f} nact={na}"
        for mode in ("cold", "warm"):
            if mode == "warm":
                b.sense[:] = 0
                b.sense[o.lam > 1e-12] = 1; b.sense[o.lam < -1e-12] = 3
                flip = rng.random(b.sense.shape) < 0.03  # a few wrong rows in / right rows out
                b.sense[flip] = np.where(b.sense[flip] != 0, 0, rng.choice([1, 3], int(flip.sum())))
            use_sense = mode == "warm"
            o = orc.solve_packed(b, nthreads=threads, use_sense=use_sense)
            r = eng.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, b.sense if use_sense else None, ms=b.ms, diag=True)
            df = np.nonzero(r.exitflag != o.exitflag)[0]; di = np.nonzero((r.iter != o.iter) & (r.exitflag == o.exitflag))[0]
            ok = o.exitflag > 0
            xerr = float(np.abs(r.x[ok] - o.x[ok]).max() / (1 + np.abs(o.x[ok]).max())) if ok.any() else 0.0
            tag = f"[{k:3d}] {mode} n={n} m={m} ms={ms} {kind} seed={seed}: iters {o.iter.mean():.1f}, flags {dict(zip(*[v.tolist() for v in np.unique(o.exitflag, return_counts=True)]))}, x err {xerr:.1e}"
            if df.size or di.size or xerr > 1e-7:
                bad += 1
                print("MISMATCH", tag, "| flag diffs", df[:8].tolist(), "iter diffs", di[:8].tolist(),
                      "gpu", r.iter[di[:4]].tolist(), "oracle", o.iter[di[:4]].tolist())
            else:
                print("ok      ", tag)
            if mode == "cold" and not (o.exitflag > 0).any():
                break
    print(f"{bad} mismatching batches of {args.batches} x 2")
    eng.close()


if __name__ == "__main__":
    main()

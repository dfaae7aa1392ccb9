"""Development check (run under gpurun): CUDA path vs the oracle on a few configurations."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import daqp_b200
from daqp_b200.problems import generate_g1, generate_g0
from oracle.harness import OracleLib

def compare(name, b, use_sense=None, **settings):
    orc = OracleLib()
    o = orc.solve_packed(b, use_sense=use_sense) if not settings else orc.solve(b, use_sense=use_sense)
    eng = daqp_b200.Engine()
    t0 = time.perf_counter()
    sense = b.sense if (use_sense or (use_sense is None and b.sense.any())) else None
    r = eng.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, sense, ms=b.ms, diag=True, **settings)
    dt = time.perf_counter() - t0
    flag_eq = np.array_equal(o.exitflag, r.exitflag)
    iter_eq = np.array_equal(o.iter, r.iter)
    good = (o.exitflag > 0) & (r.exitflag > 0)
    dx = np.abs(o.x[good] - r.x[good]).max() if good.any() else 0
    dl = np.abs(o.lam[good] - r.lam[good]).max() if good.any() else 0
    df = np.abs(o.fval[good] - r.fval[good]).max() if good.any() else 0
    cnt_eq = np.array_equal(o.counts, r.counts)
    print(f"{name:32s} N={b.N} flag_eq={flag_eq} iter_eq={iter_eq} counts_eq={cnt_eq} dx={dx:.2e} dlam={dl:.2e} dfval={df:.2e} "
          f"flags={dict(zip(*np.unique(r.exitflag, return_counts=True)))} t={dt:.3f}s stats={eng.stats()}", flush=True)
    if not (flag_eq and iter_eq):
        bad = np.nonzero((o.exitflag != r.exitflag) | (o.iter != r.iter))[0]
        print("   first mismatches:", [(int(i), int(o.exitflag[i]), int(r.exitflag[i]), int(o.iter[i]), int(r.iter[i])) for i in bad[:8]], "n_bad", len(bad))
    eng.close()

if __name__ == "__main__":
    compare("G1 n10 m20", generate_g1(64, 10, 20, 0, 8))
    compare("G1 n20 m60", generate_g1(1000, 20, 60, 0, 16))
    compare("G1 n50 m150", generate_g1(2000, 50, 150, 0, 40))
    compare("G0 n50 m150", generate_g0(1000, 50, 150))
    compare("G1 n20 m60 ms10", generate_g1(1000, 20, 60, 10, 16))
    compare("G1 n30 m80 ms30", generate_g1(500, 30, 80, 30, 20))
    compare("G1 n70 m200 ms7", generate_g1(300, 70, 200, 7, 50))
    compare("G1 n120 m400 ms120", generate_g1(200, 120, 400, 120, 96))
    compare("G1 n10 m40 nact10", generate_g1(500, 10, 40, 0, 10))
    compare("G1 kappa1e8", generate_g1(500, 20, 60, 5, 16, kappa=1e8))
    b = generate_g1(500, 10, 30, 0, 8); b.A[:, 15:30] = b.A[:, 0:15]; b.bupper[:, 15:30] = b.bupper[:, 0:15]; b.blower[:, 15:30] = b.blower[:, 0:15]
    compare("duplicate rows", b)
    b = generate_g1(500, 10, 30, 0, 8); b.A[:, 1] = b.A[:, 0]; b.bupper[:, 1] = b.blower[:, 0] - 1.0; b.blower[:, 1] = b.blower[:, 0] - 2.0
    compare("infeasible", b)
    b = generate_g1(200, 20, 60, 5, 16)
    o = OracleLib().solve(b)
    b.sense[o.lam > 1e-12] = 1; b.sense[o.lam < -1e-12] = 3
    compare("warm exact", b, use_sense=True)
    rng = np.random.default_rng(0)
    b.sense[:] = np.where(rng.random(b.sense.shape) < 0.6, rng.choice([1, 3], b.sense.shape), 0)
    compare("warm random 60%", b, use_sense=True)
    compare("iter_limit 5", generate_g1(50, 20, 60, 0, 16), iter_limit=5)
    compare("unconstrained", generate_g1(50, 10, 30, 0, 0))

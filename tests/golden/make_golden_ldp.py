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


def raw_cases():
    """daqp_ldp on hand-filled workspaces (no normalisation: the rows keep their norms, spread over two decades on
    purpose so that a normalising implementation would take another path), with simple bounds, infeasible polyhedra and a
    max-radius bound (settings.fval_bound, api.jl:441-445)."""
    out = {}
    for name, (N, n, m, ms, seed, fvb) in {"rawldp_n8_m24": (20, 8, 24, 0, 911, None), "rawldp_n12_m36_ms4": (20, 12, 36, 4, 912, None),
                                           "rawldp_n30_m90_radius": (12, 30, 90, 0, 913, 5.5), "rawldp_n70_m160": (6, 70, 160, 0, 914, None)}.items():
        b = make(N, n, m, ms, seed, shift=0.7, infeasible_every=5)
        rng = np.random.Generator(np.random.Philox(key=seed + 50))
        scale = 10.0 ** rng.uniform(-1, 1, (N, m - ms))
        b.A *= scale[:, :, None]; b.bupper[:, ms:] *= scale
        fin = b.blower[:, ms:] > -1e29
        b.blower[:, ms:][fin] = (b.blower[:, ms:] * scale)[fin]
        out[name] = (b, fvb)
    # rows the caller has switched off (sense = IMMUTABLE without ACTIVE: every scan skips them, daqp.c / auxiliary.c:110)
    b, _ = out["rawldp_n8_m24"]
    b2 = QPBatch(b.n, b.m, b.ms, b.H.copy(), None, b.A.copy(), b.bupper.copy(), b.blower.copy(), b.sense.copy())
    rng = np.random.Generator(np.random.Philox(key=931))
    b2.sense[rng.uniform(size=b2.sense.shape) < 0.25] = 4
    out["rawldp_n8_m24_switched_off"] = (b2, None)
    return out


def main_raw():
    import ctypes as C
    libs = [C.CDLL(os.path.join(harness.REF_DIR, nm)) for nm in ("libdaqp_ref.so", "libdaqp_ref_strict.so")]
    for name, (b, fvb) in raw_cases().items():
        rows = []
        for p in range(b.N):
            r = [harness.raw_ldp(L, b.A[p], b.bupper[p], b.blower[p], b.sense[p] if b.sense.any() else None, b.ms, fvb) for L in libs]
            if r[0]["exitflag"] == r[1]["exitflag"] and r[0]["iter"] == r[1]["iter"] and r[0]["ws"] == r[1]["ws"]:
                rows.append((p, r[0]))
        keep = [p for p, _ in rows]
        ws = np.full((len(rows), b.n + 1), -1, np.int32)
        for q, (_, r) in enumerate(rows):
            ws[q, :len(r["ws"])] = r["ws"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), n=b.n, m=b.m, ms=b.ms, A=b.A[keep], bupper=b.bupper[keep],
                            blower=b.blower[keep], sense=b.sense[keep], fval_bound=-1.0 if fvb is None else fvb,
                            exitflag=np.array([r["exitflag"] for _, r in rows], np.int32),
                            iter=np.array([r["iter"] for _, r in rows], np.int32), u=np.array([r["u"] for _, r in rows]),
                            fval=np.array([r["fval"] for _, r in rows]), ws=ws,
                            n_active=np.array([len(r["ws"]) for _, r in rows], np.int32))
        fl = np.array([r["exitflag"] for _, r in rows])
        print(f"{name:24s} N={len(rows)} (dropped {b.N - len(rows)}) flags={dict(zip(*[v.tolist() for v in np.unique(fl, return_counts=True)]))}")


if __name__ == "__main__":
    main()
    main_raw()

"""Shared helpers for the parity tests.

fp64 parity bar (SURVEY.md §8c / BASELINE.md §3.5): exit flag equal, iteration count equal, final working set equal
(index + side), |x - x_ref|inf <= 1e-9 (1+|x_ref|inf), |fval - fval_ref| <= 1e-9 (1+|fval_ref|),
|lam - lam_ref|inf <= 1e-7 (1+|lam_ref|inf).
"""
import glob
import os

import numpy as np

from daqp_b200.problems import QPBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
X_TOL, F_TOL, L_TOL = 1e-9, 1e-9, 1e-7


def golden_names():
    """Single-solve fixtures (the wsseq_* files hold workspace sequences: see test_workspace_sequence_matches_reference)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("wsseq_", "wsshared_", "minrep_", "warmstart_", "rare_", "bnb_", "ldp_", "rawldp_"))]


def bnb_golden_names():
    """MIQPs with the reference's own branch-and-bound output (tests/golden/make_golden_bnb.py)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "bnb_*.npz")))


def rawldp_golden_names():
    """daqp_ldp on hand-filled workspaces (no normalisation), the reference's own output (make_golden_ldp.py)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "rawldp_*.npz")))


def ldp_golden_names():
    """Pure LDP inputs (H == NULL, f == NULL) with the reference's own output (tests/golden/make_golden_ldp.py)."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "ldp_*.npz")))


def load_ldp_golden(name):
    """-> (batch whose H is only an identity placeholder and whose f is None, file)"""
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, N = int(d["n"]), d["bupper"].shape[0]
    b = QPBatch(n, int(d["m"]), int(d["ms"]), np.broadcast_to(np.eye(n), (N, n, n)).copy(), None, d["A"], d["bupper"],
                d["blower"], d["sense"].astype(np.int32))
    return b, d


def rare_golden_names():
    """Inputs that drive the rare control paths of daqp_ldp (tests/golden/make_golden_rare.py): outputs of the reference
    built without reassociation, non-default settings, and the oracle's path counters."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "rare_*.npz")))


def rare_settings(d):
    return {str(k): (int(v) if str(k) in ("cycle_tol", "iter_limit") else float(v))
            for k, v in zip(d["settings_keys"], d["settings_vals"])}


# which of the eight path counters (scan, add, remove, csp, pivot, refine, refactor, cycle repair) a fixture must hit
RARE_MUST_HIT = {"rare_eqpairs_n20": (4, 5, 6, 7), "rare_eqpairs_n12_ms4": (4, 5, 6, 7), "rare_eqpairs_n50": (4, 5, 6, 7),
                 "rare_eqpairs_n70": (4, 5, 6), "rare_parallel_1e-3": (4, 5), "rare_cycle_guard": (7,),
                 "rare_cycle_exit": (7,)}


def minrep_golden_names():
    """Polyhedra with the reference's own daqp_minrep output (tests/golden/make_golden_minrep.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "minrep_n*.npz")))
    return names


def load_minrep_special():
    d = np.load(os.path.join(GOLDEN_DIR, "minrep_special.npz"))
    return {str(k): (d[str(k) + "_A"], d[str(k) + "_b"], d[str(k) + "_red"]) for k in d["names"]}


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    b = QPBatch(int(d["n"]), int(d["m"]), int(d["ms"]), d["H"], d["f"], d["A"], d["bupper"], d["blower"],
                d["sense"].astype(np.int32))
    return b, d


def assert_parity(ref_x, ref_lam, ref_fval, ref_flag, ref_iter, x, lam, fval, flag, it, what="", x_tol=None,
                  f_tol=None):
    np.testing.assert_array_equal(flag, ref_flag, err_msg=f"{what}: exit flags differ")
    np.testing.assert_array_equal(it, ref_iter, err_msg=f"{what}: iteration counts differ")
    ok = ref_flag > 0
    if not ok.any():
        return
    xs = 1 + np.abs(ref_x[ok]).max(axis=1, keepdims=True)
    lt = L_TOL if x_tol is None else max(L_TOL, 1e3 * x_tol)
    ft = F_TOL if x_tol is None else max(F_TOL, x_tol)
    ft = ft if f_tol is None else f_tol
    x_tol = X_TOL if x_tol is None else x_tol
    assert (np.abs(x[ok] - ref_x[ok]) <= x_tol * xs).all(), f"{what}: x off by {np.abs(x[ok] - ref_x[ok]).max():.3e}"
    ls = 1 + np.abs(ref_lam[ok]).max(axis=1, keepdims=True)
    assert (np.abs(lam[ok] - ref_lam[ok]) <= lt * ls).all(), \
        f"{what}: lam off by {np.abs(lam[ok] - ref_lam[ok]).max():.3e}"
    assert (np.abs(fval[ok] - ref_fval[ok]) <= ft * (1 + np.abs(ref_fval[ok]))).all(), \
        f"{what}: fval off by {np.abs(fval[ok] - ref_fval[ok]).max():.3e}"


def ws_sets(ws, n_active):
    return [sorted(int(i) for i in ws[p][: n_active[p]]) for p in range(len(n_active))]


def kkt_residuals(b: QPBatch, x, lam):
    """Solver-independent check in the spirit of the reference's Maros-Meszaros runner
    (.github/benchmarks/maros_meszaros_runner.c:79-130): stationarity, primal feasibility, complementarity."""
    N, n, m, ms = b.N, b.n, b.m, b.ms
    Afull = np.zeros((N, m, n))
    Afull[:, :ms, :] = np.eye(n)[None, :ms, :]
    Afull[:, ms:, :] = b.A
    Ax = np.einsum("bmn,bn->bm", Afull, x)
    stat = np.einsum("bij,bj->bi", b.H, x) + b.f + np.einsum("bmn,bm->bn", Afull, lam)
    pfeas = np.maximum(np.maximum(Ax - b.bupper, b.blower - Ax), 0)
    comp = np.minimum(np.abs(lam), np.minimum(np.abs(Ax - b.bupper), np.abs(Ax - b.blower)))
    return np.abs(stat).max(axis=1), pfeas.max(axis=1), comp.max(axis=1)

"""oracle/harness.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for (a) the reference compiled from /root/reference into oracle/_ref/*.so and (b) this repo's scalar
restatement oracle/libdaqp_oracle*.so. Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
Struct layouts mirror the reference ABI (include/types.h:32-74,187-264, include/api.h:15-27).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

UPDATE_UNCONSTRAINED = 64
UPDATE_ELIMINATE = 128


def _structs(real):
    P = C.POINTER

    class Problem(C.Structure):
        _fields_ = [("n", C.c_int), ("m", C.c_int), ("ms", C.c_int), ("H", P(real)), ("f", P(real)), ("A", P(real)),
                    ("bupper", P(real)), ("blower", P(real)), ("sense", P(C.c_int)), ("break_points", P(C.c_int)),
                    ("nh", C.c_int), ("problem_type", C.c_int)]

    class Settings(C.Structure):
        _fields_ = [("primal_tol", real), ("dual_tol", real), ("zero_tol", real), ("pivot_tol", real),
                    ("progress_tol", real), ("cycle_tol", C.c_int), ("iter_limit", C.c_int), ("fval_bound", real),
                    ("eps_prox", real), ("eta_prox", real), ("rho_soft", real), ("rel_subopt", real),
                    ("abs_subopt", real), ("sing_tol", real), ("refactor_tol", real), ("time_limit", real)]

    class Result(C.Structure):
        _fields_ = [("x", P(real)), ("lam", P(real)), ("fval", real), ("soft_slack", real), ("exitflag", C.c_int),
                    ("iter", C.c_int), ("nodes", C.c_int), ("solve_time", real), ("setup_time", real)]

    class Workspace(C.Structure):  # include/types.h:187-264 (SOFT_WEIGHTS off)
        _fields_ = [("qp", C.c_void_p), ("n", C.c_int), ("m", C.c_int), ("ms", C.c_int), ("M", P(real)),
                    ("dupper", P(real)), ("dlower", P(real)), ("Rinv", P(real)), ("v", P(real)),
                    ("sense", P(C.c_int)), ("scaling", P(real)), ("RinvD", P(real)), ("x", P(real)),
                    ("xold", P(real)), ("lam", P(real)), ("lam_star", P(real)), ("u", P(real)), ("fval", real),
                    ("L", P(real)), ("D", P(real)), ("xldl", P(real)), ("zldl", P(real)), ("reuse_ind", C.c_int),
                    ("WS", P(C.c_int)), ("n_active", C.c_int), ("iterations", C.c_int), ("sing_ind", C.c_int),
                    ("prox_mask", P(C.c_int)), ("n_prox", C.c_int), ("soft_slack", real), ("settings", C.c_void_p),
                    ("bnb", C.c_void_p), ("nh", C.c_int), ("break_points", P(C.c_int)), ("avi", C.c_void_p),
                    ("eq", C.c_void_p), ("timer", C.c_void_p), ("Mu", P(real))]

    class Trace(C.Structure):
        _fields_ = [("n_active", C.c_int), ("ws", P(C.c_int)), ("sense_out", P(C.c_int)), ("n_scan", C.c_int),
                    ("n_add", C.c_int), ("n_remove", C.c_int), ("n_csp", C.c_int), ("n_pivot", C.c_int),
                    ("n_refine", C.c_int), ("n_refactor", C.c_int), ("n_cycle", C.c_int),
                    ("oplog", P(C.c_int)), ("oplog_cap", C.c_int), ("n_log", C.c_int)]

    return Problem, Settings, Result, Workspace, Trace


_F64 = _structs(C.c_double)
_F32 = _structs(C.c_float)


def build(ref: bool = True) -> None:
    """Compile the oracle restatement, and the reference into oracle/_ref when /root/reference is present."""
    have_src = ref and os.path.isdir("/root/reference/src")
    targets = ["oracle"] + (["ref"] if have_src else [])
    # the reference's own Cython interface linked against the product library (tests/test_reference_binding.py)
    if have_src and os.path.exists(os.path.join(os.path.dirname(HERE), "daqp_b200", "libdaqp_b200.so")):
        targets.append("pybind")
    subprocess.run(["make", "-C", HERE, "CC=gcc", f"PY={sys.executable}"] + targets, check=True, capture_output=True)


def have_ref(name: str = "libdaqp_ref.so") -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


@dataclass
class Solution:
    x: np.ndarray
    lam: np.ndarray
    fval: np.ndarray
    exitflag: np.ndarray
    iter: np.ndarray
    ws: list | None = None        # per problem: working-set indices in factor order
    sense: np.ndarray | None = None  # [N, m] final sense bits
    counts: np.ndarray | None = None  # [N,8] scan, add, remove, csp, pivot, refine, refactor, cycle repair (oracle only)
    seconds: float = 0.0
    soft_slack: np.ndarray | None = None
    oplog: list | None = None     # per problem: [k,2] decision log (OracleLib.solve(log_cap=...))


def default_settings(dtype=np.float64, **over):
    S = (_F64 if dtype == np.float64 else _F32)[1]
    s = S(1e-6, 1e-12, 1e-11, 1e-6, 1e-14, 10, 10000, 1e30, -1e-6, -1.0, 1e-6, 0, 0, 3.7e-11, 1e-9, 0)
    for k, v in over.items():
        setattr(s, k, v)
    return s


def _ptr(a, real):
    return None if a is None else a.ctypes.data_as(C.POINTER(real))


class RefLib:
    """The unmodified reference (daqp_quadprog & the workspace API), loaded from oracle/_ref."""

    def __init__(self, name: str = "libdaqp_ref.so"):
        self.single = "f32" in name
        self.real = C.c_float if self.single else C.c_double
        self.dtype = np.float32 if self.single else np.float64
        self.Problem, self.Settings, self.Result, self.Workspace, _ = _F32 if self.single else _F64
        self.lib = C.CDLL(os.path.join(REF_DIR, name))
        self.lib.daqp_quadprog.restype = None
        self.lib.setup_daqp_main.restype = C.c_int
        self.lib.daqp_solve.restype = None

    def _problem(self, b, p, sense, null_H=False):
        real = self.real
        mA = b.m - b.ms
        return self.Problem(b.n, b.m, b.ms, None if null_H else _ptr(b.H[p], real), _ptr(b.f[p], real) if b.f is not None else None,
                            _ptr(b.A[p], real) if mA > 0 else None, _ptr(b.bupper[p], real), _ptr(b.blower[p], real),
                            sense.ctypes.data_as(C.POINTER(C.c_int)) if sense is not None else None, None, 0, 0)

    def solve(self, b, settings=None, use_sense: bool | None = None, want_ws: bool = False, null_H: bool = False) -> Solution:
        """One daqp_quadprog call per problem (api.c:62-79). want_ws drives the same calls through
        setup_daqp_main + daqp_solve so that the final working set can be read from the workspace. null_H passes
        H = NULL (with b.f = None: the LDP min |x|^2 over the constraints; b.H is then only a placeholder)."""
        import time
        N, n, m = b.N, b.n, b.m
        b = b.astype(self.dtype)
        x = np.zeros((N, n), self.dtype); lam = np.zeros((N, m), self.dtype)
        fval = np.zeros(N, self.dtype); flag = np.zeros(N, np.int32); it = np.zeros(N, np.int32)
        ws = [] if want_ws else None
        slack = np.zeros(N, self.dtype)
        sense_out = np.zeros((N, m), np.int32) if want_ws else None
        if use_sense is None:
            use_sense = bool(np.any(b.sense))
        sp = C.byref(settings) if settings is not None else None
        t0 = time.perf_counter()
        for p in range(N):
            sense = b.sense[p].copy() if use_sense else None
            qp = self._problem(b, p, sense, null_H)
            res = self.Result(_ptr(x[p], self.real), _ptr(lam[p], self.real), 0, 0, 0, 0, 0, 0, 0)
            if not want_ws:
                self.lib.daqp_quadprog(C.byref(res), C.byref(qp), sp)
            else:
                work = self.Workspace()
                work.settings = C.cast(sp, C.c_void_p) if sp is not None else None
                st = self.real(0)
                rc = self.lib.setup_daqp_main(C.byref(qp), C.byref(work), C.byref(st),
                                              UPDATE_UNCONSTRAINED | UPDATE_ELIMINATE)
                res.exitflag = rc
                if rc >= 0:
                    self.lib.daqp_solve(C.byref(res), C.byref(work))
                    ws.append([work.WS[i] for i in range(work.n_active)])
                    sense_out[p] = [work.sense[i] for i in range(m)]
                    if sp is not None:
                        work.settings = None
                    self.lib.free_daqp_workspace(C.byref(work))
                    self.lib.free_daqp_ldp(C.byref(work))
                else:
                    ws.append([])
            fval[p], flag[p], it[p] = res.fval, res.exitflag, res.iter
            slack[p] = res.soft_slack
        return Solution(x, lam, fval, flag, it, ws, sense_out, None, time.perf_counter() - t0, slack)


def raw_ldp(lib, A, bupper, blower, sense=None, ms=0, fval_bound=None):
    """daqp_ldp on a HAND-FILLED workspace, the way interfaces/daqp-julia/src/api.jl:428-459 does it: calloc a
    DAQPWorkspace, allocate_daqp_workspace(n, 0) + allocate_daqp_settings, store the caller's M = A (row-major [m-ms][n],
    NOT normalised), dupper, dlower, sense, m, ms in the struct, call daqp_ldp, read the struct back. `lib` is any CDLL
    that exports those symbols (the reference, or the library under test). Returns a dict."""
    _, Settings, _, Workspace, _ = _F64
    A = np.ascontiguousarray(A, np.float64); bu = np.ascontiguousarray(bupper, np.float64)
    bl = np.ascontiguousarray(blower, np.float64)
    m = bu.shape[0]; n = A.shape[1]
    se = np.zeros(m, np.int32) if sense is None else np.ascontiguousarray(sense, np.int32).copy()
    lib.daqp_ldp.restype = C.c_int
    for f in ("allocate_daqp_workspace", "allocate_daqp_settings", "free_daqp_workspace"):
        getattr(lib, f).restype = None
    work = Workspace()
    lib.allocate_daqp_workspace(C.byref(work), C.c_int(n), C.c_int(0))
    lib.allocate_daqp_settings(C.byref(work))
    if fval_bound is not None:
        C.cast(work.settings, C.POINTER(Settings)).contents.fval_bound = fval_bound
    work.M = _ptr(A, C.c_double) if A.size else None
    work.dupper = _ptr(bu, C.c_double); work.dlower = _ptr(bl, C.c_double)
    work.sense = se.ctypes.data_as(C.POINTER(C.c_int))
    work.m = m; work.ms = ms
    flag = lib.daqp_ldp(C.byref(work))
    k = work.n_active
    out = {"exitflag": int(flag), "iter": int(work.iterations), "u": np.array([work.u[i] for i in range(n)]),
           "ws": [int(work.WS[i]) for i in range(k)], "lam_star": np.array([work.lam_star[i] for i in range(k)]),
           "fval": float(work.fval), "sense": se.copy()}
    lib.free_daqp_workspace(C.byref(work))
    return out


def ref_solve_sequence(reflib, b, steps, settings=None, use_sense=None):
    """The reference's workspace flow on every problem of the batch: setup_daqp() + daqp_solve(), then per step
    daqp_update_ldp(DAQP_UPDATE_v + DAQP_UPDATE_d) with new (f, bupper, blower) + daqp_solve() on the KEPT workspace
    (src/api.c:88-160,214-260, src/utils.c:58-221). steps = list of (f[N,n], bupper[N,m], blower[N,m]).
    Returns one Solution per solve (len(steps) + 1)."""
    self = reflib
    N, n, m = b.N, b.n, b.m
    nsol = len(steps) + 1
    outs = [dict(x=np.zeros((N, n)), lam=np.zeros((N, m)), fval=np.zeros(N), flag=np.zeros(N, np.int32),
                 it=np.zeros(N, np.int32), ws=[], slack=np.zeros(N)) for _ in range(nsol)]
    if use_sense is None:
        use_sense = bool(np.any(b.sense))
    sp = C.byref(settings) if settings is not None else None
    self.lib.daqp_update_ldp.restype = C.c_int
    for p in range(N):
        f = b.f[p].copy(); bu = b.bupper[p].copy(); bl = b.blower[p].copy()
        sense = b.sense[p].copy() if use_sense else None
        mA = m - b.ms
        qp = self.Problem(n, m, b.ms, _ptr(b.H[p], self.real), _ptr(f, self.real),
                          _ptr(b.A[p], self.real) if mA > 0 else None, _ptr(bu, self.real), _ptr(bl, self.real),
                          sense.ctypes.data_as(C.POINTER(C.c_int)) if sense is not None else None, None, 0, 0)
        work = self.Workspace()
        work.settings = C.cast(sp, C.c_void_p) if sp is not None else None
        st = self.real(0)
        rc = self.lib.setup_daqp_main(C.byref(qp), C.byref(work), C.byref(st), 0)
        for k in range(nsol):
            o = outs[k]
            if k > 0 and rc >= 0:
                f[:] = steps[k - 1][0][p]; bu[:] = steps[k - 1][1][p]; bl[:] = steps[k - 1][2][p]
                rc = self.lib.daqp_update_ldp(4 + 8, C.byref(work), C.byref(qp))
                if rc >= 0:
                    rc = 1
            res = self.Result(_ptr(o["x"][p], self.real), _ptr(o["lam"][p], self.real), 0, 0, 0, 0, 0, 0, 0)
            if rc >= 0:
                self.lib.daqp_solve(C.byref(res), C.byref(work))
                o["ws"].append([work.WS[i] for i in range(work.n_active)])
            else:
                res.exitflag = rc
                o["ws"].append([])
            o["fval"][p], o["flag"][p], o["it"][p], o["slack"][p] = res.fval, res.exitflag, res.iter, res.soft_slack
            if rc < 0 and k == 0:
                break
            if rc < 0:
                rc = 1  # an update that failed (infeasible bounds) does not destroy the workspace
        if work.n > 0 or True:
            try:
                if sp is not None:
                    work.settings = None
                if rc >= 0 or nsol > 1:
                    self.lib.free_daqp_workspace(C.byref(work))
                    self.lib.free_daqp_ldp(C.byref(work))
            except Exception:
                pass
    return [Solution(o["x"], o["lam"], o["fval"], o["flag"], o["it"], o["ws"], None, None, 0.0, o["slack"]) for o in outs]


class RefDriver:
    """pthread loop around the unmodified reference's daqp_quadprog (oracle/ref_driver.c -> oracle/_ref)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_driver.so"))
        self.lib.ref_solve_packed.restype = C.c_double
        self.dtype = np.float64

    def solve_packed(self, b, settings=None, nthreads: int = 1, use_sense: bool | None = None) -> "Solution":
        N, n, m = b.N, b.n, b.m
        real = C.c_double
        x = np.zeros((N, n)); lam = np.zeros((N, m)); fval = np.zeros(N)
        flag = np.zeros(N, np.int32); it = np.zeros(N, np.int32)
        if use_sense is None:
            use_sense = bool(np.any(b.sense))
        sp = C.byref(settings) if settings is not None else None
        secs = self.lib.ref_solve_packed(
            N, n, m, b.ms, _ptr(b.H, real), _ptr(b.f, real), _ptr(b.A, real), _ptr(b.bupper, real),
            _ptr(b.blower, real), b.sense.ctypes.data_as(C.POINTER(C.c_int)) if use_sense else None, sp,
            _ptr(x, real), _ptr(lam, real), _ptr(fval, real), flag.ctypes.data_as(C.POINTER(C.c_int)),
            it.ctypes.data_as(C.POINTER(C.c_int)), nthreads)
        return Solution(x, lam, fval, flag, it, None, None, None, secs)


class OracleLib:
    """This repo's scalar restatement (oracle/daqp_oracle.c)."""

    def __init__(self, single: bool = False, variant: str | None = None):
        """variant="gpuarith": the restatement compiled with FMA contraction and the kernels' reciprocal form of the C1
        quotients -- a voter in the path-stability filters of the fixture generators, nothing else."""
        self.single = single
        self.real = C.c_float if single else C.c_double
        self.dtype = np.float32 if single else np.float64
        self.Problem, self.Settings, self.Result, _, self.Trace = _F32 if single else _F64
        name = "libdaqp_oracle_f32.so" if single else ("libdaqp_oracle_" + variant + ".so" if variant else "libdaqp_oracle.so")
        self.lib = C.CDLL(os.path.join(HERE, name))
        self.lib.orc_quadprog.restype = None
        self.lib.orc_solve_packed.restype = C.c_double

    def solve(self, b, settings=None, use_sense: bool | None = None, log_cap: int = 0, null_H: bool = False) -> Solution:
        """log_cap > 0 also records up to that many (code, value) decisions per problem in Solution.oplog: 1 add
        (2*constraint + lower), 2 remove (constraint), 3 refactor, 4 refine, 5 cycle repair, 7 exit (flag)."""
        import time
        N, n, m = b.N, b.n, b.m
        b = b.astype(self.dtype)
        real = self.real
        x = np.zeros((N, n), self.dtype); lam = np.zeros((N, m), self.dtype)
        fval = np.zeros(N, self.dtype); flag = np.zeros(N, np.int32); it = np.zeros(N, np.int32)
        counts = np.zeros((N, 8), np.int32)
        slack = np.zeros(N, self.dtype)
        sense_out = np.zeros((N, m), np.int32)
        ws = []
        logs = []
        logbuf = np.zeros((max(log_cap, 1), 2), np.int32)
        wsbuf = np.zeros(n + m + 2, np.int32)
        if use_sense is None:
            use_sense = bool(np.any(b.sense))
        sp = C.byref(settings) if settings is not None else None
        mA = m - b.ms
        t0 = time.perf_counter()
        for p in range(N):
            sense = b.sense[p].copy() if use_sense else None
            qp = self.Problem(n, m, b.ms, None if null_H else _ptr(b.H[p], real), _ptr(b.f[p], real) if b.f is not None else None,
                              _ptr(b.A[p], real) if mA > 0 else None, _ptr(b.bupper[p], real),
                              _ptr(b.blower[p], real),
                              sense.ctypes.data_as(C.POINTER(C.c_int)) if sense is not None else None, None, 0, 0)
            res = self.Result(_ptr(x[p], real), _ptr(lam[p], real), 0, 0, 0, 0, 0, 0, 0)
            tr = self.Trace(0, wsbuf.ctypes.data_as(C.POINTER(C.c_int)),
                            sense_out[p].ctypes.data_as(C.POINTER(C.c_int)), 0, 0, 0, 0, 0, 0, 0, 0,
                            logbuf.ctypes.data_as(C.POINTER(C.c_int)) if log_cap else None, log_cap, 0)
            self.lib.orc_quadprog(C.byref(res), C.byref(qp), sp, C.byref(tr))
            if log_cap:
                logs.append(logbuf[:min(tr.n_log, log_cap)].copy())
            fval[p], flag[p], it[p] = res.fval, res.exitflag, res.iter
            slack[p] = res.soft_slack
            counts[p] = (tr.n_scan, tr.n_add, tr.n_remove, tr.n_csp, tr.n_pivot, tr.n_refine, tr.n_refactor, tr.n_cycle)
            ws.append(wsbuf[:tr.n_active].tolist())
        return Solution(x, lam, fval, flag, it, ws, sense_out, counts, time.perf_counter() - t0, slack,
                        logs if log_cap else None)

    def solve_packed(self, b, settings=None, nthreads: int = 1, use_sense: bool | None = None) -> Solution:
        """Whole batch inside C (no Python in the timed loop); returns wall seconds measured in C."""
        N, n, m = b.N, b.n, b.m
        b = b.astype(self.dtype)
        real = self.real
        x = np.zeros((N, n), self.dtype); lam = np.zeros((N, m), self.dtype)
        fval = np.zeros(N, self.dtype); flag = np.zeros(N, np.int32); it = np.zeros(N, np.int32)
        counts = np.zeros((N, 8), np.int32)
        if use_sense is None:
            use_sense = bool(np.any(b.sense))
        sp = C.byref(settings) if settings is not None else None
        secs = self.lib.orc_solve_packed(
            N, n, m, b.ms, _ptr(b.H, real), _ptr(b.f, real), _ptr(b.A, real), _ptr(b.bupper, real),
            _ptr(b.blower, real), b.sense.ctypes.data_as(C.POINTER(C.c_int)) if use_sense else None, sp,
            _ptr(x, real), _ptr(lam, real), _ptr(fval, real), flag.ctypes.data_as(C.POINTER(C.c_int)),
            it.ctypes.data_as(C.POINTER(C.c_int)), counts.ctypes.data_as(C.POINTER(C.c_int)), nthreads)
        return Solution(x, lam, fval, flag, it, None, None, counts, secs)

    # ---- minimal representation (oracle restatement of daqp_minrep, fp64 only) --------------------------------
    def minrep(self, A, b) -> np.ndarray:
        """Reference order (one probe after the other): orc_minrep. A[m-ms, n], b[m]."""
        assert not self.single
        A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
        mA, n = A.shape
        m = b.shape[0]
        red = np.zeros(m, np.int32)
        self.lib.orc_minrep.restype = None
        self.lib.orc_minrep(red.ctypes.data_as(C.POINTER(C.c_int)), _ptr(A, self.real), _ptr(b, self.real),
                            C.c_int(n), C.c_int(m), C.c_int(m - mA))
        return red

    def minrep_independent(self, A, b):
        """Every probe against the full polyhedron (the batched semantics): (is_redundant, exitflag, iter)."""
        assert not self.single
        A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
        mA, n = A.shape
        m = b.shape[0]
        red = np.zeros(m, np.int32); flag = np.zeros(m, np.int32); it = np.zeros(m, np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.lib.orc_minrep_independent.restype = None
        self.lib.orc_minrep_independent(ip(red), ip(flag), ip(it), _ptr(A, self.real), _ptr(b, self.real),
                                        C.c_int(n), C.c_int(m), C.c_int(m - mA))
        return red, flag, it


def ref_minrep(A, b, name: str = "libdaqp_ref.so") -> np.ndarray:
    """The reference's own daqp_minrep (include/api.h:54) from oracle/_ref. A[m-ms, n], b[m]."""
    lib = C.CDLL(os.path.join(REF_DIR, name))
    lib.daqp_minrep.restype = None
    A = np.ascontiguousarray(A, np.float64).copy(); b = np.ascontiguousarray(b, np.float64).copy()
    mA, n = A.shape
    m = b.shape[0]
    red = np.zeros(m, np.int32)
    dp = C.POINTER(C.c_double)
    lib.daqp_minrep(red.ctypes.data_as(C.POINTER(C.c_int)), A.ctypes.data_as(dp), b.ctypes.data_as(dp),
                    C.c_int(n), C.c_int(m), C.c_int(m - mA))
    return red


def init_active(b, x=None, lam=None, sense=None, which: str = "oracle", name: str = "libdaqp_ref.so") -> np.ndarray:
    """daqp_primal_init_active (x given) / daqp_dual_init_active (lam given) per problem of the batch, through the oracle
    restatement (which="oracle") or the reference itself (which="ref", oracle/_ref). Returns the new sense [N, m]."""
    Problem = _F64[0]
    if which == "ref":
        lib = C.CDLL(os.path.join(REF_DIR, name))
        fp, fd = lib.daqp_primal_init_active, lib.daqp_dual_init_active
    else:
        lib = C.CDLL(os.path.join(HERE, "libdaqp_oracle.so"))
        fp, fd = lib.orc_primal_init_active, lib.orc_dual_init_active
    fp.restype = None; fd.restype = None
    out = np.ascontiguousarray(b.sense if sense is None else sense, dtype=np.int32).copy()
    real = C.c_double
    mA = b.m - b.ms
    for p in range(b.N):
        qp = Problem(b.n, b.m, b.ms, _ptr(b.H[p], real), _ptr(b.f[p], real), _ptr(b.A[p], real) if mA > 0 else None,
                     _ptr(b.bupper[p], real), _ptr(b.blower[p], real), out[p].ctypes.data_as(C.POINTER(C.c_int)),
                     None, 0, 0)
        if x is not None:
            v = np.ascontiguousarray(x[p], np.float64)
            fp(C.byref(qp), _ptr(v, real))
        else:
            v = np.ascontiguousarray(lam[p], np.float64)
            fd(C.byref(qp), _ptr(v, real))
    return out

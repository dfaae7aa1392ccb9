"""daqp_b200 -- B200-native batched dual active-set QP engine behind the DAQP C API.

Host-side mirror of the reference's Python interface for the hot path (reference:
interfaces/daqp-python/daqp.pyx:68-221 ``daqp.solve``), plus the batch entry points the reference lacks.
Everything here calls the C ABI in ``libdaqp_b200.so`` (include/daqp_b200.h) through ctypes; there is no
Python/CPU solve path -- if the CUDA library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DAQP_B200_LIB", os.path.join(_HERE, "libdaqp_b200.so"))

DAQP_INF = 1e30
EXIT_OPTIMAL, EXIT_SOFT_OPTIMAL = 1, 2
EXIT_INFEASIBLE, EXIT_CYCLE, EXIT_UNBOUNDED, EXIT_ITERLIMIT = -1, -2, -3, -4
EXIT_NONCONVEX, EXIT_OVERDETERMINED_INITIAL, EXIT_TIMELIMIT, EXIT_UNSUPPORTED = -5, -6, -7, -8
ACTIVE, LOWER, IMMUTABLE, SOFT, BINARY = 1, 2, 4, 8, 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class DAQPProblem(C.Structure):  # include/daqp_b200.h == reference include/types.h:14-50
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("ms", C.c_int), ("H", _dp), ("f", _dp), ("A", _dp),
                ("bupper", _dp), ("blower", _dp), ("sense", _ip), ("break_points", _ip), ("nh", C.c_int),
                ("problem_type", C.c_int)]


class DAQPSettings(C.Structure):  # reference include/types.h:52-74
    _fields_ = [("primal_tol", C.c_double), ("dual_tol", C.c_double), ("zero_tol", C.c_double),
                ("pivot_tol", C.c_double), ("progress_tol", C.c_double), ("cycle_tol", C.c_int),
                ("iter_limit", C.c_int), ("fval_bound", C.c_double), ("eps_prox", C.c_double),
                ("eta_prox", C.c_double), ("rho_soft", C.c_double), ("rel_subopt", C.c_double),
                ("abs_subopt", C.c_double), ("sing_tol", C.c_double), ("refactor_tol", C.c_double),
                ("time_limit", C.c_double)]


class DAQPResult(C.Structure):  # reference include/api.h:15-27
    _fields_ = [("x", _dp), ("lam", _dp), ("fval", C.c_double), ("soft_slack", C.c_double), ("exitflag", C.c_int),
                ("iter", C.c_int), ("nodes", C.c_int), ("solve_time", C.c_double), ("setup_time", C.c_double)]


_fp = C.POINTER(C.c_float)


class DAQPProblemF32(C.Structure):  # the reference's DAQPProblem with c_float = float (include/types.h:8-12,32-49)
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("ms", C.c_int), ("H", _fp), ("f", _fp), ("A", _fp), ("bupper", _fp),
                ("blower", _fp), ("sense", _ip), ("break_points", _ip), ("nh", C.c_int), ("problem_type", C.c_int)]


class DAQPResultF32(C.Structure):  # include/api.h:15-27 with c_float = float
    _fields_ = [("x", _fp), ("lam", _fp), ("fval", C.c_float), ("soft_slack", C.c_float), ("exitflag", C.c_int),
                ("iter", C.c_int), ("nodes", C.c_int), ("solve_time", C.c_float), ("setup_time", C.c_float)]


class DAQPB200Diag(C.Structure):
    _fields_ = [("n_active", _ip), ("ws", _ip), ("counts", _ip), ("sense", C.POINTER(C.c_ubyte)), ("soft_slack", _dp),
                ("trace", _ip), ("trace_cap", C.c_int)]


class DAQPB200Stats(C.Structure):
    _fields_ = [("setup_launches", C.c_int), ("solve_launches", C.c_int), ("setup_ms", C.c_double),
                ("solve_ms", C.c_double), ("warps_per_sm", C.c_int), ("scratch_bytes", C.c_longlong)]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library. Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m daqp_b200.build` "
                               "(daqp_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.daqp_quadprog.restype = None
        L.daqp_default_settings.restype = None
        L.daqp_quadprog_batch.restype = C.c_int
        L.daqp_b200_create.restype = C.c_int
        L.daqp_b200_destroy.restype = None
        L.daqp_b200_solve_packed.restype = C.c_int
        L.daqp_b200_solve_device.restype = C.c_int
        L.daqp_b200_get_stats.restype = C.c_int
        L.daqp_b200_set_scratch_limit.restype = None
        L.daqp_b200_set_scratch_limit.argtypes = [C.c_void_p, C.c_longlong]
        L.daqp_b200_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"daqp_b200 error {rc}: {lib().daqp_b200_last_error().decode()}")


def default_settings(**over) -> DAQPSettings:
    s = DAQPSettings()
    lib().daqp_default_settings(C.byref(s))
    for k, v in over.items():
        if not hasattr(s, k):
            raise TypeError(f"unknown setting {k!r}")
        setattr(s, k, v)
    return s


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def solve(H, f, A, bupper, blower=None, sense=None, primal_start=None, dual_start=None, **settings):
    """Single QP through the drop-in ``daqp_quadprog`` symbol (same call shape and return value as the
    reference's ``daqp.solve``: ``x, fval, exitflag, info``). bupper/blower longer than A's row count means the
    leading entries are simple bounds (reference daqp.pyx:96-104)."""
    A = _f64(A)
    H = _f64(H); f = _f64(f); bupper = _f64(bupper)
    m = bupper.shape[0]
    # (H = None with f = None is the LDP min |x|^2 s.t. the constraints; without general rows every bound is a simple one)
    mA, n = (A.shape if A is not None and A.size else (0, H.shape[0] if H is not None else m))
    blower = np.full(m, -DAQP_INF) if blower is None else _f64(blower)
    sense = np.zeros(m, dtype=np.intc) if sense is None else np.ascontiguousarray(sense, dtype=np.intc)
    x = np.empty(n); lam = np.empty(m)
    qp = DAQPProblem(n, m, m - mA, _p(H), _p(f), _p(A) if mA else None, _p(bupper), _p(blower), _p(sense, _ip),
                     None, 0, 0)
    if (primal_start is not None or dual_start is not None) and m > 0:  # daqp.pyx:24-38 (sense is a private copy here)
        sense = sense.copy(); qp.sense = _p(sense, _ip)
        L = lib()
        L.daqp_dual_init_active.restype = None; L.daqp_primal_init_active.restype = None
        if dual_start is not None:
            L.daqp_dual_init_active(C.byref(qp), _p(_f64(dual_start)))
        else:
            L.daqp_primal_init_active(C.byref(qp), _p(_f64(primal_start)))
    st = default_settings(**settings)
    res = DAQPResult(_p(x), _p(lam) if m else None, 0, 0, 0, 0, 0, 0, 0)
    lib().daqp_quadprog(C.byref(res), C.byref(qp), C.byref(st))
    return x, res.fval, res.exitflag, {"solve_time": res.solve_time, "setup_time": res.setup_time,
                                       "iterations": res.iter, "nodes": res.nodes, "lam": lam,
                                       "soft_slack": res.soft_slack}


class BatchResult:
    __slots__ = ("x", "lam", "fval", "exitflag", "iter", "n_active", "ws", "counts", "sense", "soft_slack")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))

    def working_sets(self):
        return [self.ws[p, : self.n_active[p]].tolist() for p in range(len(self.n_active))]


class Engine:
    """Owns one DAQPB200Handle (device scratch + streams)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        _check(lib().daqp_b200_create(C.byref(self._h), device))

    def close(self):
        if self._h:
            lib().daqp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scratch_limit(self, nbytes: int):
        lib().daqp_b200_set_scratch_limit(self._h, int(nbytes))

    def stats(self, reset: bool = False) -> dict:
        s = DAQPB200Stats()
        _check(lib().daqp_b200_get_stats(self._h, C.byref(s), int(reset)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    # -- host arrays -----------------------------------------------------------------------------------------
    def solve_batch(self, H, f, A, bupper, blower, sense=None, ms: int | None = None, diag: bool = False,
                    out: BatchResult | None = None, primal_start=None, dual_start=None, **settings) -> BatchResult:
        """Homogeneous batch in host memory: H[N,n,n], f[N,n]|None, A[N,m-ms,n], bupper/blower[N,m],
        sense[N,m]|None (``daqp_b200_solve_packed``). Pass pinned arrays for asynchronous copies.
        ``primal_start[N,n]`` / ``dual_start[N,m]`` warm-start the working sets like the reference's
        ``daqp.solve(..., primal_start=, dual_start=)`` (daqp.pyx:24-38): the init helpers set the ACTIVE bits of a copy of
        ``sense`` (dual_start wins when both are given, as in the reference)."""
        H = _f64(H); f = _f64(f); A = _f64(A); bupper = _f64(bupper); blower = _f64(blower)
        if dual_start is not None:
            sense = self.init_active_batch(A, bupper, blower, sense, lam=dual_start, ms=ms)
        elif primal_start is not None:
            sense = self.init_active_batch(A, bupper, blower, sense, x=primal_start, ms=ms)
        N, n = H.shape[0], H.shape[1]
        m = bupper.shape[1]
        mA = A.shape[1] if A is not None and A.size else 0
        ms = m - mA if ms is None else ms
        if sense is not None:
            sense = np.ascontiguousarray(sense, dtype=np.intc)
        r = out or BatchResult(x=np.empty((N, n)), lam=np.empty((N, m)), fval=np.zeros(N),
                               exitflag=np.empty(N, np.intc), iter=np.empty(N, np.intc))
        d = None
        if diag:
            ldm = (max(m, 1) + 3) // 4 * 4
            ns = 0 if sense is None else int(((sense & SOFT) != 0).sum(axis=1).max(initial=0))
            r.n_active = np.zeros(N, np.intc); r.ws = np.zeros((N, n + ns + 1), np.intc)
            r.counts = np.zeros((N, 8), np.intc); r.sense = np.zeros((N, ldm), np.uint8)
            r.soft_slack = np.zeros(N)
            d = DAQPB200Diag(_p(r.n_active, _ip), _p(r.ws, _ip), _p(r.counts, _ip),
                             r.sense.ctypes.data_as(C.POINTER(C.c_ubyte)), _p(r.soft_slack))
        st = default_settings(**settings)
        _check(lib().daqp_b200_solve_packed(self._h, N, n, m, ms, _p(H), _p(f), _p(A), _p(bupper), _p(blower),
                                            _p(sense, _ip), C.byref(st), _p(r.x), _p(r.lam), _p(r.fval),
                                            _p(r.exitflag, _ip), _p(r.iter, _ip), C.byref(d) if d else None))
        if diag:
            r.sense = r.sense[:, :m]
        return r

    def solve_batch_f32(self, H, f, A, bupper, blower, sense=None, ms: int | None = None, diag: bool = False,
                        **settings) -> BatchResult:
        """Same batch in fp32 arithmetic end to end (``daqp_b200_solve_packed_f32``): the batched form of the reference
        built with -DDAQP_SINGLE_PRECISION. Inputs are converted to float32; results are float32 arrays."""
        L = lib()
        L.daqp_b200_solve_packed_f32.restype = C.c_int
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        H = f32(H); f = f32(f); A = f32(A); bupper = f32(bupper); blower = f32(blower)
        N, n = H.shape[0], H.shape[1]
        m = bupper.shape[1]
        mA = A.shape[1] if A is not None and A.size else 0
        ms = m - mA if ms is None else ms
        if sense is not None:
            sense = np.ascontiguousarray(sense, dtype=np.intc)
        r = BatchResult(x=np.empty((N, n), np.float32), lam=np.empty((N, m), np.float32), fval=np.zeros(N, np.float32),
                        exitflag=np.empty(N, np.intc), iter=np.empty(N, np.intc))
        d = None
        if diag:
            ldm = (max(m, 1) + 3) // 4 * 4
            r.n_active = np.zeros(N, np.intc); r.ws = np.zeros((N, n + 1), np.intc)
            r.counts = np.zeros((N, 8), np.intc); r.sense = np.zeros((N, ldm), np.uint8)
            d = DAQPB200Diag(_p(r.n_active, _ip), _p(r.ws, _ip), _p(r.counts, _ip),
                             r.sense.ctypes.data_as(C.POINTER(C.c_ubyte)), None)
        st = default_settings(**settings)
        _check(L.daqp_b200_solve_packed_f32(self._h, N, n, m, ms, _p(H, fp), _p(f, fp), _p(A, fp), _p(bupper, fp),
                                            _p(blower, fp), _p(sense, _ip), C.byref(st), _p(r.x, fp), _p(r.lam, fp),
                                            _p(r.fval, fp), _p(r.exitflag, _ip), _p(r.iter, _ip),
                                            C.byref(d) if d else None))
        if diag:
            r.sense = r.sense[:, :m]
        return r

    # -- device arrays (torch CUDA tensors) -------------------------------------------------------------------
    def solve_batch_device(self, H, f, A, bupper, blower, sense=None, ms: int | None = None, out=None,
                           diag=None, stream=None, **settings):
        """Same on CUDA tensors (float64, contiguous); enqueues on the current torch stream and returns a dict of
        output tensors without synchronising (``daqp_b200_solve_device``)."""
        import torch
        N, n = H.shape[0], H.shape[1]
        m = bupper.shape[1]
        mA = A.shape[1] if A is not None and A.numel() else 0
        ms = m - mA if ms is None else ms
        dev = H.device
        for t in (H, f, A, bupper, blower):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        if sense is not None:
            assert sense.is_cuda and sense.dtype == torch.int32 and sense.is_contiguous()
        if out is None:
            out = {"x": torch.empty((N, n), dtype=torch.float64, device=dev),
                   "lam": torch.empty((N, m), dtype=torch.float64, device=dev),
                   "fval": torch.zeros(N, dtype=torch.float64, device=dev),
                   "exitflag": torch.empty(N, dtype=torch.int32, device=dev),
                   "iter": torch.empty(N, dtype=torch.int32, device=dev)}
        d = None
        if diag is not None:
            d = DAQPB200Diag(C.cast(diag["n_active"].data_ptr(), _ip), C.cast(diag["ws"].data_ptr(), _ip),
                             C.cast(diag["counts"].data_ptr(), _ip),
                             C.cast(diag["sense"].data_ptr(), C.POINTER(C.c_ubyte)),
                             C.cast(diag["soft_slack"].data_ptr(), _dp) if "soft_slack" in diag else None,
                             C.cast(diag["trace"].data_ptr(), _ip) if "trace" in diag else None,
                             int((diag["trace"].shape[1] - 1) // 2) if "trace" in diag else 0)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        if stream == 0:
            stream = 1  # cudaStreamLegacy: name torch's default stream explicitly (NULL means "the engine's own")
        ptr = lambda t, ty=_dp: None if t is None else C.cast(t.data_ptr(), ty)
        st = default_settings(**settings)
        _check(lib().daqp_b200_solve_device(self._h, N, n, m, ms, ptr(H), ptr(f), ptr(A), ptr(bupper), ptr(blower),
                                            ptr(sense, _ip), C.byref(st), ptr(out["x"]), ptr(out["lam"]),
                                            ptr(out["fval"]), ptr(out["exitflag"], _ip), ptr(out["iter"], _ip),
                                            C.byref(d) if d else None, C.c_void_p(stream)))
        return out

    def solve_batch_device_f32(self, H, f, A, bupper, blower, sense=None, ms: int | None = None, out=None, stream=None,
                               **settings):
        """fp32 CUDA tensors through ``daqp_b200_solve_device_f32`` (asynchronous on the current torch stream)."""
        import torch
        L = lib()
        L.daqp_b200_solve_device_f32.restype = C.c_int
        N, n = H.shape[0], H.shape[1]
        m = bupper.shape[1]
        mA = A.shape[1] if A is not None and A.numel() else 0
        ms = m - mA if ms is None else ms
        dev = H.device
        for t in (H, f, A, bupper, blower):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        if out is None:
            out = {"x": torch.empty((N, n), dtype=torch.float32, device=dev),
                   "lam": torch.empty((N, m), dtype=torch.float32, device=dev),
                   "fval": torch.zeros(N, dtype=torch.float32, device=dev),
                   "exitflag": torch.empty(N, dtype=torch.int32, device=dev),
                   "iter": torch.empty(N, dtype=torch.int32, device=dev)}
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        if stream == 0:
            stream = 1
        fp = C.POINTER(C.c_float)
        ptr = lambda t, ty=fp: None if t is None else C.cast(t.data_ptr(), ty)
        st = default_settings(**settings)
        _check(L.daqp_b200_solve_device_f32(self._h, N, n, m, ms, ptr(H), ptr(f), ptr(A), ptr(bupper), ptr(blower),
                                            ptr(sense, _ip), C.byref(st), ptr(out["x"]), ptr(out["lam"]),
                                            ptr(out["fval"]), ptr(out["exitflag"], _ip), ptr(out["iter"], _ip),
                                            None, C.c_void_p(stream)))
        return out

    # -- warm-start initialisers (reference daqp_primal_init_active / daqp_dual_init_active) ------------------------------
    def init_active_batch(self, A, bupper, blower, sense=None, x=None, lam=None, ms: int | None = None) -> np.ndarray:
        """New ``sense[N, m]`` with the ACTIVE / LOWER bits the reference's init helpers would set from the primal iterate
        ``x[N, n]`` or the dual iterate ``lam[N, m]`` (``daqp_b200_init_active``). ``sense`` (or zeros) is not modified."""
        L = lib()
        L.daqp_b200_init_active.restype = C.c_int
        A = _f64(A); bupper = _f64(bupper); blower = _f64(blower); x = _f64(x); lam = _f64(lam)
        N, m = bupper.shape
        mA = A.shape[1] if A is not None and A.size else 0
        ms = m - mA if ms is None else ms
        n = A.shape[2] if mA else (x.shape[1] if x is not None else max(ms, 1))  # n is not used by the dual pass
        out = np.zeros((N, m), np.intc) if sense is None else np.ascontiguousarray(sense, dtype=np.intc).copy()
        _check(L.daqp_b200_init_active(self._h, N, n, m, ms, _p(x), _p(lam), _p(A), _p(bupper), _p(blower), _p(out, _ip)))
        return out

    def init_active_device(self, sense, A, bupper, blower, x=None, lam=None, ms: int | None = None, stream=None):
        """CUDA tensors; ``sense`` (int32 [N, m]) is updated IN PLACE, asynchronously on the current torch stream
        (``daqp_b200_init_active_device``)."""
        import torch
        L = lib()
        L.daqp_b200_init_active_device.restype = C.c_int
        N, m = bupper.shape
        mA = A.shape[1] if A is not None and A.numel() else 0
        ms = m - mA if ms is None else ms
        n = A.shape[2] if mA else (x.shape[1] if x is not None else max(ms, 1))
        assert sense.is_cuda and sense.dtype == torch.int32 and sense.is_contiguous()
        for t in (A, bupper, blower, x, lam):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream(bupper.device).cuda_stream
        if stream == 0:
            stream = 1
        ptr = lambda t, ty=_dp: None if t is None else C.cast(t.data_ptr(), ty)
        _check(L.daqp_b200_init_active_device(self._h, N, n, m, ms, ptr(x), ptr(lam), ptr(A), ptr(bupper), ptr(blower),
                                              ptr(sense, _ip), C.c_void_p(stream)))
        return sense

    # -- minimal representation of polyhedra (batched LDP consumer) ----------------------------------------------
    def minrep_batch(self, A, b, ms: int | None = None, info: bool = False, **settings):
        """P polyhedra {x : [I(ms); A[q]] x <= b[q]} in host memory: A[P,m-ms,n], b[P,m] -> is_redundant[P,m]
        (``daqp_b200_minrep_batch``; the batched form of the reference's ``daqp.minrep``, daqp.pyx:636-652). With
        ``info`` also returns the exit flag and iteration count of every LDP."""
        L = lib()
        L.daqp_b200_minrep_batch.restype = C.c_int
        A = _f64(A); b = _f64(b)
        P, m = b.shape
        mA, n = A.shape[1], A.shape[2]
        ms = m - mA if ms is None else ms
        red = np.empty((P, m), np.intc); flag = np.empty((P, m), np.intc); it = np.empty((P, m), np.intc)
        st = default_settings(**settings)
        _check(L.daqp_b200_minrep_batch(self._h, P, n, m, ms, _p(A), _p(b), C.byref(st), _p(red, _ip),
                                        _p(flag, _ip), _p(it, _ip)))
        return (red, flag, it) if info else red

    def minrep_batch_device(self, A, b, ms: int | None = None, out=None, stream=None, **settings):
        """CUDA tensors (float64, contiguous) through ``daqp_b200_minrep_device``; asynchronous on the current torch
        stream. Returns {"is_redundant", "exitflag", "iter"} int32 tensors of shape [P, m]."""
        import torch
        L = lib()
        L.daqp_b200_minrep_device.restype = C.c_int
        P, m = b.shape
        mA, n = A.shape[1], A.shape[2]
        ms = m - mA if ms is None else ms
        for t in (A, b):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        if out is None:
            out = {k: torch.empty((P, m), dtype=torch.int32, device=b.device) for k in ("is_redundant", "exitflag", "iter")}
        if stream is None:
            stream = torch.cuda.current_stream(b.device).cuda_stream
        if stream == 0:
            stream = 1
        ptr = lambda t, ty=_dp: C.cast(t.data_ptr(), ty)
        st = default_settings(**settings)
        _check(L.daqp_b200_minrep_device(self._h, P, n, m, ms, ptr(A), ptr(b), C.byref(st),
                                         ptr(out["is_redundant"], _ip), ptr(out["exitflag"], _ip),
                                         ptr(out["iter"], _ip), C.c_void_p(stream)))
        return out

    @staticmethod
    def alloc_diag(N: int, n: int, m: int, device, ns: int = 0):
        """ns = the largest number of soft constraints per problem (rows of ``ws`` hold n + ns + 1 entries)."""
        import torch
        ldm = (max(m, 1) + 3) // 4 * 4
        return {"n_active": torch.zeros(N, dtype=torch.int32, device=device),
                "soft_slack": torch.zeros(N, dtype=torch.float64, device=device),
                "ws": torch.zeros((N, n + ns + 1), dtype=torch.int32, device=device),
                "counts": torch.zeros((N, 8), dtype=torch.int32, device=device),
                "sense": torch.zeros((N, ldm), dtype=torch.uint8, device=device)}


class BatchModel:
    """Batched counterpart of the reference's ``daqp.Model`` (interfaces/daqp-python/daqp.pyx:224-572): ``setup`` once,
    then ``update(f=..., bupper=..., blower=...)`` + ``solve()`` per step. The QP -> LDP transform (Cholesky, R^-1,
    M = A R^-1, row scaling) stays on the device; ``solve(warm=True)`` continues every problem from the LDL' factor,
    multipliers and working set its previous solve ended with (``daqp_b200_workspace_*``)."""

    def __init__(self, engine: "Engine | None" = None):
        self._eng = engine
        self._w = C.c_void_p()
        self.shape = None

    def setup(self, H, f, A, bupper, blower, sense=None, ms: int | None = None, **settings):
        L = lib()
        L.daqp_b200_workspace_setup.restype = C.c_int
        L.daqp_b200_workspace_update.restype = C.c_int
        L.daqp_b200_workspace_solve.restype = C.c_int
        L.daqp_b200_workspace_free.restype = None
        self.close()
        H = _f64(H); f = _f64(f); A = _f64(A); bupper = _f64(bupper); blower = _f64(blower)
        N, n = H.shape[0], H.shape[1]
        m = bupper.shape[1]
        mA = A.shape[1] if A is not None and A.size else 0
        ms = m - mA if ms is None else ms
        ns = 0
        if sense is not None:
            sense = np.ascontiguousarray(sense, dtype=np.intc)
            ns = int(((sense & SOFT) != 0).sum(axis=1).max(initial=0))
        st = default_settings(**settings)
        h = self._eng._h if self._eng is not None else None
        _check(L.daqp_b200_workspace_setup(h, N, n, m, ms, _p(H), _p(f), _p(A), _p(bupper), _p(blower),
                                           _p(sense, _ip), C.byref(st), C.byref(self._w)))
        self.shape = (N, n, m, ms, ns)
        return self

    def setup_shared(self, H, A, K: int, sense=None, ms: int | None = None, m: int | None = None, **settings):
        """Shared workspace (``daqp_b200_workspace_setup_shared``): G matrix sets H[G,n,n], A[G,m-ms,n] (sense[G,m] or
        None), K problems per set -- the parametric / MPC case of ONE controller evaluated for many states, where only
        f and the bounds differ. The QP -> LDP transform runs once per set and the K problems of a set stream the same
        matrices out of L2. Problem p = g * K + k; ``update(f[G*K,n], bupper[G*K,m], blower[G*K,m])`` must give all
        three arrays before the first ``solve``. Every problem gets the result of the reference's
        ``setup_daqp(H_g, f_p, A_g, b_p)`` + ``daqp_solve``. ``m`` is only needed when A is empty (m == ms)."""
        L = lib()
        L.daqp_b200_workspace_setup_shared.restype = C.c_int
        L.daqp_b200_workspace_update.restype = C.c_int
        L.daqp_b200_workspace_solve.restype = C.c_int
        L.daqp_b200_workspace_free.restype = None
        self.close()
        H = _f64(H); A = _f64(A)
        G, n = H.shape[0], H.shape[1]
        mA = A.shape[1] if A is not None and A.size else 0
        if sense is not None:
            sense = np.ascontiguousarray(sense, dtype=np.intc)
            m = sense.shape[1]
        if m is None:
            m = mA + (0 if ms is None else ms)
        ms = m - mA if ms is None else ms
        ns = 0 if sense is None else int(((sense & SOFT) != 0).sum(axis=1).max(initial=0))
        st = default_settings(**settings)
        h = self._eng._h if self._eng is not None else None
        _check(L.daqp_b200_workspace_setup_shared(h, G, int(K), n, m, ms, _p(H), _p(A), _p(sense, _ip), C.byref(st),
                                                  C.byref(self._w)))
        self.shape = (G * int(K), n, m, ms, ns)
        return self

    def update(self, f=None, bupper=None, blower=None):
        f = _f64(f); bupper = _f64(bupper); blower = _f64(blower)
        _check(lib().daqp_b200_workspace_update(self._w, _p(f), _p(bupper), _p(blower)))
        return self

    def solve(self, warm: bool = True, diag: bool = False, out: BatchResult | None = None) -> BatchResult:
        """``out``: reuse result arrays (pinned host memory keeps the read-back asynchronous)."""
        N, n, m, ms, ns = self.shape
        r = out or BatchResult(x=np.empty((N, n)), lam=np.empty((N, m)), fval=np.zeros(N),
                               exitflag=np.empty(N, np.intc), iter=np.empty(N, np.intc))
        d = None
        if diag:
            ldm = (max(m, 1) + 3) // 4 * 4
            r.n_active = np.zeros(N, np.intc); r.ws = np.zeros((N, n + ns + 1), np.intc)
            r.counts = np.zeros((N, 8), np.intc); r.sense = np.zeros((N, ldm), np.uint8); r.soft_slack = np.zeros(N)
            d = DAQPB200Diag(_p(r.n_active, _ip), _p(r.ws, _ip), _p(r.counts, _ip),
                             r.sense.ctypes.data_as(C.POINTER(C.c_ubyte)), _p(r.soft_slack))
        _check(lib().daqp_b200_workspace_solve(self._w, int(warm), _p(r.x), _p(r.lam), _p(r.fval),
                                               _p(r.exitflag, _ip), _p(r.iter, _ip), C.byref(d) if d else None))
        if diag:
            r.sense = r.sense[:, :m]
        return r

    # -- closed loops that live on the GPU: torch CUDA tensors in, torch CUDA tensors out, nothing synchronises ----------
    def update_device(self, f=None, bupper=None, blower=None, stream=None):
        import torch
        L = lib()
        L.daqp_b200_workspace_update_device.restype = C.c_int
        for t in (f, bupper, blower):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        ptr = lambda t: None if t is None else C.cast(t.data_ptr(), _dp)
        _check(L.daqp_b200_workspace_update_device(self._w, ptr(f), ptr(bupper), ptr(blower), self._stream(stream)))
        return self

    def solve_device(self, warm: bool = True, out=None, stream=None):
        import torch
        L = lib()
        L.daqp_b200_workspace_solve_device.restype = C.c_int
        N, n, m, ms, ns = self.shape
        if out is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            out = {"x": torch.empty((N, n), dtype=torch.float64, device=dev),
                   "lam": torch.empty((N, m), dtype=torch.float64, device=dev),
                   "fval": torch.zeros(N, dtype=torch.float64, device=dev),
                   "exitflag": torch.empty(N, dtype=torch.int32, device=dev),
                   "iter": torch.empty(N, dtype=torch.int32, device=dev)}
        ptr = lambda t, ty=_dp: C.cast(t.data_ptr(), ty)
        _check(L.daqp_b200_workspace_solve_device(self._w, int(warm), ptr(out["x"]), ptr(out["lam"]), ptr(out["fval"]),
                                                  ptr(out["exitflag"], _ip), ptr(out["iter"], _ip), self._stream(stream)))
        return out

    @staticmethod
    def _stream(stream):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        return C.c_void_p(1 if stream == 0 else stream)  # cudaStreamLegacy names torch's default stream explicitly

    def close(self):
        if self._w:
            lib().daqp_b200_workspace_free(self._w)
            self._w = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_engine = None


def solve_batch(H, f, A, bupper, blower, sense=None, **kw) -> BatchResult:
    """Module-level convenience: ``Engine().solve_batch`` on a process-wide engine."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine.solve_batch(H, f, A, bupper, blower, sense, **kw)


def solve_batch_multi(H, f, A, bupper, blower, sense=None, ms: int | None = None, devices=None, out: BatchResult | None = None,
                      **settings):
    """One process, several GPUs (``daqp_b200_solve_packed_multi``): the batch is cut into contiguous blocks, one host
    thread + engine per device, no exchange between devices. ``devices``: list of CUDA device indices (default: all
    visible). Returns ``(BatchResult, seconds_per_device)``."""
    L = lib()
    L.daqp_b200_solve_packed_multi.restype = C.c_int
    H = _f64(H); f = _f64(f); A = _f64(A); bupper = _f64(bupper); blower = _f64(blower)
    N, n = H.shape[0], H.shape[1]
    m = bupper.shape[1]
    mA = A.shape[1] if A is not None and A.size else 0
    ms = m - mA if ms is None else ms
    if sense is not None:
        sense = np.ascontiguousarray(sense, dtype=np.intc)
    if devices is None:
        import torch
        devices = list(range(torch.cuda.device_count()))
    dv = np.asarray(devices, dtype=np.intc)
    r = out or BatchResult(x=np.empty((N, n)), lam=np.empty((N, m)), fval=np.zeros(N), exitflag=np.empty(N, np.intc),
                           iter=np.empty(N, np.intc))
    secs = np.zeros(len(dv))
    st = default_settings(**settings)
    _check(L.daqp_b200_solve_packed_multi(C.c_int(len(dv)), _p(dv, _ip), N, n, m, ms, _p(H), _p(f), _p(A), _p(bupper),
                                          _p(blower), _p(sense, _ip), C.byref(st), _p(r.x), _p(r.lam), _p(r.fval),
                                          _p(r.exitflag, _ip), _p(r.iter, _ip), _p(secs)))
    return r, secs


def minrep(A, b):
    """Drop-in for the reference's ``daqp.minrep(A, b)`` (interfaces/daqp-python/daqp.pyx:636-652): which constraints
    of {x : A x <= b} are redundant. ``b`` longer than A's row count means the leading entries are simple bounds.
    Goes through the ``daqp_minrep`` symbol (reference include/api.h:54); the m LDPs run concurrently on the GPU."""
    L = lib()
    L.daqp_minrep.restype = None
    A = _f64(A); b = _f64(b)
    mA, n = A.shape
    m = b.shape[0]
    red = np.zeros(m, np.intc)
    L.daqp_minrep(_p(red, _ip), _p(A), _p(b), C.c_int(n), C.c_int(m), C.c_int(m - mA))
    if m and red[0] < 0:
        raise RuntimeError(f"daqp_minrep failed: {L.daqp_b200_last_error().decode()}")
    return red


def minrep_batch(A, b, **kw):
    """Module-level convenience: ``Engine().minrep_batch`` on a process-wide engine."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine.minrep_batch(A, b, **kw)


def ldp_batch(A, bupper, blower=None, sense=None, **settings) -> BatchResult:
    """P raw LDPs  min |u|^2  s.t.  blower <= [I(ms); A] u <= bupper  in one call (``daqp_b200_ldp_batch``): A[P, m-ms, n]
    as given (no normalisation), bupper / blower [P, m] (ms = m - A.shape[1] leading simple bounds), sense [P, m] or None.
    The batched form of the feasibility checks Julia's polyhedral tools run through ``daqp_ldp`` on a hand-filled
    workspace (api.jl:428-459): ``exitflag`` 1 = non-empty (``x`` = its point of least norm), -1 = empty or outside the
    radius ``fval_bound``. ``fval`` = |u|^2 / 2; ``ws`` / ``n_active`` = working sets in factor order."""
    A = _f64(A); bupper = _f64(bupper)
    P, m = bupper.shape
    mA, n = A.shape[1], A.shape[2]
    blower = np.full((P, m), -DAQP_INF) if blower is None else _f64(blower)
    se = None if sense is None else np.ascontiguousarray(sense, dtype=np.intc)
    r = BatchResult(x=np.zeros((P, n)), lam=np.zeros((P, m)), fval=np.zeros(P), exitflag=np.empty(P, np.intc),
                    iter=np.empty(P, np.intc), n_active=np.zeros(P, np.intc), ws=np.zeros((P, n + 1), np.intc))
    d = DAQPB200Diag(_p(r.n_active, _ip), _p(r.ws, _ip), None, None, None)
    st = default_settings(**settings)
    L = lib()
    L.daqp_b200_ldp_batch.restype = C.c_int
    _check(L.daqp_b200_ldp_batch(None, P, n, m, m - mA, _p(A) if A.size else None, _p(bupper), _p(blower), _p(se, _ip),
                                 C.byref(st), _p(r.x), _p(r.lam), _p(r.fval), _p(r.exitflag, _ip), _p(r.iter, _ip), C.byref(d)))
    return r


def quadprog_batch(problems: list[dict], **settings):
    """Array-of-struct batch through ``daqp_quadprog_batch``: a list of dicts with keys H, f, A, bupper, blower,
    sense (the last three optional as in ``solve``); problems may differ in size. Returns a list of
    ``(x, fval, exitflag, info)`` tuples."""
    N = len(problems)
    qps = (DAQPProblem * N)(); res = (DAQPResult * N)()
    keep = []
    for i, pr in enumerate(problems):
        H = _f64(pr.get("H")); f = _f64(pr.get("f")); A = _f64(pr.get("A")); bu = _f64(pr["bupper"])
        m = bu.shape[0]
        mA = A.shape[0] if A is not None and A.size else 0
        n = H.shape[0] if H is not None else (A.shape[1] if mA else m)  # H = None, f = None: the LDP min |x|^2
        bl = np.full(m, -DAQP_INF) if pr.get("blower") is None else _f64(pr["blower"])
        se = None if pr.get("sense") is None else np.ascontiguousarray(pr["sense"], dtype=np.intc)
        x = np.empty(n); lam = np.empty(m)
        keep.append((H, f, A, bu, bl, se, x, lam))
        qps[i] = DAQPProblem(n, m, m - mA, _p(H), _p(f), _p(A) if mA else None, _p(bu), _p(bl), _p(se, _ip), None, 0, 0)
        res[i] = DAQPResult(_p(x), _p(lam) if m else None, 0, 0, 0, 0, 0, 0, 0)
    st = default_settings(**settings)
    _check(lib().daqp_quadprog_batch(N, qps, res, C.byref(st)))
    return [(keep[i][6], res[i].fval, res[i].exitflag,
             {"iterations": res[i].iter, "lam": keep[i][7], "solve_time": res[i].solve_time,
              "setup_time": res[i].setup_time, "nodes": res[i].nodes, "soft_slack": res[i].soft_slack})
            for i in range(N)]


def quadprog_batch_f32(problems: list[dict], **settings):
    """The same array-of-struct batch in fp32 arithmetic (``daqp_quadprog_batch_f32``: the structs of the reference's
    -DDAQP_SINGLE_PRECISION build). Mixed shapes in one call; x / lam come back as float32 arrays."""
    L = lib()
    L.daqp_quadprog_batch_f32.restype = C.c_int
    N = len(problems)
    qps = (DAQPProblemF32 * N)(); res = (DAQPResultF32 * N)()
    keep = []
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    pf = lambda a: None if a is None else a.ctypes.data_as(_fp)
    for i, pr in enumerate(problems):
        H = f32(pr["H"]); f = f32(pr.get("f")); A = f32(pr.get("A")); bu = f32(pr["bupper"])
        m = bu.shape[0]
        n = H.shape[0]
        mA = A.shape[0] if A is not None and A.size else 0
        bl = np.full(m, -DAQP_INF, np.float32) if pr.get("blower") is None else f32(pr["blower"])
        se = None if pr.get("sense") is None else np.ascontiguousarray(pr["sense"], dtype=np.intc)
        x = np.empty(n, np.float32); lam = np.empty(m, np.float32)
        keep.append((H, f, A, bu, bl, se, x, lam))
        qps[i] = DAQPProblemF32(n, m, m - mA, pf(H), pf(f), pf(A) if mA else None, pf(bu), pf(bl), _p(se, _ip), None, 0, 0)
        res[i] = DAQPResultF32(pf(x), pf(lam) if m else None, 0, 0, 0, 0, 0, 0, 0)
    st = default_settings(**settings)
    _check(L.daqp_quadprog_batch_f32(N, qps, res, C.byref(st)))
    return [(keep[i][6], res[i].fval, res[i].exitflag,
             {"iterations": res[i].iter, "lam": keep[i][7], "solve_time": res[i].solve_time,
              "setup_time": res[i].setup_time, "nodes": res[i].nodes})
            for i in range(N)]

"""Warm-start initialisers (reference daqp_primal_init_active / daqp_dual_init_active: include/api.h:57-58,
src/api.c:577-631) -- the callers on the input side of the hot path.

CPU part: the oracle restatement against golden vectors produced by the reference itself and against the reference
compiled here. GPU part (-m gpu): the batched CUDA kernel through the C ABI; sense bits must be EQUAL (integer output),
and the warm-started solves must reproduce the reference's iteration counts (1 from the optimum, core_tests.jl:520-543).
"""
import glob
import os

import numpy as np
import pytest

from common import GOLDEN_DIR
from daqp_b200.problems import QPBatch

NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "warmstart_*.npz")))
ITERATES = [("x_opt", "x"), ("x_in", "x"), ("x_out", "x"), ("x_far", "x"), ("lam_opt", "lam"), ("lam_noisy", "lam")]


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, m, ms = int(d["n"]), int(d["m"]), int(d["ms"])
    b = QPBatch(n, m, ms, d["H"], d["f"], d["A"], d["bupper"], d["blower"], np.zeros(d["bupper"].shape, np.int32))
    return b, d


@pytest.mark.parametrize("name", NAMES)
def test_oracle_init_active_matches_golden(oracle_libs, name):
    b, d = load(name)
    for key, kind in ITERATES:
        for tag, s0 in (("", None), ("_s0", d["sense0"])):
            got = oracle_libs.init_active(b, sense=s0, **{kind: d[key]})
            np.testing.assert_array_equal(got, d[f"sense_{key}{tag}"], err_msg=f"{name} {key}{tag}")
    assert (d["iter_x_opt"] == 1).all() and (d["iter_lam_opt"] == 1).all()  # the reference's own gate
    s0 = d["sense0"]
    keep = (s0 & 4) != 0
    assert (d["sense_x_opt_s0"][keep] == s0[keep]).all()  # IMMUTABLE rows are left alone


def test_oracle_init_active_vs_live_reference(oracle_libs):
    if not oracle_libs.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    b, d = load("warmstart_n20_m60_ms5")
    rng = np.random.default_rng(3)
    for scale in (0.0, 1e-10, 1e-9, 1e-8):
        x = d["x_opt"] + scale * rng.standard_normal(d["x_opt"].shape)
        for lib in ("libdaqp_ref.so", "libdaqp_ref_strict.so"):
            np.testing.assert_array_equal(oracle_libs.init_active(b, x=x),
                                          oracle_libs.init_active(b, x=x, which="ref", name=lib))


# ---- GPU -------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def engine(cuda_lib):
    import daqp_b200
    e = daqp_b200.Engine()
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_init_active_matches_golden(engine, name):
    import torch
    b, d = load(name)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for key, kind in ITERATES:
        for tag, s0 in (("", None), ("_s0", d["sense0"])):
            got = engine.init_active_batch(b.A, b.bupper, b.blower, s0, ms=b.ms, **{kind: d[key]})
            np.testing.assert_array_equal(got, d[f"sense_{key}{tag}"], err_msg=f"{name} {key}{tag} (host entry)")
            se = t(np.zeros_like(d["sense0"]) if s0 is None else s0).to(torch.int32)
            engine.init_active_device(se, t(b.A), t(b.bupper), t(b.blower), ms=b.ms, **{kind: t(d[key])})
            torch.cuda.synchronize()
            np.testing.assert_array_equal(se.cpu().numpy(), d[f"sense_{key}{tag}"], err_msg=f"{name} {key}{tag} (device)")


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_warm_started_solves_match_reference(engine, name):
    """solve_batch(primal_start= / dual_start=) -- the batched daqp.solve(..., primal_start=, dual_start=) of
    daqp.pyx:24-38 -- reproduces the reference's iteration counts: 1 from the optimum."""
    b, d = load(name)
    for key, kw in (("x_opt", "primal_start"), ("lam_opt", "dual_start"), ("x_out", "primal_start")):
        r = engine.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, None, ms=b.ms, **{kw: d[key]})
        np.testing.assert_array_equal(r.exitflag, d[f"flag_{key}"], err_msg=f"{name} {key}")
        np.testing.assert_array_equal(r.iter, d[f"iter_{key}"], err_msg=f"{name} {key}")
        assert np.abs(r.x - d[f"x_{key}"]).max() <= 1e-9 * (1 + np.abs(d[f"x_{key}"]).max())
    assert (r.iter >= 1).all()


@pytest.mark.gpu
def test_drop_in_init_symbols_and_device_loop(cuda_lib, engine):
    import torch
    import daqp_b200
    b, d = load("warmstart_n10_m30_ms3")
    # the drop-in single-problem symbols behind daqp.solve(primal_start=, dual_start=)
    for p in range(4):
        for kw, key in (("primal_start", "x_opt"), ("dual_start", "lam_opt")):
            x, fval, flag, info = daqp_b200.solve(b.H[p], b.f[p], b.A[p], b.bupper[p], b.blower[p], **{kw: d[key][p]})
            assert flag == 1 and info["iterations"] == 1
            assert np.abs(x - d["x_opt"][p]).max() < 1e-9 * (1 + np.abs(d["x_opt"][p]).max())
    # closed loop on the device: solve -> init_active(lam) -> solve again starts at the optimum (one iteration)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    H, f, A, bu, bl = t(b.H), t(b.f), t(b.A), t(b.bupper), t(b.blower)
    out = engine.solve_batch_device(H, f, A, bu, bl, None, ms=b.ms)
    sense = torch.zeros(b.bupper.shape, dtype=torch.int32, device=dev)
    engine.init_active_device(sense, A, bu, bl, lam=out["lam"], ms=b.ms)
    out2 = engine.solve_batch_device(H, f, A, bu, bl, sense, ms=b.ms)
    torch.cuda.synchronize()
    assert (out2["exitflag"] == 1).all() and (out2["iter"] == 1).all()
    assert (out2["x"] - out["x"]).abs().max().item() < 1e-9


@pytest.mark.gpu
def test_init_active_at_scale(engine):
    """C3 shape, 20 000 problems: every bit equal to a numpy restatement evaluated with the same left-to-right sums is
    too slow in Python, so check the size-independent property instead -- bits from the solver's own lam select exactly
    the rows with a non-zero multiplier, and bits from its x contain them."""
    from daqp_b200.problems import generate_g1
    b = generate_g1(20000, 50, 150, 0, 40, seed=77)
    r = engine.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, None, ms=0)
    sl = engine.init_active_batch(b.A, b.bupper, b.blower, None, lam=r.lam)
    np.testing.assert_array_equal((sl & 1) != 0, np.abs(r.lam) > 1e-12)
    np.testing.assert_array_equal((sl & 2) != 0, r.lam < -1e-12)
    sx = engine.init_active_batch(b.A, b.bupper, b.blower, None, x=r.x)
    ax = np.einsum("bmn,bn->bm", b.A, r.x)
    near = np.minimum(np.abs(ax - b.bupper), np.abs(ax - b.blower))
    assert ((sx & 1) != 0)[near < 5e-10].all() and not ((sx & 1) != 0)[near > 2e-9].any()


@pytest.mark.gpu
def test_first_violating_matches_reference_rule(cuda_lib):
    """daqp_first_violating (reference src/api.c:562-574) through the drop-in symbol and the batched entry: the first
    index with x_i or A_i x outside [bl - tol, bu + tol], simple bounds first, else m -- checked against the same loop in
    numpy (same left-to-right accumulation)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    n, m, ms, N = 7, 40, 3, 257
    A = rng.standard_normal((m - ms, n)); bu = rng.random(m) + 0.5; bl = -bu
    X = rng.standard_normal((N, n)) * np.linspace(0.01, 1.0, N)[:, None]
    want = np.full(N, m, np.intc)
    for p in range(N):
        for i in range(m):
            if i < ms: v = X[p, i]
            else:
                v = 0.0
                for j in range(n): v += A[i - ms, j] * X[p, j]
            if v > bu[i] + 1e-9 or v < bl[i] - 1e-9:
                want[p] = i
                break
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    got = np.zeros(N, np.intc)
    cuda_lib.daqp_b200_first_violating_batch.restype = C.c_int
    assert cuda_lib.daqp_b200_first_violating_batch(None, N, n, m, ms, dp(X), dp(A), dp(bu), dp(bl), C.c_double(1e-9),
                                                    got.ctypes.data_as(C.POINTER(C.c_int))) == 0
    np.testing.assert_array_equal(got, want)
    assert (want < m).any() and (want == m).any()
    cuda_lib.daqp_first_violating.restype = C.c_int
    for p in (0, 100, 256):
        x = np.ascontiguousarray(X[p])
        assert cuda_lib.daqp_first_violating(dp(x), dp(A), dp(bu), dp(bl), n, m, ms, C.c_double(1e-9)) == want[p]

"""GPU suite: the reference's OWN Python interface bound to this library.

oracle/_ref/pybind/daqp*.so is interfaces/daqp-python/daqp.pyx + daqp.pxd of the reference, unmodified, cythonized and
compiled against the reference's headers but LINKED AGAINST daqp_b200/libdaqp_b200.so instead of the reference's C
sources (oracle/Makefile target `pybind`, built in the container where /root/reference exists; the binary travels to
the GPU box). Every call below therefore goes daqp.solve / daqp.Model (reference Cython) -> daqp_quadprog, setup_daqp_main,
daqp_solve, daqp_update_ldp, daqp_primal/dual_init_active, daqp_set_primal_start, allocate_daqp_settings,
free_daqp_workspace, free_daqp_ldp, daqp_minrep (this library, CUDA).

The cases restate the in-scope tests of the reference's interfaces/daqp-python/test/example_test.py (line numbers in each
docstring); AVI, proximal-point and hierarchy tests are out of this repository's scope."""
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PYBIND = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "pybind")


@pytest.fixture(scope="module")
def daqp(cuda_lib):
    if not glob.glob(os.path.join(PYBIND, "daqp*.so")):
        pytest.skip("oracle/_ref/pybind not built (needs /root/reference at build time)")
    sys.path.insert(0, PYBIND)
    try:
        import daqp as mod
    finally:
        sys.path.remove(PYBIND)
    assert os.path.dirname(mod.__file__) == PYBIND
    return mod


def demo_qp():
    H = np.eye(2); f = np.array([1.0, 1.0]); A = np.array([[1.0, 1.0], [1.0, -1.0]])
    return H, f, A, np.array([1.0, 2, 3, 4]), -np.array([1.0, 2, 3, 4]), np.zeros(4, np.intc)


def model_qp():  # example_test.py:175-183
    return (np.eye(2), np.array([2.0, 2.0]), np.eye(2), np.array([1.0, 1.0]), np.array([-1.0, -1.0]), np.zeros(2, np.intc))


def warm_qp():  # example_test.py:95-100
    return (np.eye(2), np.array([1.0, 1.0]), np.array([[1.0, 1.0]]), np.array([2.0, 2.0, 10.0]),
            np.array([-2.0, -2.0, -1.0]), np.zeros(3, np.intc))


def test_python_demo(daqp):
    """example_test.py:17-26"""
    x, fval, exitflag, info = daqp.solve(*demo_qp())
    assert exitflag == 1 and info["iterations"] >= 1 and info["nodes"] == 1
    np.testing.assert_allclose(x, [-1.0, -1.0], atol=1e-9)  # the unconstrained optimum: feasible, on the bound x1 >= -1


@pytest.mark.parametrize("kind", ["dual", "primal"])
def test_warm_start(daqp, kind):
    """example_test.py:93-148: same optimum, no more iterations than the cold start, sense not mutated."""
    H, f, A, bu, bl, sense = warm_qp()
    x_cold, fval_cold, ef_cold, info_cold = daqp.solve(H, f, A, bu, bl, sense)
    assert ef_cold == 1
    before = sense.copy()
    kw = {"dual_start": info_cold["lam"]} if kind == "dual" else {"primal_start": x_cold}
    x_warm, fval_warm, ef_warm, info_warm = daqp.solve(H, f, A, bu, bl, sense, **kw)
    assert ef_warm == 1
    np.testing.assert_allclose(x_warm, x_cold, atol=1e-8)
    np.testing.assert_allclose(fval_warm, fval_cold, atol=1e-8)
    assert info_warm["iterations"] <= info_cold["iterations"]
    np.testing.assert_array_equal(sense, before)


def test_model_setup_solve_update(daqp):
    """example_test.py:185-237: setup + solve -> [-1,-1]; update(f) -> [1,1]; update(bounds) -> [-0.5,-0.5]."""
    H, f, A, bu, bl, sense = model_qp()
    d = daqp.Model()
    exitflag, setup_time = d.setup(H, f, A, bu, bl, sense)
    assert exitflag >= 0
    x, fval, ef, info = d.solve()
    assert ef == 1
    np.testing.assert_allclose(x, [-1.0, -1.0], atol=1e-6)
    assert d.update(f=np.array([-2.0, -2.0])) == 0
    x2, _, ef2, _ = d.solve()
    assert ef2 == 1
    np.testing.assert_allclose(x2, [1.0, 1.0], atol=1e-6)
    d = daqp.Model()
    d.setup(H, f, A, bu, bl, sense)
    d.update(bupper=np.array([0.5, 0.5]), blower=np.array([-0.5, -0.5]))
    x3, _, ef3, _ = d.solve()
    assert ef3 == 1
    np.testing.assert_allclose(x3, [-0.5, -0.5], atol=1e-6)


def test_model_guards_and_copies(daqp):
    """example_test.py:195-205,267-276"""
    d = daqp.Model()
    with pytest.raises(RuntimeError):
        d.solve()
    with pytest.raises(RuntimeError):
        d.update(f=np.array([1.0, 1.0]))
    d.setup(*model_qp())
    x1, _, _, _ = d.solve()
    x1[0] = 999.0
    x2, _, _, _ = d.solve()
    assert x2[0] != 999.0


def test_model_keeps_equalities(daqp):
    """example_test.py:239-265: six equalities among eleven rows, persistent model; update of the bounds returns 0."""
    n, neq = 10, 6
    H = np.eye(n); f = np.zeros(n)
    A = np.vstack((np.eye(n)[:neq], np.ones((1, n))))
    bu = np.concatenate((np.full(n, 10.0), np.zeros(neq), [10.0]))
    bl = np.concatenate((np.full(n, -10.0), np.zeros(neq), [-10.0]))
    sense = np.concatenate((np.zeros(n, np.intc), np.full(neq, 5, np.intc), np.zeros(1, np.intc)))
    d = daqp.Model()
    flag, _ = d.setup(H, f, A, bu, bl, sense)
    assert flag > 0
    x, _, exitflag, _ = d.solve()
    assert exitflag == 1
    np.testing.assert_allclose(x, np.zeros(n), atol=1e-8)
    bu[-1] = 9.0; bl[-1] = -9.0
    assert d.update(bupper=bu, blower=bl) == 0


def test_model_settings(daqp):
    """example_test.py:278-327: getter / setter, settings survive setup and a second setup."""
    d = daqp.Model()
    s = d.settings
    assert {"iter_limit", "primal_tol", "eps_prox", "time_limit"} <= set(s) and s["time_limit"] == 0.0
    tol = s["primal_tol"]
    d.settings = {"iter_limit": 42, "time_limit": 5.0}
    assert d.settings["iter_limit"] == 42 and d.settings["time_limit"] == 5.0 and d.settings["primal_tol"] == tol
    H, f, A, bu, bl, sense = model_qp()
    d = daqp.Model()
    d.settings = {"iter_limit": 123}
    d.setup(H, f, A, bu, bl, sense)
    assert d.settings["iter_limit"] == 123
    d.settings = {"iter_limit": 77}
    d.setup(H, np.array([-2.0, -2.0]), A, bu, bl, sense)
    assert d.settings["iter_limit"] == 77
    x, _, ef, _ = d.solve()
    assert ef == 1
    np.testing.assert_allclose(x, [1.0, 1.0], atol=1e-6)


def test_model_matches_quadprog_and_warm_start(daqp):
    """example_test.py:329-381"""
    H, f, A, bu, bl, sense = model_qp()
    x_ref, fval_ref, ef_ref, _ = daqp.solve(H, f, A, bu, bl, sense)
    d = daqp.Model()
    d.setup(H, f, A, bu, bl, sense)
    x_ws, fval_ws, ef_ws, info_cold = d.solve()
    assert ef_ref == 1 and ef_ws == 1
    np.testing.assert_allclose(x_ws, x_ref, atol=1e-8)
    np.testing.assert_allclose(fval_ws, fval_ref, atol=1e-8)
    before = sense.copy()
    d2 = daqp.Model()
    d2.setup(H, f, A, bu, bl, sense, dual_start=info_cold["lam"])
    x_warm, _, ef_warm, info_warm = d2.solve()
    assert ef_warm == 1 and info_warm["iterations"] <= info_cold["iterations"]
    np.testing.assert_allclose(x_warm, x_ws, atol=1e-8)
    d3 = daqp.Model()
    d3.setup(H, f, A, bu, bl, sense, primal_start=x_ws)
    np.testing.assert_array_equal(sense, before)
    x3, _, ef3, _ = d3.solve()
    assert ef3 == 1
    np.testing.assert_allclose(x3, x_ws, atol=1e-8)


def test_zero_eps_rejects_singular_hessian(daqp):
    """example_test.py:444-456: singular H with eps_prox = 0 => exit flag -5."""
    H = np.diag([1.0, 0.0]); f = np.ones(2); A = np.zeros((0, 2)); bu = np.ones(2)
    _, _, exitflag, _ = daqp.solve(H, f, A, bu, -bu, np.zeros(2, np.intc), eps_prox=0.0)
    assert exitflag == -5


def test_model_against_oracle_on_random_problems(daqp, oracle_libs):
    """Beyond the literals: the reference's Model driven through this library reproduces the oracle (x, iterations) on
    generated problems with simple bounds, across update(f, b) + warm solves (wsseq fixture of the reference itself)."""
    from daqp_b200.problems import generate_g1
    b = generate_g1(6, 12, 30, 3, 9, seed=4321)
    o = oracle_libs.OracleLib().solve(b)
    for p in range(b.N):
        x, fval, flag, info = daqp.solve(b.H[p], b.f[p], b.A[p], b.bupper[p], b.blower[p], np.zeros(b.m, np.intc))
        assert flag == o.exitflag[p] and info["iterations"] == o.iter[p]
        np.testing.assert_allclose(x, o.x[p], atol=1e-9 * (1 + np.abs(o.x[p]).max()))
        np.testing.assert_allclose(info["lam"], o.lam[p], atol=1e-7 * (1 + np.abs(o.lam[p]).max()))
    d = np.load(os.path.join(HERE, "golden", "wsseq_n10_m30_ms3.npz"))
    for p in range(4):
        m = daqp.Model()
        flag, _ = m.setup(d["H"][p], d["f"][p], d["A"][p], d["bupper"][p], d["blower"][p], np.zeros(int(d["m"]), np.intc))
        assert flag > 0
        x, _, ef, info = m.solve()
        assert ef == d["flag_0"][p] and info["iterations"] == d["iter_0"][p]
        np.testing.assert_allclose(x, d["x_0"][p], atol=1e-9 * (1 + np.abs(d["x_0"][p]).max()))
        live = True  # (after a solve that did not end optimal the next one restarts cold here: DESIGN.md §8)
        for k in range(int(d["K"])):
            assert m.update(f=d[f"f{k}"][p], bupper=d[f"bu{k}"][p], blower=d[f"bl{k}"][p]) in (0, -1)
            x, _, ef, info = m.solve()
            assert ef == d[f"flag_{k + 1}"][p]
            live = live and d[f"flag_{k}"][p] > 0
            if ef > 0 and live:
                assert info["iterations"] == d[f"iter_{k + 1}"][p]
                np.testing.assert_allclose(x, d[f"x_{k + 1}"][p], atol=1e-9 * (1 + np.abs(d[f"x_{k + 1}"][p]).max()))


def test_minrep_through_reference_binding(daqp):
    """daqp.minrep (daqp.pyx:636-652) -> daqp_minrep of this library, against the committed output of the reference."""
    d = np.load(os.path.join(HERE, "golden", "minrep_n3_m20.npz"))
    A, b, want = d["A"], d["b"], d["is_redundant"]
    for q in range(min(4, A.shape[0])):
        got = daqp.minrep(np.ascontiguousarray(A[q]), np.ascontiguousarray(b[q]))
        np.testing.assert_array_equal(np.asarray(got).astype(int), want[q].astype(int))

"""CPU suite: the oracle restatement against (1) golden vectors produced by the unmodified reference,
(2) the reference itself when it is compiled here (oracle/_ref), (3) construction-known optima."""
import numpy as np
import pytest

from common import (GOLDEN_DIR, RARE_MUST_HIT, assert_parity, bnb_golden_names, golden_names, kkt_residuals, ldp_golden_names,
                    load_golden, load_ldp_golden, rare_golden_names, rare_settings, rawldp_golden_names, ws_sets)
from daqp_b200.problems import generate_g0, generate_g1


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(oracle_libs, name):
    b, d = load_golden(name)
    o = oracle_libs.OracleLib().solve(b, use_sense=bool(d["use_sense"]))
    assert_parity(d["x"], d["lam"], d["fval"], d["exitflag"], d["iter"], o.x, o.lam, o.fval, o.exitflag, o.iter, name)
    started = d["exitflag"] >= -4  # failures raised by the setup leave no working set
    got = [sorted(w) for w in o.ws]
    want = ws_sets(d["ws"], d["n_active"])
    for p in np.nonzero(started)[0]:
        assert got[p] == want[p], f"{name}[{p}]: working sets differ"


@pytest.mark.parametrize("name", rare_golden_names())
def test_rare_paths_oracle_matches_reference_bit_for_bit(oracle_libs, name):
    """The rare control paths of daqp_ldp -- pivot_last, refactor-on-exit, refinement, the cycle guard with its repair
    and EXIT_CYCLE (daqp.c:28-85, auxiliary.c:379-396,498-593), NONCONVEX (utils.c:246-377), zero rows
    (utils.c:595-606) -- against outputs of the reference built without reassociation: x, fval, iterations, exit flags,
    working sets and (for solved problems) lam EQUAL; and the counters prove the paths were taken. lam of a problem that
    ends in EXIT_CYCLE is not compared: the reference reports the entering row's entry from a multiplier buffer it has
    not written yet (api.c:463-466 after the pointer swap of auxiliary.c:159-160)."""
    from oracle import harness
    b, d = load_golden(name)
    over = rare_settings(d)
    st = harness.default_settings(**over) if over else None
    o = oracle_libs.OracleLib().solve(b, settings=st, use_sense=bool(d["use_sense"]))
    for key in ("x", "fval", "iter", "exitflag"):
        np.testing.assert_array_equal(getattr(o, key), d[key], err_msg=f"{name}: {key}")
    ok = d["exitflag"] > 0
    np.testing.assert_array_equal(o.lam[ok], d["lam"][ok])
    want = [list(w[:k]) for w, k in zip(d["ws"], d["n_active"])]
    for p in np.nonzero(d["exitflag"] >= -4)[0]:
        assert list(o.ws[p]) == want[p], f"{name}[{p}]: working set (factor order)"
    np.testing.assert_array_equal(o.counts, d["counts"])
    for c in RARE_MUST_HIT.get(name, ()):
        assert o.counts[:, c].sum() > 0, f"{name}: path counter {c} never fired"
    if name == "rare_nonconvex":
        assert (d["exitflag"] == -5).sum() == 8 and (d["exitflag"] == 1).sum() == 4
    if name == "rare_zero_rows":
        assert (d["exitflag"][0::2] == 1).all() and (d["exitflag"][1::2] == -1).all()
    if name in ("rare_cycle_guard", "rare_cycle_exit", "rare_eqpairs_n50"):
        assert (d["exitflag"] == -2).any(), "EXIT_CYCLE not reached"  # the last one with the default settings


def test_rare_paths_live_reference(oracle_libs):
    """The same families on fresh seeds against the live strict build (this container only)."""
    if not oracle_libs.have_ref("libdaqp_ref_strict.so"):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("mgr", os.path.join(os.path.dirname(__file__), "golden", "make_golden_rare.py"))
    mgr = importlib.util.module_from_spec(spec); spec.loader.exec_module(mgr)
    from oracle import harness
    ref = oracle_libs.RefLib("libdaqp_ref_strict.so")
    orc = oracle_libs.OracleLib()
    hits = np.zeros(8, np.int64)
    for seed, eps in [(501, 3e-5), (502, 2e-5), (503, 1e-4)]:
        b = mgr.near_dependent_equalities(40, 16, 50, 2, 12, eps, seed)
        r = ref.solve(b, want_ws=True); o = orc.solve(b)
        for a, c in ((r.x, o.x), (r.fval, o.fval), (r.iter, o.iter), (r.exitflag, o.exitflag)):
            np.testing.assert_array_equal(a, c)
        assert all(list(a) == list(c) for a, c, f in zip(r.ws, o.ws, r.exitflag) if f >= -4)
        hits += o.counts.sum(axis=0)
    st = harness.default_settings(progress_tol=1.0, cycle_tol=4)
    b = generate_g1(40, 16, 50, 2, 12, seed=504)
    r = ref.solve(b, settings=st, want_ws=True); o = orc.solve(b, settings=st)
    for a, c in ((r.x, o.x), (r.fval, o.fval), (r.iter, o.iter), (r.exitflag, o.exitflag)):
        np.testing.assert_array_equal(a, c)
    hits += o.counts.sum(axis=0)
    assert (hits[4:] > 0).all(), f"rare path counters {hits[4:]}"


def test_known_answers():
    """Literals of the reference's own tests (example_test.py:175-237, 00_basic_qp.cpp:28)."""
    for name, want in [("lit_model_qp", [-1, -1]), ("lit_model_qp_flipped", [1, 1]), ("lit_model_qp_half", [-0.5, -0.5]),
                       ("lit_eigen_basic", [-1, -1])]:
        _, d = load_golden(name)
        assert d["exitflag"][0] == 1
        np.testing.assert_allclose(d["x"][0], want, atol=1e-6)
    assert load_golden("lit_python_demo")[1]["exitflag"][0] == 1
    assert (load_golden("warm_exact")[1]["iter"] == 1).all()
    assert (load_golden("trivially_infeasible")[1]["exitflag"] == -1).all()


@pytest.mark.parametrize("shape", [(10, 20, 0, 8), (20, 60, 0, 16), (24, 70, 9, 18), (50, 150, 0, 40)])
def test_oracle_recovers_constructed_optimum(oracle_libs, shape):
    """generate_test_QP gives xref and the optimal active set by construction; the reference's own gate is
    |x - xref| < 1e-4 (core_tests.jl:26-30)."""
    n, m, ms, na = shape
    b = generate_g1(40, n, m, ms, na, seed=900 + n)
    o = oracle_libs.OracleLib().solve(b)
    assert (o.exitflag == 1).all()
    assert np.abs(o.x - b.xref).max() < 1e-8
    for p in range(b.N):
        assert sorted(o.ws[p]) == np.nonzero(b.active_ref[p])[0].tolist()
        assert (np.sign(o.lam[p]) == b.active_ref[p]).all()
    stat, pf, comp = kkt_residuals(b, o.x, o.lam)
    assert stat.max() < 1e-8 and pf.max() < 1e-8 and comp.max() < 1e-8


@pytest.mark.parametrize("libname", ["libdaqp_ref_strict.so", "libdaqp_ref.so"])
def test_oracle_vs_live_reference(oracle_libs, libname):
    """Against the reference compiled here from /root/reference. The -O2 -ffp-contract=off build must agree BIT FOR
    BIT; the default fast-math build to the parity tolerances with identical iteration counts and working sets."""
    if not oracle_libs.have_ref(libname):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = oracle_libs.RefLib(libname)
    orc = oracle_libs.OracleLib()
    batches = [generate_g1(60, 10, 20, 0, 8, seed=1), generate_g1(60, 20, 60, 10, 16, seed=2),
               generate_g1(20, 50, 150, 0, 40, seed=3), generate_g0(40, 20, 60, seed=4),
               generate_g1(40, 20, 60, 5, 16, kappa=1e9, seed=5), generate_g1(60, 10, 40, 10, 10, seed=6)]
    for b in batches:
        r = ref.solve(b, want_ws=True)
        o = orc.solve(b)
        if "strict" in libname:
            for a, c in ((r.x, o.x), (r.lam, o.lam), (r.fval, o.fval), (r.iter, o.iter), (r.exitflag, o.exitflag)):
                np.testing.assert_array_equal(a, c)
        else:
            assert_parity(r.x, r.lam, r.fval, r.exitflag, r.iter, o.x, o.lam, o.fval, o.exitflag, o.iter, libname)
        assert [list(w) for w in r.ws] == [list(w) for w in o.ws]


def test_oracle_threads_and_packed_entry(oracle_libs):
    b = generate_g1(64, 12, 30, 3, 9, seed=11)
    orc = oracle_libs.OracleLib()
    one = orc.solve(b)
    for nt in (1, 3):
        pk = orc.solve_packed(b, nthreads=nt)
        np.testing.assert_array_equal(one.x, pk.x)
        np.testing.assert_array_equal(one.iter, pk.iter)
        np.testing.assert_array_equal(one.counts, pk.counts)


def test_oracle_out_of_scope_flags(oracle_libs):
    b = generate_g1(4, 8, 20, 0, 6, seed=12)
    b.sense[:, 0] = 16  # binary
    assert (oracle_libs.OracleLib().solve(b, use_sense=True).exitflag == -8).all()
    b = generate_g1(4, 8, 20, 0, 6, seed=13)
    b.H[:, 0, :] = 0; b.H[:, :, 0] = 0  # singular H -> proximal driver in the reference
    assert (oracle_libs.OracleLib().solve(b).exitflag == -8).all()


def test_oracle_soft_constraints_vs_live_reference(oracle_libs):
    """Soft constraints (sense & 8: +rho_soft on the pivot, n + ns + 1 factor rows, soft_slack, exit flag 2;
    factorization.c:48-52, auxiliary.c:69-84,534-535, daqp.c:59-62): bit for bit against the strict reference build."""
    if not oracle_libs.have_ref("libdaqp_ref_strict.so"):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    from daqp_b200.problems import soften
    ref = oracle_libs.RefLib("libdaqp_ref_strict.so")
    orc = oracle_libs.OracleLib()
    seen = set()
    for cfg, frac, shift in [((60, 10, 20, 0, 8), 0.2, 0.5), ((40, 20, 60, 5, 16), 0.5, 2.0), ((30, 12, 40, 12, 10), 1.0, 1.0),
                             ((8, 50, 150, 0, 40), 0.2, 0.5)]:
        b = soften(generate_g1(*cfg, seed=77), frac, shift, 5)
        r = ref.solve(b, want_ws=True, use_sense=True)
        o = orc.solve(b, use_sense=True)
        for a, c in ((r.x, o.x), (r.lam, o.lam), (r.fval, o.fval), (r.iter, o.iter), (r.exitflag, o.exitflag),
                     (r.soft_slack, o.soft_slack)):
            np.testing.assert_array_equal(a, c)
        assert [list(w) for w in r.ws] == [list(w) for w in o.ws]
        seen |= set(np.unique(o.exitflag).tolist())
    assert {1, 2} <= seen


def test_fp32_oracle_vs_live_fp32_reference(oracle_libs):
    """c_float = float: the oracle compiled with -DORC_SINGLE against the reference compiled with
    -DDAQP_SINGLE_PRECISION (include/types.h:8-12): same exit flags and iteration counts, x within fp32 accuracy."""
    if not oracle_libs.have_ref("libdaqp_ref_f32.so"):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = oracle_libs.RefLib("libdaqp_ref_f32.so")
    orc = oracle_libs.OracleLib(single=True)
    for cfg in [(200, 10, 20, 0, 8), (100, 20, 60, 4, 16), (30, 50, 150, 0, 40)]:
        b = generate_g1(*cfg, seed=3200 + cfg[1])
        r, o = ref.solve(b), orc.solve(b)
        np.testing.assert_array_equal(r.exitflag, o.exitflag)
        assert (r.iter == o.iter).mean() >= 0.9  # near-ties resolve differently under fast-math in fp32: a rate, not bit parity
        assert np.abs(r.x - o.x).max() <= 1e-4 * (1 + np.abs(r.x).max())


def test_oracle_decision_log(oracle_libs):
    """OracleLib.solve(log_cap=...) records every working-set decision (the log tests/trace_diff.py compares with the
    kernel's DAQPB200Diag.trace): as many adds / removes as the path counters say, one objective value per add made by the
    feasibility scan, the exit flag last."""
    b = generate_g1(20, 12, 36, 4, 9, seed=77)
    o = oracle_libs.OracleLib().solve(b, log_cap=4096)
    for p in range(b.N):
        log = o.oplog[p]
        codes = log[:, 0]
        assert codes[-1] == 7 and log[-1, 1] == o.exitflag[p]
        assert (codes == 1).sum() == o.counts[p, 1] and (codes == 2).sum() == o.counts[p, 2]
        assert (codes == 8).sum() == (codes == 9).sum() == (codes == 1).sum()  # no warm start: every add comes from a scan
        adds = log[codes == 1, 1] >> 1
        assert ((adds >= 0) & (adds < b.m)).all()
        live = set()
        for c, v in log:  # the log replays to the final working set
            if c == 1: live.add(int(v) >> 1)
            elif c == 2: live.discard(int(v))
        assert live == set(o.ws[p])


@pytest.mark.parametrize("name", ldp_golden_names())
def test_pure_ldp_oracle_matches_reference(oracle_libs, name):
    """H == NULL, f == NULL (the LDP min |x|^2 itself, utils.c:103-110): the oracle -- which walks it with an identity
    Hessian -- reproduces the reference's exit flags, iteration counts and working sets, x / lam to 1e-9 of the default
    build's, and is BIT-identical to the live strict build (with H = NULL, and with H = I: the substitution is exact)."""
    b, d = load_ldp_golden(name)
    use_sense = bool(d["use_sense"])
    o = oracle_libs.OracleLib().solve(b, use_sense=use_sense, null_H=True)
    np.testing.assert_array_equal(o.exitflag, d["exitflag"])
    np.testing.assert_array_equal(o.iter, d["iter"])
    ok = d["exitflag"] > 0
    assert ok.any()
    np.testing.assert_allclose(o.x[ok], d["x"][ok], atol=1e-9)
    np.testing.assert_allclose(o.lam[ok], d["lam"][ok], atol=1e-7)
    for p in range(b.N):
        assert list(o.ws[p]) == list(d["ws"][p, :d["n_active"][p]]), f"{name}[{p}]: working set"
    if oracle_libs.have_ref("libdaqp_ref_strict.so"):
        ref = oracle_libs.RefLib("libdaqp_ref_strict.so")
        for null_H in (True, False):
            r = ref.solve(b, use_sense=use_sense, null_H=null_H)
            np.testing.assert_array_equal(r.exitflag, o.exitflag)
            np.testing.assert_array_equal(r.iter, o.iter)
            np.testing.assert_array_equal(r.x[ok], o.x[ok])
            np.testing.assert_array_equal(r.lam[ok], o.lam[ok])


@pytest.mark.parametrize("name", rawldp_golden_names())
def test_raw_ldp_fixtures_are_what_the_reference_gives(oracle_libs, name):
    """daqp_ldp on hand-filled workspaces (api.jl:428-459): where the reference is compiled here, its live output equals
    the file; every optimal u satisfies the polyhedron it was projected onto, and the radius fixture has both outcomes."""
    import ctypes as C
    import os
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, m, ms = int(d["n"]), int(d["m"]), int(d["ms"])
    fvb = None if d["fval_bound"] < 0 else float(d["fval_bound"])
    assert (d["exitflag"] == 1).any() and (d["exitflag"] == -1).any()
    for p in np.nonzero(d["exitflag"] == 1)[0]:
        Ax = np.concatenate([d["u"][p, :ms], d["A"][p] @ d["u"][p]])
        scale = np.concatenate([np.ones(ms), np.linalg.norm(d["A"][p], axis=1)])
        on = (d["sense"][p] & 4) == 0  # rows the caller switched off (IMMUTABLE without ACTIVE) do not constrain u
        assert (Ax <= d["bupper"][p] + 1e-5 * scale)[on].all() and (Ax >= d["blower"][p] - 1e-5 * scale)[on].all()
        assert abs(d["fval"][p] - d["u"][p] @ d["u"][p]) < 1e-9 * (1 + d["fval"][p])
    if oracle_libs.have_ref():
        L = C.CDLL(os.path.join(oracle_libs.REF_DIR, "libdaqp_ref.so"))
        for p in range(d["exitflag"].shape[0]):
            se = d["sense"][p] if d["sense"].any() else None
            r = oracle_libs.raw_ldp(L, d["A"][p], d["bupper"][p], d["blower"][p], se, ms, fvb)
            assert r["exitflag"] == d["exitflag"][p] and r["iter"] == d["iter"][p]
            assert r["ws"] == list(d["ws"][p, :d["n_active"][p]])
            np.testing.assert_array_equal(r["u"], d["u"][p])


@pytest.mark.parametrize("name", bnb_golden_names())
def test_bnb_fixtures_are_what_the_reference_gives(oracle_libs, name):
    """Branch-and-bound fixtures (reference src/bnb.c through daqp_quadprog): every binary constraint sits on one of its
    bounds at the recorded optimum, the literals of core_tests.jl:150-178 have their known answers, and -- where the
    reference is compiled here -- the live reference reproduces the file."""
    b, d = load_golden(name)
    assert (d["exitflag"] == 1).all()
    for p in range(b.N):
        for i in np.nonzero(b.sense[p] & 16)[0]:
            val = d["x"][p, i] if i < b.ms else b.A[p, i - b.ms] @ d["x"][p]
            assert min(abs(val - b.bupper[p, i]), abs(val - b.blower[p, i])) < 1e-5
    if name == "bnb_lit_x011":
        np.testing.assert_allclose(d["x"][0], [0, 1, 1], atol=1e-6)
    if name.startswith("bnb_lit_zero_dual"):
        np.testing.assert_allclose(d["x"][0], 0, atol=1e-9)
    if oracle_libs.have_ref():
        r = oracle_libs.RefLib("libdaqp_ref.so").solve(b, use_sense=True)
        np.testing.assert_array_equal(r.exitflag, d["exitflag"])
        np.testing.assert_allclose(r.x, d["x"], atol=1e-12)

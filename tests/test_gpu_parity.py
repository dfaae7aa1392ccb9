"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed golden vectors.

fp64 tolerances are stated in tests/common.py (x 1e-9, fval 1e-9, lam 1e-7, relative); exit flags, iteration counts,
final working sets and per-problem operation counts must be EQUAL.
"""
import os

import numpy as np
import pytest

from common import (GOLDEN_DIR, RARE_MUST_HIT, assert_parity, bnb_golden_names, golden_names, kkt_residuals, ldp_golden_names,
                    load_golden, load_ldp_golden, rare_golden_names, rare_settings, rawldp_golden_names, ws_sets)
from daqp_b200.problems import generate_config, generate_g0, generate_g1, soften

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(cuda_lib):
    import daqp_b200
    e = daqp_b200.Engine()
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracle(oracle_libs):
    return oracle_libs.OracleLib()


def run_gpu(engine, b, use_sense=None, **settings):
    if use_sense is None:
        use_sense = bool(b.sense.any())
    return engine.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, b.sense if use_sense else None, ms=b.ms, diag=True,
                              **settings)


def check_vs_oracle(engine, oracle, b, what, use_sense=None, settings=None, x_tol=None, f_tol=None):
    from oracle import harness
    settings = settings or {}
    st = harness.default_settings(**settings) if settings else None
    o = oracle.solve(b, settings=st, use_sense=use_sense)
    r = run_gpu(engine, b, use_sense=use_sense, **settings)
    assert_parity(o.x, o.lam, o.fval, o.exitflag, o.iter, r.x, r.lam, r.fval, r.exitflag, r.iter, what, x_tol=x_tol,
                  f_tol=f_tol)
    started = o.exitflag >= -4
    got = r.working_sets()
    for p in np.nonzero(started)[0]:
        assert got[p] == list(o.ws[p]), f"{what}[{p}]: working set (factor order) differs"
        assert (r.sense[p] & 3 == o.sense[p] & 3)[o.ws[p]].all(), f"{what}[{p}]: active side differs"
    np.testing.assert_array_equal(r.counts[started], o.counts[started], err_msg=f"{what}: scan/add/remove counts")
    return o, r


@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_golden(engine, name):
    """Outputs of the unmodified reference, generated in the build container (tests/golden/make_golden.py)."""
    b, d = load_golden(name)
    r = run_gpu(engine, b, use_sense=bool(d["use_sense"]))
    assert_parity(d["x"], d["lam"], d["fval"], d["exitflag"], d["iter"], r.x, r.lam, r.fval, r.exitflag, r.iter, name)
    want = ws_sets(d["ws"], d["n_active"])
    got = [sorted(w) for w in r.working_sets()]
    for p in np.nonzero(d["exitflag"] >= -4)[0]:
        assert got[p] == want[p], f"{name}[{p}]: final active set differs from the reference"


@pytest.mark.parametrize("name", rare_golden_names())
def test_rare_paths_cuda_matches_reference(engine, name):
    """The rare control paths of daqp_ldp in the PLAIN instantiation of the solve kernel (and its team mode, n = 70):
    pivot_last, refactor-on-exit, refinement, cycle guard + repair + EXIT_CYCLE, NONCONVEX, zero rows. Reference =
    the build without reassociation (tests/golden/make_golden_rare.py). Exit flags, iteration counts, working sets in
    factor order and all eight path counters EQUAL; the counters prove the paths ran on the GPU. Values: these inputs are
    ill-conditioned on purpose (pivots ~1e-10), so x / lam are held to 1e-5 relative."""
    b, d = load_golden(name)
    over = rare_settings(d)
    r = run_gpu(engine, b, use_sense=bool(d["use_sense"]), **over)
    off = np.nonzero((r.iter != d["iter"]) | (r.exitflag != d["exitflag"]))[0]
    # The near-dependent-equality families sit on pivots of ~1e-10 computed by cancellation from O(1) terms: the path can
    # turn on the last bits of a sum. The fixtures keep only problems whose path the reference's own builds and few-ulp
    # input perturbations agree on; what is left may still part from a GPU summation order on an isolated problem, which
    # is tolerated up to 1 in 50 and REPORTED -- every other problem, and every problem of every other fixture, must
    # follow the reference's path exactly. The one such problem today (rare_eqpairs_n12_ms4[33]) was traced decision by
    # decision (tests/trace_diff.py, profiles/trace_diff_rare_eqpairs_r02.log): both sides make the SAME 38 first
    # decisions and enter the same 4-step cycle; with D ~ 1e-10 the objective is only reproducible to ~1e-11, and the
    # GPU's value drifts by +9e-12 during the first two periods, which the guard's 1e-14 progress test (daqp.c:67) counts
    # as progress -- so the repair fires two periods (8 iterations) later and then proceeds identically to the same exit.
    allowed = max(1, b.N // 50) if name.startswith("rare_eqpairs") else 0
    if off.size:
        print(f"\n{name}: path differs from the reference on problems {off.tolist()}: iterations {r.iter[off].tolist()} vs "
              f"{d['iter'][off].tolist()}, flags {r.exitflag[off].tolist()} vs {d['exitflag'][off].tolist()}")
    assert off.size <= allowed, f"{name}: path differs on {off.size} problems ({off.tolist()}), allowed {allowed}"
    np.testing.assert_array_equal(r.exitflag[off] > 0, d["exitflag"][off] > 0)  # ... and even those end the same way
    same = np.ones(b.N, bool); same[off] = False
    assert_parity(d["x"][same], d["lam"][same], d["fval"][same], d["exitflag"][same], d["iter"][same], r.x[same],
                  r.lam[same], r.fval[same], r.exitflag[same], r.iter[same], name, x_tol=1e-5, f_tol=1e-6)
    want = [list(w[:k]) for w, k in zip(d["ws"], d["n_active"])]
    got = r.working_sets()
    started = (d["exitflag"] >= -4) & same
    for p in np.nonzero(started)[0]:
        assert got[p] == want[p], f"{name}[{p}]: working set (factor order) differs"
    np.testing.assert_array_equal(r.counts[started], d["counts"][started], err_msg=f"{name}: path counters")
    for c in RARE_MUST_HIT.get(name, ()):
        assert r.counts[:, c].sum() > 0, f"{name}: path counter {c} never fired on the GPU"


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_configs_vs_oracle(engine, oracle, cfg):
    """BASELINE.json configs at sizes the scalar oracle finishes in seconds."""
    N = {"C1": 1000, "C2": 3000, "C3": 1200}[cfg]
    b = generate_config(cfg, N=N)
    o, r = check_vs_oracle(engine, oracle, b, cfg)
    assert (r.exitflag == 1).all()
    assert np.abs(r.x - b.xref).max() < 1e-8  # construction-known optimum


@pytest.mark.parametrize("shape", [(20, 60, 10, 16), (30, 80, 30, 20), (70, 200, 7, 50), (33, 95, 5, 25),
                                   (64, 130, 0, 50), (65, 131, 1, 50), (120, 400, 120, 96)])
def test_shapes_vs_oracle(engine, oracle, shape):
    """Simple bounds, odd sizes (padding paths), n on both sides of a 64-column lane group, C4 shape."""
    n, m, ms, na = shape
    N = 150 if n >= 100 else 400
    check_vs_oracle(engine, oracle, generate_g1(N, n, m, ms, na, seed=2000 + n + ms), f"G1{shape}")


@pytest.mark.parametrize("shape", [(50, 300, 0, 40), (20, 520, 5, 14), (33, 700, 0, 20), (63, 257, 10, 50), (40, 768, 0, 30),
                                   (30, 800, 0, 20)])
def test_many_rows_screening_in_blocks(engine, oracle, shape):
    """m > 256 in the single-warp kernel: the fp32 screening scan takes the rows through its ring in blocks of 128 (up to
    m = 768; beyond that the exact scan), same decisions as the oracle."""
    n, m, ms, na = shape
    check_vs_oracle(engine, oracle, generate_g1(150, n, m, ms, na, seed=8100 + m), f"G1{shape}")


def test_g0_distribution(engine, oracle):
    check_vs_oracle(engine, oracle, generate_g0(800, 50, 150), "G0 n50 m150")


def test_degenerate_and_infeasible(engine, oracle):
    check_vs_oracle(engine, oracle, generate_g1(600, 10, 40, 10, 10, seed=31), "vertex (nact = n)")
    # cond(H) = 1e9: flags / iterations / working sets must still be equal; values are compared at kappa * eps
    # (the reference itself is 2.5e-5 away from the constructed optimum on this set)
    check_vs_oracle(engine, oracle, generate_g1(400, 20, 60, 5, 16, kappa=1e9, seed=32), "kappa 1e9", x_tol=1e-6)
    b = generate_g1(400, 10, 30, 0, 8, seed=33)
    b.A[:, 15:30] = b.A[:, 0:15]; b.bupper[:, 15:30] = b.bupper[:, 0:15]; b.blower[:, 15:30] = b.blower[:, 0:15]
    check_vs_oracle(engine, oracle, b, "duplicate rows (singular steps)")
    b = generate_g1(400, 10, 30, 0, 8, seed=34)
    b.A[:, 1] = b.A[:, 0]; b.bupper[:, 1] = b.blower[:, 0] - 1.0; b.blower[:, 1] = b.blower[:, 0] - 2.0
    o, r = check_vs_oracle(engine, oracle, b, "infeasible")
    assert (r.exitflag == -1).all()
    b = generate_g1(64, 10, 30, 0, 8, seed=35)
    b.blower[::2, 3] = b.bupper[::2, 3] + 1  # every other problem trivially infeasible, the rest solvable
    o, r = check_vs_oracle(engine, oracle, b, "mixed trivially infeasible")
    assert (r.exitflag[::2] == -1).all() and (r.exitflag[1::2] == 1).all()


def test_warm_start_and_equalities(engine, oracle):
    b = generate_g1(300, 20, 60, 5, 16, seed=41)
    o = oracle.solve(b)
    b.sense[o.lam > 1e-12] = 1
    b.sense[o.lam < -1e-12] = 3
    o2, r2 = check_vs_oracle(engine, oracle, b, "warm start from the optimal active set", use_sense=True)
    assert (r2.iter == 1).all()  # core_tests.jl:520-543
    rng = np.random.default_rng(5)
    b.sense[:] = np.where(rng.random(b.sense.shape) < 0.6, rng.choice([1, 3], b.sense.shape), 0)
    check_vs_oracle(engine, oracle, b, "over-determined warm start", use_sense=True)
    b = generate_g1(300, 20, 60, 0, 16, seed=42)
    for p in range(b.N):
        act = np.nonzero(b.active_ref[p])[0][:4]
        b.sense[p, act] = 5
    check_vs_oracle(engine, oracle, b, "equalities", use_sense=True)
    b.sense[:] = 0
    for p in range(b.N):
        for i in np.nonzero(b.active_ref[p])[0][:4]:
            if b.active_ref[p, i] > 0: b.blower[p, i] = b.bupper[p, i]
            else: b.bupper[p, i] = b.blower[p, i]
    check_vs_oracle(engine, oracle, b, "equalities via bl == bu", use_sense=False)


@pytest.mark.parametrize("n,mode", [(9, None), (31, None), (32, None), (33, None), (63, None), (64, "0"), (64, "4"), (65, "0"),
                                    (65, "4"), (100, "0"), (100, "4"), (127, "0"), (127, "4")])
def test_warm_start_activation_across_shapes(engine, oracle, n, mode, monkeypatch):
    """The two-pass activation (Gram pass + right-looking LDL', single-warp kernel for n <= 63, team kernel above) against
    the oracle's row-by-row daqp_activate_constraints on both sides of every register-segment / team boundary: warm starts
    from a perturbed neighbour's active set (wrong rows in, right rows missing), from the exact set, with equality rows,
    with simple bounds, with fewer rows than the two-pass threshold, and with more rows than dimensions (falls back to the
    row-by-row path and its drop rules). Flags, iterations, working sets in factor order, path counters equal."""
    if mode is not None:  # n + 1 > 64: the single-warp kernel ("0") and the team kernel ("4")
        monkeypatch.setenv("DAQP_B200_TEAM", mode)
    m, ms, na = 3 * n, (n // 3 if n % 2 else 0), max(4, (7 * n) // 10)
    b = generate_g1(16, n, m, ms, na, seed=1200 + n)
    nb = generate_g1(16, n, m, ms, na, seed=1200 + n)
    rng = np.random.default_rng(n)
    nb.f = b.f * (1 + 0.05 * rng.standard_normal(b.f.shape))
    on = oracle.solve(nb)
    b.sense[on.lam > 1e-12] = 1
    b.sense[on.lam < -1e-12] = 3
    check_vs_oracle(engine, oracle, b, f"n={n}: neighbour's active set", use_sense=True)
    o = oracle.solve(generate_g1(16, n, m, ms, na, seed=1200 + n))
    b.sense[:] = 0
    b.sense[o.lam > 1e-12] = 1
    b.sense[o.lam < -1e-12] = 3
    _, r = check_vs_oracle(engine, oracle, b, f"n={n}: exact active set", use_sense=True)
    assert (r.iter == 1).all()
    for p in range(b.N):  # two of the active rows as equalities, the others as warm-start bits
        for i in np.nonzero(b.sense[p])[0][:2]:
            b.sense[p, i] |= 4
    check_vs_oracle(engine, oracle, b, f"n={n}: with equalities", use_sense=True)
    b.sense[:] = 0
    for p in range(b.N):  # five rows only: below the two-pass threshold
        act = np.nonzero(o.lam[p])[0][:5]
        b.sense[p, act] = np.where(o.lam[p, act] > 0, 1, 3)
    check_vs_oracle(engine, oracle, b, f"n={n}: five rows", use_sense=True)
    b.sense[:] = np.where(rng.random(b.sense.shape) < 0.5, rng.choice([1, 3], b.sense.shape), 0)  # ~1.5 n rows
    check_vs_oracle(engine, oracle, b, f"n={n}: more rows than dimensions", use_sense=True)


def test_soft_constraints(engine, oracle):
    """sense & 8 (reference factorization.c:48-52, auxiliary.c:69-84,534-535, daqp.c:59-62): the working set grows past
    n (up to n + ns rows of the factor), exit flag 2, soft_slack reported."""
    seen = set()
    for cfg, frac, shift in [((200, 10, 20, 0, 8), 0.2, 0.5), ((100, 20, 60, 5, 16), 0.5, 2.0),
                             ((60, 12, 40, 12, 10), 1.0, 1.0), ((24, 50, 150, 0, 40), 0.2, 0.5)]:
        b = soften(generate_g1(*cfg, seed=77), frac, shift, 5)
        # fval is dominated by soft_slack = rho_soft * sum lam*^2 here (1e5 against |u|^2 ~ 1): it inherits twice the
        # relative tolerance of lam (1e-7), so it is held to 1e-6 instead of 1e-9
        o, r = check_vs_oracle(engine, oracle, b, f"soft {cfg} {frac}", use_sense=True, f_tol=1e-6)
        np.testing.assert_allclose(r.soft_slack, o.soft_slack, rtol=1e-6, atol=1e-12)
        seen |= set(np.unique(r.exitflag).tolist())
    assert {1, 2} <= seen
    b = soften(generate_g1(3, 10, 20, 0, 8, seed=78), 0.5, 1.0, 6)  # drop-in symbol reports soft_slack too
    import daqp_b200
    o = oracle.solve(b, use_sense=True)
    for p in range(b.N):
        x, fval, flag, info = daqp_b200.solve(b.H[p], b.f[p], b.A[p], b.bupper[p], b.blower[p], b.sense[p])
        assert flag == o.exitflag[p] and info["iterations"] == o.iter[p]
        np.testing.assert_allclose(info["soft_slack"], o.soft_slack[p], rtol=1e-6, atol=1e-12)


def test_c4_mpc_step_warm_start(engine, oracle):
    """BASELINE.json config 4 shape (n=120, m=400, box on every variable + 280 general rows), warm-started the way an
    MPC loop does: sense bits pre-set from the optimal active set of a neighbour whose linear term differs by 5 %."""
    b = generate_g1(20, 120, 400, 120, 96, seed=404)
    nb = generate_g1(20, 120, 400, 120, 96, seed=404)
    rng = np.random.default_rng(44)
    nb.f = nb.f * (1 + 0.05 * rng.standard_normal(nb.f.shape))
    on = oracle.solve(nb)
    assert (on.exitflag == 1).all()
    b.sense[on.lam > 1e-12] = 1
    b.sense[on.lam < -1e-12] = 3
    cold = oracle.solve(b, use_sense=False)
    o, r = check_vs_oracle(engine, oracle, b, "C4 warm-started MPC step", use_sense=True)
    assert (r.exitflag == 1).all() and r.iter.mean() < 0.5 * cold.iter.mean()  # the warm start pays (287 -> 23)


def test_c4_at_scale_cold_and_warm(engine, oracle):
    """BASELINE.json config 4 (n=120, m=400, ms=120) at N = 2000, cold and warm-started from a neighbour's active set:
    the team mode of the solve kernel (one CTA of four warps per problem, n > 64). Exit flags, iteration counts,
    working sets in factor order, operation counts equal to the oracle's; x / lam / fval to the fp64 tolerances."""
    N = 2000
    b = generate_g1(N, 120, 400, 120, 96, seed=4404)
    o, r = check_vs_oracle(engine, oracle, b, "C4 cold, N=2000")
    assert (r.exitflag == 1).all() and np.abs(r.x - b.xref).max() < 1e-5  # construction-known optimum (kappa = 100)
    nb = generate_g1(N, 120, 400, 120, 96, seed=4404)
    rng = np.random.default_rng(45)
    nb.f = nb.f * (1 + 0.05 * rng.standard_normal(nb.f.shape))
    rn = run_gpu(engine, nb)  # the neighbour's optimum (checked against the oracle on a sample)
    on = oracle.solve(nb.slice(0, 100))
    np.testing.assert_array_equal(rn.iter[:100], on.iter)
    assert (rn.exitflag == 1).all()
    b.sense[rn.lam > 1e-12] = 1
    b.sense[rn.lam < -1e-12] = 3
    o2, r2 = check_vs_oracle(engine, oracle, b, "C4 warm, N=2000", use_sense=True)
    assert (r2.exitflag == 1).all() and r2.iter.mean() < 0.5 * r.iter.mean()


@pytest.mark.parametrize("shape", [(66, 140, 0, 50), (95, 300, 40, 70), (96, 200, 96, 80), (127, 260, 3, 100),
                                   (100, 600, 0, 75)])
@pytest.mark.parametrize("mode", ["4", "0"])
def test_team_mode_shapes(engine, oracle, shape, mode, monkeypatch):
    """Sizes on both sides of the team kernel's three / four row segments (n + 1 = 67 .. 128), with m beyond 512 -- through
    the team kernel (DAQP_B200_TEAM=4: wherever it can run) and through the single-warp kernel (=0); the library's own
    choice between the two is by measured speed (daqp_b200.cu) and takes one or the other of these paths."""
    monkeypatch.setenv("DAQP_B200_TEAM", mode)
    n, m, ms, na = shape
    check_vs_oracle(engine, oracle, generate_g1(120, n, m, ms, na, seed=7000 + n + ms), f"team={mode} G1{shape}")


@pytest.mark.parametrize("mode", ["4", "0"])
def test_team_mode_rare_paths(engine, oracle, mode, monkeypatch):
    """Singular steps, infeasibility, warm starts (incl. over-determined ones) and equalities at n > 64, in both kernels."""
    monkeypatch.setenv("DAQP_B200_TEAM", mode)
    b = generate_g1(100, 70, 160, 0, 50, seed=7101)
    b.A[:, 80:160] = b.A[:, 0:80]; b.bupper[:, 80:160] = b.bupper[:, 0:80]; b.blower[:, 80:160] = b.blower[:, 0:80]
    check_vs_oracle(engine, oracle, b, "team: duplicate rows")
    b = generate_g1(60, 80, 200, 10, 60, seed=7102)
    b.A[:, 1] = b.A[:, 0]; b.bupper[:, 1] = b.blower[:, 0] - 1.0; b.blower[:, 1] = b.blower[:, 0] - 2.0
    o, r = check_vs_oracle(engine, oracle, b, "team: infeasible")
    assert (r.exitflag == -1).all()
    b = generate_g1(80, 72, 180, 12, 72, seed=7103)
    # (a vertex of 72 constraints in 72 variables: the multipliers solve a square system whose conditioning sets their
    # accuracy -- 3.8e-7 relative between two fp64 summation orders; x is unaffected)
    check_vs_oracle(engine, oracle, b, "team: vertex (nact = n)", x_tol=1e-8)
    rng = np.random.default_rng(7)
    b.sense[:] = np.where(rng.random(b.sense.shape) < 0.5, rng.choice([1, 3], b.sense.shape), 0)
    check_vs_oracle(engine, oracle, b, "team: over-determined warm start", use_sense=True, x_tol=1e-8)
    b = generate_g1(80, 90, 220, 0, 60, seed=7104)
    for p in range(b.N):
        b.sense[p, np.nonzero(b.active_ref[p])[0][:5]] = 5
    check_vs_oracle(engine, oracle, b, "team: equalities", use_sense=True)
    check_vs_oracle(engine, oracle, generate_g1(60, 100, 250, 20, 80, kappa=1e8, seed=7105), "team: kappa 1e8", x_tol=1e-6)


def test_odd_batch_sizes_keep_staging_aligned(engine, oracle):
    """Host entry with odd N between the first-chunk and chunk sizes: every staging buffer must stay 8-byte aligned
    (per-problem bytes = 3228 at n=10, m=20)."""
    for N in (2001, 1025, 6143):
        b = generate_g1(N, 10, 20, 0, 8, seed=900 + N)
        r = run_gpu(engine, b)
        assert (r.exitflag == 1).all() and np.abs(r.x - b.xref).max() < 1e-8
    check_vs_oracle(engine, oracle, generate_g1(1031, 10, 20, 0, 8, seed=77), "odd N")


def test_c5_mixed_sizes_one_call(cuda_lib, oracle):
    """BASELINE.json config 5 size mix through daqp_quadprog_batch: n in {8,16,...,128}, m = 4 n, the number of active
    constraints at the optimum drawn from U{0..n} (divergent iteration counts), one call for all shapes -- the shape
    groups run side by side on the library's lanes, largest estimated cost first. This test is the fp64 pass (exact path
    parity with the oracle); test_c5_mixed_sizes_fp32_one_call is the fp32 pass BASELINE.json names."""
    import daqp_b200
    rng = np.random.default_rng(55)
    probs, refs = [], []
    for k, n in enumerate([8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128] * 2):
        na = int(rng.integers(0, n + 1))
        b = generate_g1(1, n, 4 * n, 0, na, seed=5500 + k)
        refs.append(oracle.solve(b))
        probs.append(dict(H=b.H[0], f=b.f[0], A=b.A[0], bupper=b.bupper[0], blower=b.blower[0]))
    out = daqp_b200.quadprog_batch(probs)
    its = []
    for (x, fval, flag, info), o in zip(out, refs):
        assert flag == o.exitflag[0] == 1 and info["iterations"] == o.iter[0]
        np.testing.assert_allclose(x, o.x[0], rtol=0, atol=1e-9 * (1 + np.abs(o.x[0]).max()))
        np.testing.assert_allclose(info["lam"], o.lam[0], rtol=0, atol=1e-7 * (1 + np.abs(o.lam[0]).max()))
        its.append(info["iterations"])
    assert max(its) > 10 * max(1, min(its))  # the iteration counts really diverge


def test_c5_mixed_sizes_fp32_one_call(cuda_lib, oracle_libs, capsys):
    """BASELINE.json config 5 as stated: fp32, mixed sizes n in {8..128}, m = 4 n, random active-set sizes, ONE call
    (daqp_quadprog_batch_f32, the single-precision ABI of the reference). Against the oracle compiled with c_float = float
    (pinned on the reference's -DDAQP_SINGLE_PRECISION build): every problem solved to the constructed optimum within the
    reference's own gate (1e-4, core_tests.jl:26-30, relative to |x|), exit flags equal, and the path-mismatch RATE
    (iteration counts that differ from the fp32 oracle) reported and bounded. The reference's tolerances sit far below
    fp32 epsilon (dual_tol 1e-12, sing_tol 3.7e-11), so two correct fp32 implementations part on near-ties."""
    import daqp_b200
    orc = oracle_libs.OracleLib(single=True)
    rng = np.random.default_rng(56)
    probs, batches = [], []
    for k, n in enumerate([8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128] * 3):
        na = int(rng.integers(0, n + 1))
        b = generate_g1(1, n, 4 * n, 0, na, seed=5600 + k)
        batches.append(b)
        probs.append(dict(H=b.H[0], f=b.f[0], A=b.A[0], bupper=b.bupper[0], blower=b.blower[0]))
    out = daqp_b200.quadprog_batch_f32(probs)
    mism, worst, by_n, solved, flag_mism = 0, 0.0, {}, 0, 0
    for (x, fval, flag, info), b in zip(out, batches):
        o = orc.solve(b)
        assert x.dtype == np.float32
        flag_mism += int(flag != o.exitflag[0])
        if flag != 1 or o.exitflag[0] != 1:
            continue
        solved += 1
        err = np.abs(x - b.xref[0]).max() / (1 + np.abs(b.xref[0]).max())
        worst = max(worst, err)
        assert err < 1e-3, f"n={b.n}: fp32 solution {err:.2e} away from the constructed optimum"
        d = int(info["iterations"] != o.iter[0])
        mism += d
        by_n[b.n] = by_n.get(b.n, 0) + d
    rate = mism / max(1, solved)
    with capsys.disabled():
        print(f"\nC5 fp32, {len(out)} problems, n = 8..128: {solved} optimal in both, exit flags differ on {flag_mism}; "
              f"path-mismatch rate vs the fp32 oracle {rate:.3f} (by n: {by_n}), worst relative |x - xref| {worst:.2e}")
    assert solved >= 0.95 * len(out) and flag_mism <= 0.05 * len(out) and rate <= 0.25


def test_fp32_path_vs_fp32_oracle(engine, oracle_libs):
    """fp32 arithmetic end to end (BASELINE.json config 5 precision) against the oracle compiled with c_float = float,
    which is pinned against the reference's -DDAQP_SINGLE_PRECISION build. The reference's tolerances sit far below
    fp32 epsilon (dual_tol 1e-12, sing_tol 3.7e-11), so two correct fp32 implementations may take different paths on a
    near-tie: the bar (SURVEY.md §7 "fp32") is every problem solved to the constructed optimum and a reported, bounded
    path-mismatch RATE, not bit parity."""
    orc = oracle_libs.OracleLib(single=True)
    total = mism = 0
    for cfg in [(300, 10, 20, 0, 8), (200, 20, 60, 4, 16), (60, 50, 150, 0, 40), (40, 32, 128, 0, 20)]:
        b = generate_g1(*cfg, seed=3200 + cfg[1])
        o = orc.solve(b)
        r = engine.solve_batch_f32(b.H, b.f, b.A, b.bupper, b.blower, None, ms=b.ms)
        assert r.x.dtype == np.float32
        ok = (o.exitflag == 1)
        assert ok.mean() > 0.95 and (r.exitflag[ok] == 1).mean() > 0.98
        both = ok & (r.exitflag == 1)
        scale = 1 + np.abs(b.xref[both]).max(axis=1)
        assert (np.abs(r.x[both] - b.xref[both]).max(axis=1) <= 1e-4 * scale).all()  # the reference's own test gate
        assert np.median(np.abs(r.x[both] - o.x[both]).max(axis=1)) <= 1e-5
        total += int(both.sum()); mism += int((r.iter[both] != o.iter[both]).sum())
    print(f"fp32 path mismatch rate vs the fp32 oracle: {mism}/{total}")
    assert mism <= 0.05 * total


WS_GOLDEN = ["wsseq_n10_m30_ms3", "wsseq_n20_m60_ms5", "wsseq_n50_m150", "wsseq_soft_n12_m40", "wsseq_equalities_n16_m48"]


@pytest.mark.parametrize("name", WS_GOLDEN)
def test_workspace_sequence_matches_reference(cuda_lib, name):
    """Persistent workspace: setup once, then update(f, b) + warm solve, against the UNMODIFIED reference driven through
    setup_daqp / daqp_update_ldp(UPDATE_v + UPDATE_d) / daqp_solve on a kept workspace (tests/golden/
    make_golden_workspace.py). Exit flags, iteration counts and working sets (factor order) must be equal for every
    solve of a problem as long as its previous solves ended (soft-)optimal: the state a reference workspace is left in
    after an INFEASIBLE exit (a zero pivot in D) is not something to reproduce."""
    import os
    import daqp_b200
    from common import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    use_sense = bool(d["use_sense"])
    mdl = daqp_b200.BatchModel().setup(d["H"], d["f"], d["A"], d["bupper"], d["blower"],
                                       d["sense"].astype(np.int32) if use_sense else None, ms=int(d["ms"]))
    N = d["H"].shape[0]
    live = np.ones(N, bool)
    checked = 0
    for k in range(int(d["K"]) + 1):
        if k > 0:
            mdl.update(f=d[f"f{k-1}"], bupper=d[f"bu{k-1}"], blower=d[f"bl{k-1}"])
        r = mdl.solve(warm=True, diag=True)
        flag, it = d[f"flag_{k}"], d[f"iter_{k}"]
        np.testing.assert_array_equal(r.exitflag[live], flag[live], err_msg=f"{name} solve {k}: exit flags")
        np.testing.assert_array_equal(r.iter[live], it[live], err_msg=f"{name} solve {k}: iteration counts")
        ok = live & (flag > 0)
        want_ws = d[f"ws_{k}"]
        got = r.working_sets()
        for p in np.nonzero(ok)[0]:
            assert got[p] == want_ws[p, : d[f"nact_{k}"][p]].tolist(), f"{name} solve {k} problem {p}: working set"
        f_tol = 1e-6 if "soft" in name else None
        assert_parity(d[f"x_{k}"][ok], d[f"lam_{k}"][ok], d[f"fval_{k}"][ok], flag[ok], it[ok], r.x[ok], r.lam[ok],
                      r.fval[ok], r.exitflag[ok], r.iter[ok], f"{name} solve {k}", f_tol=f_tol)
        checked += int(ok.sum())
        live &= flag > 0
    assert checked >= N  # the sequences are not vacuous
    cold = mdl.solve(warm=False)  # same data, cold start: same optimum
    ok = live
    if ok.any():
        np.testing.assert_allclose(cold.x[ok], r.x[ok], rtol=0, atol=1e-8 * (1 + np.abs(r.x[ok]).max()))
    mdl.close()


WS_SHARED = ["wsshared_n10_m30", "wsshared_n20_m60_ms5", "wsshared_n50_m150"]


@pytest.mark.parametrize("name", WS_SHARED)
def test_shared_workspace_matches_reference(cuda_lib, name):
    """Shared workspace: G matrix sets, K problems per set that differ only in f and the bounds. Every problem must get
    what the UNMODIFIED reference gives it through its own workspace -- setup_daqp(H_g, f_p, A_g, b_p) + daqp_solve, then
    update + warm solve per step (tests/golden/make_golden_shared.py): exit flags, iteration counts and working sets equal
    for as long as the problem's previous solves ended optimal. The same sequence through a PLAIN workspace on the
    replicated matrices must agree (same kernels; only the matrix indexing and the place where d is summed differ)."""
    import os
    import daqp_b200
    from common import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    G, K, ms = int(d["G"]), int(d["Kp"]), int(d["ms"])
    N = G * K
    mdl = daqp_b200.BatchModel().setup_shared(d["H"], d["A"], K, ms=ms, m=int(d["m"]))
    with pytest.raises(RuntimeError, match="update first"):
        mdl.solve()
    plain = daqp_b200.BatchModel().setup(np.repeat(d["H"], K, axis=0), d["f"], np.repeat(d["A"], K, axis=0), d["bupper"],
                                         d["blower"], None, ms=ms)
    live = np.ones(N, bool)
    checked = 0
    for k in range(int(d["K"]) + 1):
        if k == 0:
            mdl.update(f=d["f"], bupper=d["bupper"], blower=d["blower"])
        else:
            mdl.update(f=d[f"f{k-1}"], bupper=d[f"bu{k-1}"], blower=d[f"bl{k-1}"])
            plain.update(f=d[f"f{k-1}"], bupper=d[f"bu{k-1}"], blower=d[f"bl{k-1}"])
        r = mdl.solve(warm=True, diag=True)
        q = plain.solve(warm=True, diag=True)
        np.testing.assert_array_equal(r.exitflag, q.exitflag)
        np.testing.assert_array_equal(r.iter, q.iter)
        okq = q.exitflag > 0
        # (d = b * scaling + M v is summed by the setup kernel in the plain workspace's first solve and by the update kernel
        # here: same value up to the summation order)
        np.testing.assert_allclose(r.x[okq], q.x[okq], rtol=0, atol=1e-11 * (1 + np.abs(q.x[okq]).max()))
        np.testing.assert_allclose(r.lam[okq], q.lam[okq], rtol=0, atol=1e-9 * (1 + np.abs(q.lam[okq]).max()))
        flag, it = d[f"flag_{k}"], d[f"iter_{k}"]
        np.testing.assert_array_equal(r.exitflag[live], flag[live], err_msg=f"{name} solve {k}: exit flags")
        np.testing.assert_array_equal(r.iter[live], it[live], err_msg=f"{name} solve {k}: iteration counts")
        ok = live & (flag > 0)
        got = r.working_sets()
        for p in np.nonzero(ok)[0]:
            assert got[p] == d[f"ws_{k}"][p, : d[f"nact_{k}"][p]].tolist(), f"{name} solve {k} problem {p}: working set"
        assert_parity(d[f"x_{k}"][ok], d[f"lam_{k}"][ok], d[f"fval_{k}"][ok], flag[ok], it[ok], r.x[ok], r.lam[ok],
                      r.fval[ok], r.exitflag[ok], r.iter[ok], f"{name} solve {k}")
        checked += int(ok.sum())
        live &= flag > 0
    assert checked >= N
    mdl.close(); plain.close()


def test_workspace_device_entry_points(cuda_lib):
    """update_device / solve_device (torch CUDA tensors, asynchronous) give the same sequence as the host entry points."""
    import os
    import torch
    import daqp_b200
    from common import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, "wsseq_n20_m60_ms5.npz"))
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    host = daqp_b200.BatchModel().setup(d["H"], d["f"], d["A"], d["bupper"], d["blower"], None, ms=int(d["ms"]))
    devm = daqp_b200.BatchModel().setup(d["H"], d["f"], d["A"], d["bupper"], d["blower"], None, ms=int(d["ms"]))
    for k in range(int(d["K"]) + 1):
        if k > 0:
            host.update(f=d[f"f{k-1}"], bupper=d[f"bu{k-1}"], blower=d[f"bl{k-1}"])
            devm.update_device(t(d[f"f{k-1}"]), t(d[f"bu{k-1}"]), t(d[f"bl{k-1}"]))
        rh = host.solve(warm=True)
        rd = devm.solve_device(warm=True)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(rd["exitflag"].cpu().numpy(), rh.exitflag)
        np.testing.assert_array_equal(rd["iter"].cpu().numpy(), rh.iter)
        ok = rh.exitflag > 0
        np.testing.assert_array_equal(rd["x"].cpu().numpy()[ok], rh.x[ok])
        np.testing.assert_array_equal(rd["lam"].cpu().numpy()[ok], rh.lam[ok])
    host.close(); devm.close()


def test_settings_and_limits(engine, oracle):
    b = generate_g1(100, 20, 60, 0, 16, seed=51)
    o, r = check_vs_oracle(engine, oracle, b, "iter_limit", settings={"iter_limit": 7})
    assert (r.exitflag == -4).all() and (r.iter == 7).all()  # core_tests.jl:32-35 semantics
    check_vs_oracle(engine, oracle, b, "loose primal_tol", settings={"primal_tol": 1e-3})
    check_vs_oracle(engine, oracle, generate_g1(50, 10, 30, 0, 0, seed=52), "unconstrained optimum")
    b = generate_g1(16, 8, 20, 0, 6, seed=53)
    b.sense[:, 0] = 16
    assert (run_gpu(engine, b, use_sense=True).exitflag == -8).all()  # binaries: out of scope, flagged
    b = generate_g1(16, 8, 20, 0, 6, seed=54)
    b.H[:, 0, :] = 0; b.H[:, :, 0] = 0
    assert (run_gpu(engine, b).exitflag == -8).all()  # singular H needs the proximal driver: flagged


@pytest.mark.parametrize("shape,mode", [((12, 36, 4, 9), None), ((70, 150, 0, 40), "4")])
def test_decision_trace_matches_oracle(engine, oracle, shape, mode, monkeypatch):
    """DAQPB200Diag.trace (extended / team instantiations): the kernel's log of working-set decisions -- add with side,
    remove, exit flag -- equals the oracle's log entry for entry (the objective words differ in the last bits)."""
    import torch
    import daqp_b200
    if mode is not None:
        monkeypatch.setenv("DAQP_B200_TEAM", mode)
    n, m, ms, na = shape
    b = generate_g1(24, n, m, ms, na, seed=9100 + n)
    cap = 1024
    o = oracle.solve(b, log_cap=cap)
    dev = torch.device("cuda:0")
    t = lambda a, ty=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=ty, device=dev).contiguous()
    diag = daqp_b200.Engine.alloc_diag(b.N, n, m, dev)
    diag["trace"] = torch.zeros((b.N, 1 + 2 * cap), dtype=torch.int32, device=dev)
    out = engine.solve_batch_device(t(b.H), t(b.f), t(b.A), t(b.bupper), t(b.blower), None, ms=ms, diag=diag)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["iter"].cpu().numpy(), o.iter)
    tr = diag["trace"].cpu().numpy()
    for p in range(b.N):
        g = tr[p, 1:1 + 2 * int(tr[p, 0])].reshape(-1, 2)
        keep = lambda log: [(int(c), int(v)) for c, v in log if c not in (8, 9)]
        assert keep(g) == keep(o.oplog[p]), f"problem {p}"
        assert (g[:, 0] == 8).sum() == (o.oplog[p][:, 0] == 8).sum()


def test_time_limit(engine, oracle):
    """settings.time_limit (daqp.c:95-103): the clock is read every 32nd iteration, so with a limit no solve can meet a
    problem either ends before its 32nd iteration, untouched, or leaves there with EXIT_TIMELIMIT (-7) and iter = 32; a
    generous limit changes nothing. Plain kernel (n = 50) and team mode (n = 70)."""
    for (n, m, na, seed) in ((50, 150, 40, 55), (70, 200, 50, 56)):
        b = generate_g1(200, n, m, 0, na, seed=seed)
        o = oracle.solve(b)
        r = run_gpu(engine, b, time_limit=1e-9)
        long = o.iter > 32
        assert long.any() and (~long).any() or long.all()
        assert (r.exitflag[long] == -7).all() and (r.iter[long] == 32).all()
        np.testing.assert_array_equal(r.exitflag[~long], o.exitflag[~long])
        np.testing.assert_array_equal(r.iter[~long], o.iter[~long])
        r = run_gpu(engine, b, time_limit=100.0)
        assert_parity(o.x, o.lam, o.fval, o.exitflag, o.iter, r.x, r.lam, r.fval, r.exitflag, r.iter, "generous time limit")


def test_no_linear_term_and_diagonal_hessian(engine, oracle):
    b = generate_g1(100, 12, 36, 4, 9, seed=61)
    b.H[:] = (np.eye(12) * np.linspace(1, 5, 12))[None]
    check_vs_oracle(engine, oracle, b, "diagonal H with simple bounds")


def test_drop_in_entry_points(cuda_lib, oracle):
    """daqp_quadprog (batch of one) and daqp_quadprog_batch (array of structs, mixed shapes)."""
    import daqp_b200
    for name, want in [("lit_model_qp", [-1, -1]), ("lit_model_qp_flipped", [1, 1]), ("lit_model_qp_half", [-0.5, -0.5]),
                       ("lit_eigen_basic", [-1, -1])]:
        b, d = load_golden(name)
        x, fval, flag, info = daqp_b200.solve(b.H[0], b.f[0], b.A[0], b.bupper[0], b.blower[0], b.sense[0])
        assert flag == 1 and info["iterations"] == int(d["iter"][0])
        np.testing.assert_allclose(x, want, atol=1e-9)
        np.testing.assert_allclose(fval, d["fval"][0], atol=1e-9)
    mixed, refs = [], []
    for seed, (n, m, ms, na) in enumerate([(10, 20, 0, 8), (20, 60, 10, 16), (10, 20, 0, 8), (7, 15, 3, 5)] * 3):
        b = generate_g1(1, n, m, ms, na, seed=700 + seed)
        o = oracle.solve(b)
        mixed.append(dict(H=b.H[0], f=b.f[0], A=b.A[0], bupper=b.bupper[0], blower=b.blower[0]))
        refs.append(o)
    out = daqp_b200.quadprog_batch(mixed)
    for (x, fval, flag, info), o in zip(out, refs):
        assert flag == o.exitflag[0] and info["iterations"] == o.iter[0]
        np.testing.assert_allclose(x, o.x[0], rtol=0, atol=1e-9 * (1 + np.abs(o.x[0]).max()))
        np.testing.assert_allclose(info["lam"], o.lam[0], rtol=0, atol=1e-7 * (1 + np.abs(o.lam[0]).max()))


def test_workspace_pieces_reset_deactivate_extract_and_ldp(cuda_lib, oracle):
    """The pieces of daqp_solve the interfaces call one by one on a workspace from setup_daqp (Julia: reset(model) =
    daqp_deactivate_constraints + reset_daqp_workspace, api.jl:384-387; Simulink: daqp_ldp + daqp_extract_result,
    DAQP_sfunc.c:567-571): a second daqp_solve continues warm (one iteration), after the reset the solve is the cold one
    again (the reference's iteration count), daqp_ldp + daqp_extract_result give what daqp_solve gives."""
    import ctypes as C
    import daqp_b200 as dq
    from oracle import harness
    Problem, Settings, Result, Workspace, _ = harness._F64
    L = dq.lib()
    for f in ("daqp_solve", "reset_daqp_workspace", "daqp_deactivate_constraints", "daqp_extract_result", "free_daqp_workspace",
              "free_daqp_ldp"):
        getattr(L, f).restype = None
    L.setup_daqp.restype = C.c_int; L.daqp_ldp.restype = C.c_int
    b = generate_g1(1, 12, 36, 4, 9, seed=88)
    o = oracle.solve(b)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    se = np.zeros(b.m, np.intc)
    qp = Problem(b.n, b.m, b.ms, dp(b.H[0]), dp(b.f[0]), dp(b.A[0]), dp(b.bupper[0]), dp(b.blower[0]),
                 se.ctypes.data_as(C.POINTER(C.c_int)), None, 0, 0)
    work = Workspace()
    t = C.c_double(0)
    assert L.setup_daqp(C.byref(qp), C.byref(work), C.byref(t)) == 1
    def solve():
        x = np.zeros(b.n); lam = np.zeros(b.m)
        res = Result(dp(x), dp(lam), 0, 0, 0, 0, 0, 0, 0)
        L.daqp_solve(C.byref(res), C.byref(work))
        return x, lam, res
    x1, lam1, r1 = solve()
    assert r1.exitflag == 1 and r1.iter == o.iter[0]
    np.testing.assert_allclose(x1, o.x[0], atol=1e-9)
    x2, _, r2 = solve()                      # kept factor and working set: nothing to do
    assert r2.exitflag == 1 and r2.iter == 1
    L.daqp_deactivate_constraints(C.byref(work)); L.reset_daqp_workspace(C.byref(work))
    assert work.n_active == 0 and all((work.sense[i] & 1) == 0 for i in range(b.m))
    x3, lam3, r3 = solve()                   # cold again
    assert r3.exitflag == 1 and r3.iter == o.iter[0]
    np.testing.assert_array_equal(x3, x1)
    L.daqp_deactivate_constraints(C.byref(work)); L.reset_daqp_workspace(C.byref(work))
    assert L.daqp_ldp(C.byref(work)) == 1 and work.iterations == o.iter[0]
    x4 = np.zeros(b.n); lam4 = np.zeros(b.m)
    res4 = Result(dp(x4), dp(lam4), 0, 0, 0, 0, 0, 0, 0)
    L.daqp_extract_result(C.byref(res4), C.byref(work))
    np.testing.assert_array_equal(x4, x1); np.testing.assert_array_equal(lam4, lam1)
    assert res4.iter == o.iter[0] and abs(res4.fval - o.fval[0]) <= 1e-9 * (1 + abs(o.fval[0]))
    L.free_daqp_workspace(C.byref(work)); L.free_daqp_ldp(C.byref(work))


def test_device_entry_point_and_chunking(engine, oracle):
    """Device-pointer entry on the torch stream; a tiny scratch limit forces the multi-chunk path."""
    import torch
    import daqp_b200
    b = generate_g1(700, 20, 60, 4, 16, seed=81)
    o = oracle.solve_packed(b)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    eng = daqp_b200.Engine()
    eng.set_scratch_limit(12 << 20)  # ~ 200 problems per chunk
    out = eng.solve_batch_device(t(b.H), t(b.f), t(b.A), t(b.bupper), t(b.blower), None, ms=b.ms)
    torch.cuda.synchronize()
    st = eng.stats()
    assert st["solve_launches"] >= 3
    assert_parity(o.x, o.lam, o.fval, o.exitflag, o.iter, out["x"].cpu().numpy(), out["lam"].cpu().numpy(),
                  out["fval"].cpu().numpy(), out["exitflag"].cpu().numpy(), out["iter"].cpu().numpy(), "device entry")
    eng.close()


def test_full_size_properties(engine):
    """BASELINE.json C3 at a size the oracle cannot cover: size-independent properties only.
    (1) every problem reports OPTIMAL, (2) x equals the construction-known optimum, (3) the active set equals the
    constructed one, (4) KKT residuals vanish, (5) solving the same batch twice is bit-identical (idempotence /
    no dependence on warp scheduling)."""
    b = generate_config("C3", N=20000)
    r = run_gpu(engine, b)
    assert (r.exitflag == 1).all()
    assert np.abs(r.x - b.xref).max() < 1e-6  # reference gate: 1e-4 (core_tests.jl:26-30)
    assert (np.sign(r.lam) == b.active_ref).all()
    stat, pf, comp = kkt_residuals(b, r.x, r.lam)
    assert stat.max() < 1e-7 and pf.max() < 1e-7 and comp.max() < 1e-7
    r2 = run_gpu(engine, b)
    np.testing.assert_array_equal(r.x, r2.x)
    np.testing.assert_array_equal(r.iter, r2.iter)
    # a batch is a set of independent problems: permuting it permutes the answers
    perm = np.random.default_rng(0).permutation(b.N)[:4000]
    sub = engine.solve_batch(b.H[perm], b.f[perm], b.A[perm], b.bupper[perm], b.blower[perm], None, ms=b.ms)
    np.testing.assert_array_equal(sub.x, r.x[perm])


@pytest.mark.parametrize("name", ldp_golden_names())
def test_pure_ldp_inputs_match_reference(cuda_lib, name):
    """H == NULL, f == NULL through daqp_quadprog_batch / daqp_quadprog (the LDP min |x|^2 over the constraints, reference
    utils.c:103-110; the form Julia's polyhedral tools use): exit flags (infeasible polyhedra included) and iteration
    counts equal to the reference's, x / lam within the fp64 bar, fval untouched (the reference sets none without f)."""
    import daqp_b200
    b, d = load_ldp_golden(name)
    use_sense = bool(d["use_sense"])
    probs = [{"A": b.A[p], "bupper": b.bupper[p], "blower": b.blower[p], **({"sense": b.sense[p]} if use_sense else {})}
             for p in range(b.N)]
    out = daqp_b200.quadprog_batch(probs)
    flag = np.array([o[2] for o in out]); it = np.array([o[3]["iterations"] for o in out])
    np.testing.assert_array_equal(flag, d["exitflag"])
    np.testing.assert_array_equal(it, d["iter"])
    for p in np.nonzero(d["exitflag"] > 0)[0]:
        np.testing.assert_allclose(out[p][0], d["x"][p], atol=1e-9 * (1 + np.abs(d["x"][p]).max()), err_msg=f"{name}[{p}] x")
        np.testing.assert_allclose(out[p][3]["lam"], d["lam"][p], atol=1e-7 * (1 + np.abs(d["lam"][p]).max()), err_msg=f"{name}[{p}]")
        assert out[p][1] == 0.0  # fval as handed in
    x, fval, fl, info = daqp_b200.solve(None, None, b.A[0], b.bupper[0], b.blower[0], b.sense[0] if use_sense else None)
    assert fl == d["exitflag"][0] and info["iterations"] == d["iter"][0]
    # a linear term without a Hessian is an LP: the reference's proximal driver, flagged out of scope here
    assert daqp_b200.quadprog_batch([{"f": np.ones(b.n), "A": b.A[0], "bupper": b.bupper[0], "blower": b.blower[0]}])[0][2] == -8


@pytest.mark.parametrize("name", rawldp_golden_names())
def test_raw_ldp_batch_and_hand_filled_workspace(cuda_lib, name):
    """Raw LDPs (M = A as given, no normalisation, Rinv == NULL): (1) daqp_b200_ldp_batch, all polyhedra of the fixture in one
    call, (2) the reference's own flow on a hand-filled DAQPWorkspace -- allocate_daqp_workspace, allocate_daqp_settings,
    fields stored by the caller, daqp_ldp, free_daqp_workspace (interfaces/daqp-julia/src/api.jl:428-459) -- one polyhedron
    at a time. Exit flags (the max-radius bound included), iteration counts and working sets in factor order equal to the
    reference's; u within 1e-9, fval = |u|^2."""
    import ctypes as C
    import os
    import daqp_b200 as dq
    from oracle import harness
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, m, ms, P = int(d["n"]), int(d["m"]), int(d["ms"]), d["exitflag"].shape[0]
    fvb = None if d["fval_bound"] < 0 else float(d["fval_bound"])
    L = dq.lib()
    L.daqp_b200_ldp_batch.restype = C.c_int
    A = np.ascontiguousarray(d["A"]); bu = np.ascontiguousarray(d["bupper"]); bl = np.ascontiguousarray(d["blower"])
    u = np.zeros((P, n)); lam = np.zeros((P, m)); fv = np.zeros(P); flag = np.zeros(P, np.intc); it = np.zeros(P, np.intc)
    nact = np.zeros(P, np.intc); ws = np.zeros((P, n + 1), np.intc)
    dg = dq.DAQPB200Diag(nact.ctypes.data_as(C.POINTER(C.c_int)), ws.ctypes.data_as(C.POINTER(C.c_int)), None, None, None)
    st = dq.default_settings(**({"fval_bound": fvb} if fvb is not None else {}))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    sense = np.ascontiguousarray(d["sense"], np.intc) if d["sense"].any() else None
    assert L.daqp_b200_ldp_batch(None, P, n, m, ms, dp(A), dp(bu), dp(bl), ip(sense) if sense is not None else None, C.byref(st), dp(u), dp(lam), dp(fv), ip(flag),
                                 ip(it), C.byref(dg)) == 0, L.daqp_b200_last_error().decode()
    np.testing.assert_array_equal(flag, d["exitflag"])
    np.testing.assert_array_equal(it, d["iter"])
    for p in range(P):
        assert ws[p, :nact[p]].tolist() == d["ws"][p, :d["n_active"][p]].tolist(), f"{name}[{p}]: working set"
    ok = d["exitflag"] > 0
    np.testing.assert_allclose(u[ok], d["u"][ok], atol=1e-9 * (1 + np.abs(d["u"][ok]).max()))
    np.testing.assert_allclose(2 * fv[ok], d["fval"][ok], rtol=1e-9)
    rb = dq.ldp_batch(A, bu, bl, sense, **({"fval_bound": fvb} if fvb is not None else {}))  # the Python wrapper of the same call
    np.testing.assert_array_equal(rb.exitflag, flag); np.testing.assert_array_equal(rb.iter, it)
    np.testing.assert_array_equal(rb.x, u)
    for p in range(min(P, 6)):
        r = harness.raw_ldp(L, A[p], bu[p], bl[p], sense[p] if sense is not None else None, ms, fvb)
        assert r["exitflag"] == d["exitflag"][p] and r["iter"] == d["iter"][p], f"{name}[{p}] via daqp_ldp"
        assert r["ws"] == d["ws"][p, :d["n_active"][p]].tolist()
        if r["exitflag"] > 0:
            np.testing.assert_allclose(r["u"], d["u"][p], atol=1e-9 * (1 + np.abs(d["u"][p]).max()))
            assert abs(r["fval"] - d["fval"][p]) <= 1e-9 * (1 + d["fval"][p])


def test_raw_ldp_device_entry(cuda_lib):
    """daqp_b200_ldp_device (device arrays, asynchronous on the caller's stream) gives what the host entry gives."""
    import ctypes as C
    import torch
    import daqp_b200 as dq
    d = np.load(os.path.join(GOLDEN_DIR, "rawldp_n12_m36_ms4.npz"))
    n, m, ms, P = int(d["n"]), int(d["m"]), int(d["ms"]), d["exitflag"].shape[0]
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev).contiguous()
    A, bu, bl = t(d["A"]), t(d["bupper"]), t(d["blower"])
    u = torch.zeros((P, n), dtype=torch.float64, device=dev); lam = torch.zeros((P, m), dtype=torch.float64, device=dev)
    fv = torch.zeros(P, dtype=torch.float64, device=dev)
    flag = torch.zeros(P, dtype=torch.int32, device=dev); it = torch.zeros(P, dtype=torch.int32, device=dev)
    L = dq.lib()
    L.daqp_b200_ldp_device.restype = C.c_int
    p = lambda x: C.c_void_p(x.data_ptr())
    st = dq.default_settings()
    stream = torch.cuda.current_stream(dev).cuda_stream or 1
    assert L.daqp_b200_ldp_device(None, P, n, m, ms, p(A), p(bu), p(bl), None, C.byref(st), p(u), p(lam), p(fv), p(flag), p(it),
                                  C.c_void_p(stream)) == 0, L.daqp_b200_last_error().decode()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(flag.cpu().numpy(), d["exitflag"])
    np.testing.assert_array_equal(it.cpu().numpy(), d["iter"])
    ok = d["exitflag"] > 0
    np.testing.assert_allclose(u.cpu().numpy()[ok], d["u"][ok], atol=1e-9 * (1 + np.abs(d["u"][ok]).max()))
    np.testing.assert_allclose(2 * fv.cpu().numpy()[ok], d["fval"][ok], rtol=1e-9)


@pytest.mark.parametrize("name", bnb_golden_names())
def test_bnb_batched_node_relaxations_match_reference(cuda_lib, name):
    """Binary constraints (sense & 16): branch and bound with the node relaxations solved as batches on the GPU
    (daqp_b200_bnb behind daqp_quadprog, reference src/bnb.c:23-128) against the reference's own output on the port of its
    MIQP generator and on the literals of its tests: same exit flag, same optimum (x to 1e-6, fval to 1e-7 relative),
    every binary on one of its bounds. The walk differs from the reference's depth-first one, so iteration and node
    counts are not compared -- except where the reference's tests pin them (zero-dual endpoints: one node)."""
    import daqp_b200
    b, d = load_golden(name)
    for p in range(b.N):
        x, fval, flag, info = daqp_b200.solve(b.H[p], b.f[p], b.A[p] if b.m > b.ms else np.zeros((0, b.n)), b.bupper[p],
                                              b.blower[p], b.sense[p])
        assert flag == d["exitflag"][p] == 1, f"{name}[{p}]: exit flag {flag}"
        np.testing.assert_allclose(x, d["x"][p], atol=1e-6 * (1 + np.abs(d["x"][p]).max()), err_msg=f"{name}[{p}]")
        assert abs(fval - d["fval"][p]) <= 1e-7 * (1 + abs(d["fval"][p])), f"{name}[{p}]: fval {fval} vs {d['fval'][p]}"
        for i in np.nonzero(b.sense[p] & 16)[0]:
            val = x[i] if i < b.ms else b.A[p, i - b.ms] @ x
            assert min(abs(val - b.bupper[p, i]), abs(val - b.blower[p, i])) < 1e-5
        assert info["nodes"] >= 1 and info["iterations"] >= 1
        if name.startswith("bnb_lit_zero_dual"):
            assert info["nodes"] == 1  # core_tests.jl:160-178


def test_bnb_narrow_waves_and_infeasible(cuda_lib):
    """The same optimum whatever the wave width (1 = one relaxation per launch, the reference's own granularity), and an
    MIQP whose binaries cannot satisfy an equality ends INFEASIBLE (-1)."""
    import ctypes as C
    import daqp_b200 as d
    from daqp_b200.problems import generate_miqp
    L = d.lib()
    L.daqp_b200_bnb.restype = C.c_int
    b = generate_miqp(3, 10, 20, 8, 6, seed=711)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for p in range(b.N):
        xs = []
        for wave in (1, 4, 256):
            x = np.empty(b.n); lam = np.empty(b.m)
            se = np.ascontiguousarray(b.sense[p], np.intc)
            qp = d.DAQPProblem(b.n, b.m, b.ms, dp(b.H[p]), dp(b.f[p]), dp(b.A[p]), dp(b.bupper[p]), dp(b.blower[p]),
                               se.ctypes.data_as(C.POINTER(C.c_int)), None, 0, 0)
            res = d.DAQPResult(dp(x), dp(lam), 0, 0, 0, 0, 0, 0, 0)
            assert L.daqp_b200_bnb(None, C.byref(qp), None, C.byref(res), wave) == 0 and res.exitflag == 1
            xs.append(x.copy())
        np.testing.assert_allclose(xs[0], xs[1], atol=1e-9); np.testing.assert_allclose(xs[0], xs[2], atol=1e-9)
    # x1, x2 binary in {0, 1} and x1 + x2 = 0.5 exactly: no integer point
    H = np.eye(2); f = np.zeros(2); A = np.ones((1, 2))
    x, fval, flag, info = d.solve(H, f, A, np.array([1.0, 1.0, 0.5]), np.array([0.0, 0.0, 0.5]), np.array([16, 16, 5], np.intc))
    assert flag == -1


def test_solve_packed_multi_matches_single(engine, oracle):
    """daqp_b200_solve_packed_multi: the batch cut into per-device blocks (one host thread + engine each). With one GPU
    both blocks land on device 0; the answers must equal the single-engine call bit for bit either way."""
    import torch
    import daqp_b200
    b = generate_g1(4001, 20, 60, 4, 16, seed=93)
    b.sense[::3, 7] = 1  # some warm-start bits: sense must travel with its block
    one = engine.solve_batch(b.H, b.f, b.A, b.bupper, b.blower, b.sense, ms=b.ms)
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    r, secs = daqp_b200.solve_batch_multi(b.H, b.f, b.A, b.bupper, b.blower, b.sense, ms=b.ms, devices=devs)
    for k in ("x", "lam", "fval", "exitflag", "iter"):
        np.testing.assert_array_equal(getattr(r, k), getattr(one, k))
    assert (secs > 0).all()
    o = oracle.solve(b.slice(0, 300), use_sense=True)
    np.testing.assert_array_equal(r.iter[:300], o.iter)


def _nccl_worker(rank, world, port, tmp):
    import os, sys
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import daqp_b200
    from daqp_b200.problems import generate_g1
    from daqp_b200.sharding import scatter_solve_gather
    n, m, ms = 20, 60, 4
    dev = torch.device(f"cuda:{rank}")
    arrays, b = None, None
    if rank == 0:
        b = generate_g1(1001, n, m, ms, 16, seed=91)
        arrays = {k: torch.from_numpy(getattr(b, k)).to(dev) for k in ("H", "f", "A", "bupper", "blower")}
    eng = daqp_b200.Engine(rank)

    def solve_local(loc):
        return eng.solve_batch_device(loc["H"], loc["f"], loc["A"], loc["bupper"], loc["blower"], None, ms=ms)

    out = scatter_solve_gather(arrays, n, m, ms, solve_local, src=0, device=dev)
    ok = True
    if rank == 0:
        whole = eng.solve_batch_device(*(arrays[k] for k in ("H", "f", "A", "bupper", "blower")), None, ms=ms)
        torch.cuda.synchronize()
        ok = all(torch.equal(out[k], whole[k]) for k in ("x", "lam", "iter", "exitflag"))
        ok = ok and float((out["x"].cpu() - torch.from_numpy(b.xref)).abs().max()) < 1e-8
        # second pass: warm-start bits (int32 sense) travel with their block -> one iteration per problem
        sense = torch.zeros((1001, m), dtype=torch.int32, device=dev)
        sense[whole["lam"] > 1e-12] = 1
        sense[whole["lam"] < -1e-12] = 3
        arrays = dict(arrays, sense=sense)

    def solve_local_warm(loc):
        return eng.solve_batch_device(loc["H"], loc["f"], loc["A"], loc["bupper"], loc["blower"], loc["sense"], ms=ms)

    out2 = scatter_solve_gather(arrays, n, m, ms, solve_local_warm, src=0, device=dev)
    if rank == 0:
        ok = ok and bool((out2["iter"] == 1).all()) and bool((out2["exitflag"] == 1).all())
        ok = ok and float((out2["x"] - whole["x"]).abs().max()) < 1e-9
        open(tmp, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_ranks_nccl(cuda_lib, tmp_path):
    """Scatter -> per-rank solve -> gather over NCCL equals the single-GPU solve bit for bit (needs 2 GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    marker = str(tmp_path / "nccl.txt")
    mp.spawn(_nccl_worker, args=(2, 29533, marker), nprocs=2, join=True)
    assert open(marker).read() == "ok"

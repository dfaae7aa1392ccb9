"""Minimal representation of polyhedra (reference daqp_minrep: include/api.h:54,
src/api.c:531-556, src/utils.c:808-835).

CPU part: the oracle restatement against golden vectors produced by the reference's own daqp_minrep, and against the
reference compiled here. GPU part (-m gpu): the batched CUDA path through the C ABI -- every LDP's exit flag and
iteration count EQUAL to the oracle's independent probes, is_redundant EQUAL to the reference's golden output, plus
size-independent properties at scale (idempotence, membership).
"""
import numpy as np
import pytest

from common import GOLDEN_DIR, load_minrep_special, minrep_golden_names
from daqp_b200.problems import generate_polyhedra

import os


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return d["A"], d["b"], int(d["n"]), int(d["m"]), int(d["ms"]), d["is_redundant"]


# ---- CPU: oracle pinned on the reference ---------------------------------------------------------------------------

@pytest.mark.parametrize("name", minrep_golden_names())
def test_oracle_minrep_matches_golden(oracle_libs, name):
    A, b, n, m, ms, want = load(name)
    o = oracle_libs.OracleLib()
    for q in range(A.shape[0]):
        np.testing.assert_array_equal(o.minrep(A[q], b[q]), want[q], err_msg=f"{name}[{q}] reference order")
        red, flag, it = o.minrep_independent(A[q], b[q])
        assert ((flag == -1) == (red == 1)).all() and (it >= 1).all()
        if red.all():  # empty polyhedron: every probe infeasible; the reference's answer follows from its probing order
            assert name.endswith("_empty") and want[q][0] == 1
        else:
            np.testing.assert_array_equal(red, want[q], err_msg=f"{name}[{q}] independent probes")


def test_oracle_minrep_special_cases(oracle_libs):
    """Duplicates, a scaled duplicate, an all-zero row (the reference's first LDL' append skips its singularity test,
    factorization.c:56, and the probe ends 'optimal'), empty polyhedra (one and two leading constraints dropped before
    the rest is non-empty), an unbounded one, a single constraint."""
    o = oracle_libs.OracleLib()
    sp = load_minrep_special()
    assert sp["box_with_cuts"][2].tolist() == [0, 0, 0, 0, 0, 1, 1]
    assert sp["unbounded"][2].tolist() == [0, 1, 0, 1]
    assert sp["empty"][2].tolist() == [1, 0, 0] and sp["empty_two_rounds"][2].tolist() == [1, 1, 0, 0]
    for k, (A, b, want) in sp.items():
        np.testing.assert_array_equal(o.minrep(A, b), want, err_msg=k)


def test_oracle_minrep_vs_live_reference(oracle_libs):
    if not oracle_libs.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    o = oracle_libs.OracleLib()
    for (P, n, m, ms, seed) in [(6, 4, 30, 0, 101), (6, 8, 50, 8, 102), (3, 30, 200, 4, 103), (6, 6, 40, 3, 104)]:
        A, b = generate_polyhedra(P, n, m, ms, seed)
        for q in range(P):
            for lib in ("libdaqp_ref.so", "libdaqp_ref_strict.so"):
                np.testing.assert_array_equal(o.minrep(A[q], b[q]), oracle_libs.ref_minrep(A[q], b[q], lib))


def test_minrep_kept_rows_describe_the_same_polyhedron():
    """Solver-independent check of the golden vectors themselves: points sampled in the polyhedron of the kept rows
    satisfy the dropped rows too."""
    A, b, n, m, ms, red = load("minrep_n5_m40_ms2")
    rng = np.random.default_rng(5)
    for q in range(A.shape[0]):
        full = np.vstack([np.eye(n)[:ms], A[q]])
        keep = red[q] == 0
        x = rng.standard_normal((20000, n)) * 3
        inside = (x @ full[keep].T <= b[q][keep]).all(axis=1)
        assert inside.sum() > 0
        assert (x[inside] @ full[~keep].T <= b[q][~keep] + 1e-9).all()


# ---- GPU: the batched CUDA path ---------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def engine(cuda_lib):
    import daqp_b200
    e = daqp_b200.Engine()
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", minrep_golden_names())
def test_cuda_minrep_matches_golden_and_oracle(engine, oracle_libs, name):
    A, b, n, m, ms, want = load(name)
    o = oracle_libs.OracleLib()
    engine.stats(reset=True)
    red, flag, it = engine.minrep_batch(A, b, ms=ms, info=True)
    st = engine.stats()
    assert st["solve_launches"] >= 1, "ldp_solve_kernel did not launch"
    np.testing.assert_array_equal(red, want, err_msg=f"{name}: is_redundant differs from the reference's daqp_minrep")
    for q in range(A.shape[0]):
        ored, oflag, oit = o.minrep_independent(A[q], b[q])
        if ored.all():  # empty polyhedron: re-run with leading constraints dropped (flags are those of the last round)
            continue
        np.testing.assert_array_equal(flag[q], oflag, err_msg=f"{name}[{q}]: LDP exit flags")
        np.testing.assert_array_equal(it[q], oit, err_msg=f"{name}[{q}]: LDP iteration counts")


@pytest.mark.gpu
def test_cuda_minrep_special_cases_and_drop_in(cuda_lib, engine):
    import daqp_b200
    for k, (A, b, want) in load_minrep_special().items():
        got = daqp_b200.minrep(A, b)  # the drop-in daqp_minrep symbol, one polyhedron
        np.testing.assert_array_equal(got, want, err_msg=k)
    # simple bounds: b longer than A's row count (daqp.pyx:641-645)
    A, b, n, m, ms, want = load("minrep_n10_m60_ms10")
    np.testing.assert_array_equal(daqp_b200.minrep(A[0], b[0]), want[0])


@pytest.mark.gpu
def test_cuda_minrep_device_entry_and_chunking(cuda_lib, oracle_libs):
    import torch
    import daqp_b200
    A, b = generate_polyhedra(300, 6, 36, 2, seed=77)
    o = oracle_libs.OracleLib()
    want = np.stack([o.minrep(A[q], b[q]) for q in range(A.shape[0])])
    eng = daqp_b200.Engine()
    dev = torch.device("cuda:0")
    out = eng.minrep_batch_device(torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev), ms=2)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["is_redundant"].cpu().numpy(), want)
    # the asynchronous device entry runs ONE round: an empty polyhedron comes back all ones
    Ae = torch.tensor([[[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0]]], dtype=torch.float64, device=dev)
    be = torch.tensor([[-1.0, -1.0, 1.0]], dtype=torch.float64, device=dev)
    oe = eng.minrep_batch_device(Ae, be)
    torch.cuda.synchronize()
    assert oe["is_redundant"].cpu().numpy().tolist() == [[1, 1, 1]]
    eng.set_scratch_limit(9 << 20)  # a few dozen polyhedra per chunk
    eng.stats(reset=True)
    red = eng.minrep_batch(A, b, ms=2)
    assert eng.stats()["solve_launches"] >= 3
    np.testing.assert_array_equal(red, want)
    eng.close()


@pytest.mark.gpu
def test_cuda_minrep_properties_at_scale(engine):
    """4096 polyhedra x 64 constraints = 262144 concurrent LDPs. Size-independent properties: (1) idempotence -- the
    kept rows of a polyhedron are all non-redundant when probed again on their own; (2) membership -- points inside the
    kept rows satisfy the dropped rows; (3) every LDP ended in a decided state (optimal or infeasible)."""
    P, n, m = 4096, 8, 64
    A, b = generate_polyhedra(P, n, m, 0, seed=2024)
    red, flag, it = engine.minrep_batch(A, b, info=True)
    assert np.isin(flag, (1, -1)).all()
    assert 0.1 < red.mean() < 0.9
    rng = np.random.default_rng(1)
    for q in rng.choice(P, 24, replace=False):
        keep = red[q] == 0
        x = rng.standard_normal((4000, n)) * 2 + np.linalg.lstsq(A[q], b[q] - 1.0, rcond=None)[0] * 0
        inside = (x @ A[q][keep].T <= b[q][keep]).all(axis=1)
        if inside.any():
            assert (x[inside] @ A[q][~keep].T <= b[q][~keep] + 1e-9).all()
    # idempotence on the polyhedra that kept the most common number of rows (one homogeneous batch)
    kept = (red == 0).sum(axis=1)
    mk = np.bincount(kept).argmax()
    sel = np.nonzero(kept == mk)[0]
    A2 = np.stack([A[q][red[q] == 0] for q in sel]); b2 = np.stack([b[q][red[q] == 0] for q in sel])
    red2 = engine.minrep_batch(A2, b2)
    assert red2.sum() == 0


@pytest.mark.gpu
def test_cuda_minrep_throughput_report(cuda_lib, oracle_libs):
    """Informational: device time of the batched path (CUDA events, device arrays) next to the reference's own
    daqp_minrep on one host thread (oracle/_ref when it travelled with the snapshot, else the oracle restatement) for a
    sample of the same polyhedra. Written to gpurun_out/minrep.json (copied to profiles/minrep_r01.json by hand); the
    only assertion is that both agree on the sample."""
    import json
    import time
    import torch
    import daqp_b200
    eng = daqp_b200.Engine()
    dev = torch.device("cuda:0")
    orc = oracle_libs.OracleLib()
    rows = []
    for (P, n, m, ms) in [(4096, 8, 64, 0), (2048, 10, 100, 0), (512, 20, 150, 0), (128, 50, 300, 0)]:
        A, b = generate_polyhedra(P, n, m, ms, seed=31 + n)
        dA, db = torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev)
        out = eng.minrep_batch_device(dA, db, ms=ms)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            out = eng.minrep_batch_device(dA, db, ms=ms, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms_dev = e0.elapsed_time(e1) / reps
        red = out["is_redundant"].cpu().numpy()
        sample = min(P, 16)
        use_ref = oracle_libs.have_ref()
        t0 = time.perf_counter()
        ref = np.stack([oracle_libs.ref_minrep(A[q], b[q]) if use_ref else orc.minrep(A[q], b[q]) for q in range(sample)])
        cpu_s = time.perf_counter() - t0
        np.testing.assert_array_equal(red[:sample], ref)
        rows.append({"P": P, "n": n, "m": m, "ms": ms, "ldps": P * m, "device_ms": ms_dev,
                     "polyhedra_per_s_device": P / ms_dev * 1e3, "ldps_per_s_device": P * m / ms_dev * 1e3,
                     "mean_iterations": float(out["iter"].float().mean()), "redundant_fraction": float(red.mean()),
                     "cpu_polyhedra_per_s_1thread": sample / cpu_s, "cpu_kind": "reference" if use_ref else "port"})
    eng.close()
    os.makedirs(os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out", "minrep.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print(json.dumps(rows))

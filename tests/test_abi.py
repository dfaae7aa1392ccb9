"""CPU suite: the C-ABI shared library builds for sm_100a, loads, exports every symbol include/daqp_b200.h declares,
keeps the reference's struct layouts, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "daqp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:daqp_|setup_daqp|allocate_daqp|free_daqp)[a-z0-9_]*)\s*\(", src)))


def test_header_symbols_exported(cuda_lib):
    syms = declared_symbols()
    assert {"daqp_quadprog", "daqp_default_settings", "daqp_quadprog_batch", "daqp_b200_solve_packed",
            "daqp_b200_solve_device"} <= set(syms)
    for s in syms:
        assert hasattr(cuda_lib, s), f"{s} declared in include/daqp_b200.h but not exported"


def test_struct_layout_matches_reference_abi():
    """Field offsets of the reference structs on LP64 (include/types.h:32-74, include/api.h:15-27)."""
    import daqp_b200 as d
    P, S, R = d.DAQPProblem, d.DAQPSettings, d.DAQPResult
    assert (C.sizeof(P), P.H.offset, P.sense.offset, P.nh.offset, P.problem_type.offset) == (80, 16, 56, 72, 76)
    assert (C.sizeof(S), S.cycle_tol.offset, S.fval_bound.offset, S.time_limit.offset) == (120, 40, 48, 112)
    assert (C.sizeof(R), R.exitflag.offset, R.solve_time.offset) == (64, 32, 48)


def test_layout_against_live_reference(oracle_libs):
    """When the reference is compiled here, drive IT with this package's ctypes structs (same bytes, same answer)."""
    if not oracle_libs.have_ref():
        pytest.skip("oracle/_ref not built")
    import daqp_b200 as d
    ref = C.CDLL(os.path.join(oracle_libs.REF_DIR, "libdaqp_ref.so"))
    H = np.eye(2); f = np.array([2.0, 2.0]); A = np.eye(2); bu = np.ones(2); bl = -np.ones(2)
    x = np.zeros(2); lam = np.zeros(2)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    qp = d.DAQPProblem(2, 2, 0, dp(H), dp(f), dp(A), dp(bu), dp(bl), None, None, 0, 0)
    st = d.DAQPSettings()
    ref.daqp_default_settings(C.byref(st))
    assert (st.primal_tol, st.iter_limit, st.sing_tol) == (1e-6, 10000, 3.7e-11)
    res = d.DAQPResult(dp(x), dp(lam), 0, 0, 0, 0, 0, 0, 0)
    ref.daqp_quadprog(C.byref(res), C.byref(qp), C.byref(st))
    assert res.exitflag == 1
    np.testing.assert_allclose(x, [-1, -1], atol=1e-9)


WS_FIELDS = ["qp", "n", "m", "ms", "M", "dupper", "dlower", "Rinv", "v", "sense", "scaling", "RinvD", "x", "xold", "lam",
             "lam_star", "u", "fval", "L", "D", "xldl", "zldl", "reuse_ind", "WS", "n_active", "iterations", "sing_ind",
             "prox_mask", "n_prox", "soft_slack", "settings", "bnb", "nh", "break_points", "avi", "eq", "timer", "Mu"]


def _offsets(header_dir, header, tmp_path, tag):
    import subprocess
    src = tmp_path / f"off_{tag}.c"
    exe = tmp_path / f"off_{tag}"
    body = "".join(f'printf("%zu ", offsetof(DAQPWorkspace, {f}));' for f in WS_FIELDS)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{header}"\n'
                   f'int main(void) {{ printf("%zu ", sizeof(DAQPWorkspace)); {body} return 0; }}\n')
    subprocess.run(["gcc", "-I", header_dir, "-o", str(exe), str(src)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()


def test_workspace_layout_matches_reference_header(tmp_path):
    """DAQPWorkspace is ABI: the reference's interfaces allocate it and read its fields (daqp.pyx:264, api.jl:444-457).
    Size and every field offset of include/daqp_b200.h against the reference's include/types.h:187-264."""
    ours = _offsets(os.path.join(ROOT, "include"), "daqp_b200.h", tmp_path, "ours")
    assert int(ours[0]) == 288 and len(ours) == len(WS_FIELDS) + 1  # LP64, SOFT_WEIGHTS off
    if os.path.isdir("/root/reference/include"):
        ref = _offsets("/root/reference/include", "types.h", tmp_path, "ref")
        assert ours == ref


def test_default_settings_match_reference_constants(cuda_lib):
    import daqp_b200 as d
    s = d.default_settings()
    got = [getattr(s, k) for k, _ in s._fields_]
    want = [1e-6, 1e-12, 1e-11, 1e-6, 1e-14, 10, 10000, 1e30, -1e-6, -1.0, 1e-6, 0, 0, 3.7e-11, 1e-9, 0]
    assert got == want  # include/constants.h:15-29


def test_no_gpu_means_error_not_fallback(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import daqp_b200 as d
    with pytest.raises(RuntimeError, match="no CUDA device"):
        d.Engine()
    H = np.eye(2); x, fval, flag, info = None, None, None, None
    # the drop-in symbol reports through the exit flag (the reference API has no other channel)
    x, fval, flag, info = d.solve(H, np.ones(2), np.eye(2), np.ones(2), -np.ones(2))
    assert flag == d.EXIT_UNSUPPORTED


def test_product_does_not_reference_oracle():
    """The shipped package (and the helper scripts) must not import, link or call anything under oracle/: only tests/,
    __graft_entry__.smoke() and bench.py's CPU legs may."""
    needles = ("import oracle", "from oracle", "oracle/", "oracle.harness", "libdaqp_oracle", "daqp_oracle",
               "libdaqp_ref", "orc_quadprog")
    for pkg in ("daqp_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    txt = open(os.path.join(dirpath, fn)).read()
                    for needle in needles:
                        assert needle not in txt, f"{pkg}/{fn} reaches into the oracle ({needle})"

"""Decision-trace diff: solve a rare-path fixture on the GPU with the per-problem decision log switched on
(DAQPB200Diag.trace) and in the oracle (OracleLib.solve(log_cap=...)), and print, for every problem whose iteration
count differs, the first decision the two disagree on together with a few decisions either side.

    python tests/trace_diff.py rare_eqpairs_n12_ms4 [more fixtures]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # test tooling: the oracle is the checker here
import daqp_b200  # noqa: E402
from common import load_golden, rare_settings  # noqa: E402
from oracle import harness  # noqa: E402

CODES = {1: "add", 2: "rem", 3: "refactor", 4: "refine", 5: "cycle", 7: "exit"}


def fold(log):
    """(8, hi), (9, lo) pairs -> one ('fval', float) entry."""
    out, k = [], 0
    while k < len(log):
        if int(log[k][0]) == 8 and k + 1 < len(log):
            bits = (int(log[k][1]) & 0xffffffff) << 32 | (int(log[k + 1][1]) & 0xffffffff)
            out.append((8, float(np.array([bits], np.uint64).view(np.float64)[0])))
            k += 2
        else:
            out.append((int(log[k][0]), int(log[k][1])))
            k += 1
    return out


def fmt(e):
    c, v = e
    if c == 8:
        return f"fval {v!r}"
    if c == 1:
        return f"add {v >> 1}{'L' if v & 1 else 'U'}"
    return f"{CODES.get(c, c)} {v}"


def main(names, cap=400):
    eng = daqp_b200.Engine()
    dev = torch.device("cuda:0")
    for name in names:
        b, d = load_golden(name)
        over = rare_settings(d)
        use_sense = bool(d["use_sense"])
        o = harness.OracleLib().solve(b, settings=harness.default_settings(**over), use_sense=use_sense, log_cap=cap)
        t = lambda a, ty=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=ty, device=dev).contiguous()
        ns = int(((b.sense & 8) != 0).sum(axis=1).max(initial=0)) if use_sense else 0
        diag = daqp_b200.Engine.alloc_diag(b.N, b.n, b.m, dev, ns=ns)
        diag["trace"] = torch.zeros((b.N, 1 + 2 * cap), dtype=torch.int32, device=dev)
        out = eng.solve_batch_device(t(b.H), t(b.f), t(b.A), t(b.bupper), t(b.blower),
                                     t(b.sense, torch.int32) if use_sense else None, ms=b.ms, diag=diag, **over)
        torch.cuda.synchronize()
        it = out["iter"].cpu().numpy()
        tr = diag["trace"].cpu().numpy()
        off = np.nonzero(it != d["iter"])[0]
        print(f"{name}: N={b.N} gpu-vs-reference iteration mismatches {off.tolist()}; oracle-vs-reference "
              f"{np.nonzero(o.iter != d['iter'])[0].tolist()}")
        for p in off:
            g = fold(tr[p, 1:1 + 2 * min(int(tr[p, 0]), cap)].reshape(-1, 2))
            r = fold(o.oplog[p])
            k = 0
            while k < min(len(g), len(r)) and g[k] == r[k]:
                k += 1
            print(f"  problem {p}: gpu {len(g)} decisions (iter {it[p]}), oracle {len(r)} (iter {o.iter[p]}); first "
                  f"difference at decision {k}")
            lo = max(0, k - 4)
            print("    gpu   :", ", ".join(fmt(e) for e in g[lo:k + 8]))
            print("    oracle:", ", ".join(fmt(e) for e in r[lo:k + 8]))
            kd = 0  # first decision (objective values aside) that differs
            gd, rd = [e for e in g if e[0] != 8], [e for e in r if e[0] != 8]
            while kd < min(len(gd), len(rd)) and gd[kd] == rd[kd]:
                kd += 1
            print(f"    first differing DECISION: #{kd}: gpu {fmt(gd[kd]) if kd < len(gd) else None}, oracle "
                  f"{fmt(rd[kd]) if kd < len(rd) else None}")
            if "--full" in sys.argv:
                print("    gpu full   :", ", ".join(fmt(e) for e in g))
                print("    oracle full:", ", ".join(fmt(e) for e in r))
    eng.close()


if __name__ == "__main__":
    main([a for a in sys.argv[1:] if not a.startswith("--")] or ["rare_eqpairs_n12_ms4"])

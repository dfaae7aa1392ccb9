"""Informational throughput of the batched minimal-representation path (daqp_b200_minrep_device), next to the reference's
daqp_minrep (oracle/_ref, one thread) on a sample of the same polyhedra. Not a bench.py line: the headline metric is C3.
usage: python scripts/bench_minrep.py [--out gpurun_out/minrep.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import daqp_b200  # noqa: E402
from daqp_b200.problems import generate_polyhedra  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from oracle import harness  # CPU baseline leg only
    eng = daqp_b200.Engine()
    dev = torch.device("cuda:0")
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
        t0 = time.perf_counter()
        red_h = eng.minrep_batch(A, b, ms=ms)
        ms_host = (time.perf_counter() - t0) * 1e3
        red = out["is_redundant"].cpu().numpy()
        assert (red == red_h).all()
        sample = min(P, 16)
        t0 = time.perf_counter()
        ref = np.stack([harness.ref_minrep(A[q], b[q]) for q in range(sample)]) if harness.have_ref() else None
        cpu_s = time.perf_counter() - t0
        row = {"P": P, "n": n, "m": m, "ms": ms, "ldps": P * m, "device_ms": ms_dev, "host_call_ms": ms_host,
               "polyhedra_per_s_device": P / ms_dev * 1e3, "ldps_per_s_device": P * m / ms_dev * 1e3,
               "mean_iterations": float(out["iter"].float().mean()), "redundant_fraction": float(red.mean())}
        if ref is not None:
            row["reference_cpu_polyhedra_per_s_1thread"] = sample / cpu_s
            row["matches_reference_on_sample"] = bool((ref == red[:sample]).all())
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

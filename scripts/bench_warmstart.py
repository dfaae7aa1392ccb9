"""Device time and achieved HBM bandwidth of init_active_kernel (primal mode: one read of A per problem) on the C3 shape.
usage: python scripts/bench_warmstart.py [--out gpurun_out/warmstart.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import daqp_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--N", type=int, default=100000)
    args = ap.parse_args()
    N, n, m, ms = args.N, 50, 150, 0
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    A = torch.randn((N, m - ms, n), dtype=torch.float64, device=dev, generator=g)
    x = torch.randn((N, n), dtype=torch.float64, device=dev, generator=g)
    ax = torch.einsum("bmn,bn->bm", A, x)
    bu = ax + (torch.rand((N, m), dtype=torch.float64, device=dev, generator=g) < 0.25) * 1.0  # a quarter of the rows tight
    bl = bu - 2.0
    sense = torch.zeros((N, m), dtype=torch.int32, device=dev)
    eng = daqp_b200.Engine()
    rows = {}
    for mode, kw in (("primal", dict(x=x)), ("dual", dict(lam=ax))):
        for _ in range(3):
            eng.init_active_device(sense, A, bu, bl, ms=ms, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            eng.init_active_device(sense, A, bu, bl, ms=ms, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms_k = e0.elapsed_time(e1) / reps
        bytes_alg = N * ((m - ms) * n * 8 + n * 8 + 2 * m * 8 + 2 * m * 4) if mode == "primal" else N * (m * 8 + 2 * m * 4)
        peak = 6553.3
        pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pj):
            peak = json.load(open(pj)).get("hbm_gbs", peak)
        rows[mode] = {"N": N, "n": n, "m": m, "kernel_ms": ms_k, "algorithmic_bytes": bytes_alg,
                      "achieved_gbs": bytes_alg / ms_k / 1e6, "peak_gbs": peak, "frac": bytes_alg / ms_k / 1e6 / peak,
                      "problems_per_s": N / ms_k * 1e3}
        print(mode, json.dumps(rows[mode]), flush=True)
    tight = float(((sense & 1) != 0).float().mean())
    rows["active_fraction_set"] = tight
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

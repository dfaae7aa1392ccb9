#!/usr/bin/env python
"""bench.py -- QPs solved per second on BASELINE.json's headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--problems P]

Workload (config C3, the configuration BASELINE.json's metric is quoted on): per GPU a batch of 100 000 random dense
QPs, n=50, m=150, fp64, built by the port of the reference's own test generator generate_test_QP (kappa=100,
nActive=0.8 n), so the optimum of every problem is known by construction. One "step" = one pass of the hot path
(qp_setup_kernel + ldp_solve_kernel) over the whole batch.

  value     QP/s with the batch already resident in HBM (CUDA events around K steps, max over ranks)
  e2e       QP/s through the host C ABI daqp_b200_solve_packed: pinned HOST buffers, H2D/D2H copies inside the timed
            region (plus, for N>1, an NCCL gather of the solutions to rank 0)
  roofline  ldp_solve_kernel: algorithmic bytes of the streaming model over the kernel's
            own CUDA-event time, against the measured HBM copy peak
  cpu_baseline  the reference CPU solver (oracle/_ref, compiled from the reference sources) on a bounded sample of
            the SAME problems, 1 thread, on this box's host cores

--impl reference times the reference's own CPU implementation with all host threads on bounded samples of the same
workload. Multi-GPU: one process per GPU under torchrun; problems are independent, so the batch is sharded with no
data-path collective (weak scaling: 100k problems per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QPs solved/sec on 100k-batch n=50 m=150 random dense QPs (fp64)"
UNIT = "QP/s"
CFG = dict(n=50, m=150, ms=0, n_active=40, kappa=100.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        top = sorted(busy)[len(busy) // 2:] if busy else []
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_RESULT_FD = None


def claim_stdout():
    """stdout carries ONE line: the JSON result. Libraries loaded later write there too (NCCL prints its version banner to
    stdout when NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the rest of the run and the
    result line goes to a private duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def dist_setup(gpus: int):
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    return rank, world, local


def algorithmic_bytes(counts, n, m, w=8):
    """Streaming model per QP for ldp_solve_kernel: every feasibility scan streams the whole
    constraint matrix and both bound vectors, every LDL add re-reads the entering row, results are written once.
    counts[:,0] = scans, counts[:,1] = adds (the kernel's own per-problem counters)."""
    scans = int(counts[:, 0].sum()); adds = int(counts[:, 1].sum()); N = counts.shape[0]
    return scans * w * (m * n + 2 * m) + adds * w * n + N * (w * (n + m) + 12)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU path (oracle/_ref when present, else the oracle port) with all host
    threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import harness
    from daqp_b200.problems import generate_g1, generate_g1_torch, torch_to_batch, SEED_BASE
    cores = os.cpu_count() or 1
    per_step = args.ref_sample or max(2048, 256 * cores)
    try:
        import torch
        use_cuda = torch.cuda.is_available()
    except Exception:
        use_cuda = False
    if use_cuda:  # literally the first problems of the GPU arm's batch
        t = generate_g1_torch(per_step, CFG["n"], CFG["m"], CFG["ms"], CFG["n_active"], CFG["kappa"],
                              seed=SEED_BASE + 3, device="cuda:0")
        b = torch_to_batch(t, CFG["n"], CFG["m"], CFG["ms"])
    else:
        b = generate_g1(per_step, CFG["n"], CFG["m"], CFG["ms"], CFG["n_active"], CFG["kappa"], seed=SEED_BASE + 3)
    if harness.have_ref("libref_driver.so"):
        drv, kind = harness.RefDriver(), "reference"
    else:
        harness.build(ref=False)
        drv, kind = harness.OracleLib(), "port"
    for _ in range(args.warmup):
        drv.solve_packed(b.slice(0, min(b.N, 64 * cores)), nthreads=cores)
    secs = 0.0
    for _ in range(args.steps):
        s = drv.solve_packed(b, nthreads=cores)
        assert (s.exitflag == 1).all()
        secs += s.seconds
    val = b.N * args.steps / secs
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, b.N, note=f"reference CPU path, {b.N}-problem sample per step"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{b.N} problems of the C3 batch per step x {args.steps} steps"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def config_dict(args, per_gpu, note=None):
    c = {"workload": "C3: batch of 100000 random dense QPs n=50 m=150 ms=0 fp64 per GPU (generate_test_QP port, "
                     "kappa=100, nActive=40)", "n": CFG["n"], "m": CFG["m"], "ms": CFG["ms"],
         "problems_per_gpu": per_gpu, "generator": "G1 (reference interfaces/daqp-julia/test/utils.jl:3-53)",
         "parallelism": f"batch sharded over {args.gpus} GPU(s), no data-path collective",
         "cache": "inputs (8.5 GB) + LDP scratch (16.6 GB) per step are larger than L2 (126 MB); no explicit flush"}
    if note:
        c["note"] = note
    return c


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=100_000, help="problems per GPU (default: the C3 batch)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--ref-sample", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workspace", action="store_true", help="skip the persistent-workspace (MPC step) leg")
    args = ap.parse_args()

    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference(args, rank, world)
        if world > 1:
            import torch.distributed as dist
            dist.barrier(); dist.destroy_process_group()
        return

    import numpy as np
    import torch
    import daqp_b200
    from daqp_b200 import build
    from daqp_b200.problems import SEED_BASE, generate_g1_torch, torch_to_batch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: daqp_b200 has no CPU path")
    build.build()
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    n, m, ms = CFG["n"], CFG["m"], CFG["ms"]
    P = args.problems
    # every rank owns its own shard (born sharded): same construction, rank-specific seed
    t = generate_g1_torch(P, n, m, ms, CFG["n_active"], CFG["kappa"], seed=SEED_BASE + 3 + 1000 * rank, device=dev)
    eng = daqp_b200.Engine(local)
    out = None
    diag = eng.alloc_diag(P, n, m, dev)

    def step(d=None):
        nonlocal out
        out = eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], None, ms=ms, out=out, diag=d)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate on the real workload: construction-known optimum + counters for the byte model
    step(diag)
    torch.cuda.synchronize()
    assert bool((out["exitflag"] == 1).all()), "not all problems OPTIMAL"
    err = float((out["x"] - t["xref"]).abs().max())
    assert err < 1e-6, f"x differs from the constructed optimum by {err}"  # reference gate: 1e-4 (core_tests.jl:26-30)
    # problems whose final active set differs from the constructed one (possible only where the construction is
    # degenerate: a multiplier drawn ~0); reported, and bounded
    as_mismatch = int((torch.sign(out["lam"]).to(torch.int8) != t["active_ref"]).any(dim=1).sum())
    assert as_mismatch <= max(1, P // 10000), f"{as_mismatch} problems end on a different active set"
    counts = diag["counts"].cpu().numpy()
    iters_mean = float(out["iter"].double().mean())

    for _ in range(args.warmup):
        step()
    barrier()
    eng.stats(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    st = eng.stats(reset=True)
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    value = world * P * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (ldp_solve_kernel), from its own CUDA-event time on the launch stream
    peak, peak_src = peaks()
    alg = algorithmic_bytes(counts, n, m)
    solve_ms = st["solve_ms"] / max(1, st["solve_launches"])
    achieved = alg / (solve_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("ldp_solve_kernel_dram_bytes_per_qp", None)
            if traffic is not None:
                traffic = traffic * P  # per launch, like `achieved`
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "ldp_solve_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_per_qp": alg / P,
                "kernel_ms_per_launch": solve_ms, "setup_kernel_ms_per_launch": st["setup_ms"] / max(1, st["setup_launches"]),
                "solve_share_of_step": st["solve_ms"] / (st["solve_ms"] + st["setup_ms"]),
                "resident_problems_per_sm": st["warps_per_sm"], "mean_iterations": iters_mean,
                "scans_per_qp": float(counts[:, 0].mean()), "adds_per_qp": float(counts[:, 1].mean()),
                "removes_per_qp": float(counts[:, 2].mean())}

    # ---- end to end through the host C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    wsp = None
    shw = None
    if not args.no_e2e:
        pin = lambda x: x.cpu().pin_memory()
        h = {k: pin(t[k]) for k in ("H", "f", "A", "bupper", "blower")}
        res = daqp_b200.BatchResult(x=torch.empty((P, n), dtype=torch.float64).pin_memory().numpy(),
                                    lam=torch.empty((P, m), dtype=torch.float64).pin_memory().numpy(),
                                    fval=torch.zeros(P, dtype=torch.float64).pin_memory().numpy(),
                                    exitflag=torch.empty(P, dtype=torch.int32).pin_memory().numpy(),
                                    iter=torch.empty(P, dtype=torch.int32).pin_memory().numpy())
        hn = {k: v.numpy() for k, v in h.items()}
        gather_buf = None
        if world > 1:
            import torch.distributed as dist
            xg = torch.empty((P, n), dtype=torch.float64, device=dev)
            gather_buf = [torch.empty_like(xg) for _ in range(world)] if rank == 0 else None

        def e2e_step():
            eng.solve_batch(hn["H"], hn["f"], hn["A"], hn["bupper"], hn["blower"], None, ms=ms, out=res)
            if world > 1:
                xg.copy_(torch.from_numpy(res.x), non_blocking=True)
                dist.gather(xg, gather_buf, dst=0)

        e2e_step()
        assert (res.exitflag == 1).all() and np.abs(res.x - t["xref"].cpu().numpy()).max() < 1e-6
        ksteps = args.e2e_steps or max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        h2d = sum(v.numel() * v.element_size() for v in h.values())
        d2h = res.x.nbytes + res.lam.nbytes + res.fval.nbytes + res.exitflag.nbytes + res.iter.nbytes
        e2e = {"value": world * P * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": ksteps, "api": "daqp_b200_solve_packed (host C ABI, pinned buffers, chunked copy/solve overlap)"}
        # ---- persistent workspace: setup once, then update(f, b) + warm solve per "MPC step".
        # Reported next to the headline, not part of it: an extra object in the same line.
        wsp = None
        if rank == 0 and world == 1 and not args.no_workspace:
            try:
                rng = np.random.default_rng(7)
                mdl = daqp_b200.BatchModel(eng).setup(hn["H"], hn["f"], hn["A"], hn["bupper"], hn["blower"], None, ms=ms)
                r0 = mdl.solve(warm=True)
                assert (r0.exitflag == 1).all()

                def perturbed():  # pinned, like the e2e leg's inputs
                    sh = 0.01 * rng.standard_normal(hn["bupper"].shape)
                    arrs = (hn["f"] * (1 + 0.05 * rng.standard_normal(hn["f"].shape)), hn["bupper"] + sh, hn["blower"] + sh)
                    return tuple(torch.from_numpy(a).pin_memory().numpy() for a in arrs)

                steps_ws = [perturbed() for _ in range(4)]
                mdl.update(*steps_ws[0]); mdl.solve(warm=True, out=res)  # warm-up step
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                its, opt = [], []
                for fk, buk, blk in steps_ws[1:]:
                    mdl.update(fk, buk, blk)
                    rk = mdl.solve(warm=True, out=res)
                    its.append(float(rk.iter.mean())); opt.append(float((rk.exitflag == 1).mean()))
                dtw = (time.perf_counter() - t0) / (len(steps_ws) - 1)
                wsp = {"value": P / dtw, "unit": "QP/s per warm step (update f,b from host + solve + results to host)",
                       "ms_per_step": 1e3 * dtw, "mean_iterations_warm": sum(its) / len(its),
                       "mean_iterations_cold": iters_mean, "optimal_fraction": sum(opt) / len(opt),
                       "perturbation": "f * (1 + 0.05 N(0,1)), bounds + 0.01 N(0,1) per step",
                       "api": "daqp_b200_workspace_update + daqp_b200_workspace_solve(warm=1)"}
                mdl.close()
            except Exception as ex:  # the leg is informative: never lose the headline line over it
                wsp = {"error": repr(ex)[:200]}
        # ---- shared workspace (one controller, many states): G matrix sets, P / G problems per set that differ only in
        # f and the bounds; every step is a COLD solve of all P problems with f / bounds copied from pinned host memory.
        shw = None
        if rank == 0 and world == 1 and not args.no_workspace:
            try:
                rng = np.random.default_rng(11)
                G = 16
                K = P // G
                NS = G * K
                mdl = daqp_b200.BatchModel(eng).setup_shared(hn["H"][:G], hn["A"][:G], K, ms=ms, m=m)

                def draw():
                    sh = 0.05 * rng.standard_normal((NS, m))
                    arrs = (np.repeat(hn["f"][:G], K, axis=0) * (1 + 0.3 * rng.standard_normal((NS, n))),
                            np.repeat(hn["bupper"][:G], K, axis=0) + sh, np.repeat(hn["blower"][:G], K, axis=0) + sh)
                    return tuple(torch.from_numpy(a).pin_memory().numpy() for a in arrs)

                draws = [draw() for _ in range(3)]
                res_s = daqp_b200.BatchResult(x=res.x[:NS], lam=res.lam[:NS], fval=res.fval[:NS],
                                              exitflag=res.exitflag[:NS], iter=res.iter[:NS])
                mdl.update(*draws[0]); mdl.solve(warm=False, out=res_s)  # warm-up step
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                its, opt = [], []
                for fk, buk, blk in draws[1:]:
                    mdl.update(fk, buk, blk)
                    rk = mdl.solve(warm=False, out=res_s)
                    its.append(float(rk.iter.mean())); opt.append(float((rk.exitflag == 1).mean()))
                dts = (time.perf_counter() - t0) / (len(draws) - 1)
                shw = {"value": NS / dts, "unit": "QP/s per cold step (f,b from host + solve + results to host)",
                       "ms_per_step": 1e3 * dts, "matrix_sets": G, "problems_per_set": K,
                       "h2d_bytes_per_step": sum(a.nbytes for a in draws[0]), "mean_iterations": sum(its) / len(its),
                       "optimal_fraction": sum(opt) / len(opt),
                       "perturbation": "f * (1 + 0.3 N(0,1)), bounds + 0.05 N(0,1) around the set's base problem",
                       "api": "daqp_b200_workspace_setup_shared, then daqp_b200_workspace_update + _solve(warm=0)"}
                mdl.close()
            except Exception as ex:
                shw = {"error": repr(ex)[:200]}
        del h, hn

    # ---- the reference CPU solver on a bounded sample of the SAME problems (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import harness
        S = min(P, args.cpu_sample)
        b = torch_to_batch(t, n, m, ms, 0, S)
        if harness.have_ref("libref_driver.so"):
            drv, kind = harness.RefDriver(), "reference"
        else:
            harness.build(ref=False)
            drv, kind = harness.OracleLib(), "port"
        s = drv.solve_packed(b, nthreads=1)
        x_gpu = out["x"][:S].cpu().numpy()
        assert np.array_equal(s.exitflag, out["exitflag"][:S].cpu().numpy())
        assert np.array_equal(s.iter, out["iter"][:S].cpu().numpy()), "iteration counts differ from the CPU reference"
        assert np.abs(s.x - x_gpu).max() < 1e-9 * (1 + np.abs(s.x).max())
        cpu = {"value": S / s.seconds, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {S} problems of the timed batch, daqp_quadprog per problem, 1 thread; "
                         "exit flags, iteration counts and x checked against the GPU results",
               "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, P), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": st["setup_launches"] + st["solve_launches"], "clocks": clocks,
                "workspace": wsp if not args.no_e2e else None,
                "shared_workspace": shw if not args.no_e2e else None,
                "parity": {"max_abs_x_err_vs_constructed_optimum": err, "all_optimal": True,
                           "active_set_differs_from_construction": as_mismatch}}
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()

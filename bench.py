#!/usr/bin/env python
"""bench.py -- QPs solved per second on BASELINE.json's headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--problems P]

Workload (config C3, the configuration BASELINE.json's metric is quoted on): ONE batch of 100 000 random dense QPs,
n=50, m=150, fp64, built by the port of the reference's own test generator generate_test_QP (kappa=100, nActive=0.8 n),
so the optimum of every problem is known by construction. One "step" = one pass of the hot path (QP -> LDP transform +
dual active-set solve) over the whole batch. With --gpus N the batch is SPLIT over the N GPUs (12 500 problems per GPU
at 8: strong scaling, BASELINE.json configs[2]); problems are independent, so there is no data-path collective.

  value     QP/s with the batch already resident in HBM (CUDA events around K steps, max over ranks)
  weak      (N > 1) the same with 100 000 problems PER GPU
  e2e       QP/s through the host C ABI daqp_b200_solve_packed: pinned HOST buffers allocated NUMA-local to each rank's
            GPU, H2D/D2H copies inside the timed region (plus, for N > 1, an NCCL gather of the solutions to rank 0)
  roofline  ldp_solve_kernel: algorithmic bytes of the streaming model over the kernel's
            own CUDA-event time, against the measured HBM copy peak; next to it what binds the kernel in fact (DRAM
            bytes, L2 bytes, issue-slot utilisation from the ncu pass of this binary, profiles/ncu_solve_r02.json)
  cpu_baseline  the reference CPU solver (oracle/_ref, compiled from the reference sources) on a bounded sample of
            the SAME problems, 1 thread, on this box's host cores
  c4 / c5 / aos / latency / workspace / shared_workspace   (N = 1) the other BASELINE.json configurations and entry points

--impl reference times the reference's own CPU implementation with all host threads on bounded samples of the same
workload.
"""
from __future__ import annotations

import argparse
import ctypes
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
C4 = dict(n=120, m=400, ms=120, n_active=96, kappa=100.0, N=50_000)


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
        return self

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


def bind_numa(local: int) -> dict:
    """Run this rank on the cores of its GPU's NUMA node and prefer that node for the pages it touches next (the pinned
    staging of the e2e leg): eight ranks pulling 8 GB per step each through one node's memory controllers is what made
    the round-1 end-to-end numbers collapse at N >= 4. Best effort; reports what it managed."""
    info = {"gpu_numa_node": None, "cpus_bound": None, "mempolicy": None}
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = len(allowed)
        try:  # set_mempolicy(MPOL_PREFERRED, {node}): pages first touched from now on come from the GPU's node
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(238, 1, ctypes.byref(mask), 16 * 64 + 1)
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        except Exception as ex:
            info["mempolicy"] = repr(ex)[:60]
    except Exception as ex:
        info["error"] = repr(ex)[:120]
    return info


def algorithmic_bytes(counts, n, m, w=8):
    """Streaming model per QP for ldp_solve_kernel: every feasibility scan streams the whole
    constraint matrix and both bound vectors, every LDL add re-reads the entering row, results are written once.
    counts[:,0] = scans, counts[:,1] = adds (the kernel's own per-problem counters)."""
    scans = int(counts[:, 0].sum()); adds = int(counts[:, 1].sum()); N = counts.shape[0]
    return scans * w * (m * n + 2 * m) + adds * w * n + N * (w * (n + m) + 12)


def cpu_driver():
    from oracle import harness
    if harness.have_ref("libref_driver.so"):
        return harness.RefDriver(), "reference"
    harness.build(ref=False)
    return harness.OracleLib(), "port"


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU path (oracle/_ref when present, else the oracle port) with all host
    threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    import numpy as np
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
    drv, kind = cpu_driver()
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
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, b.N, 1, note=f"reference CPU path, {b.N}-problem sample per step"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{b.N} problems of the C3 batch per step x {args.steps} steps"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def config_dict(args, total, world, note=None):
    c = {"workload": "C3: ONE batch of 100000 random dense QPs n=50 m=150 ms=0 fp64 (generate_test_QP port, kappa=100, "
                     "nActive=40), split over the GPUs", "n": CFG["n"], "m": CFG["m"], "ms": CFG["ms"],
         "problems_total": total, "problems_per_gpu": total // max(1, world),
         "generator": "G1 (reference interfaces/daqp-julia/test/utils.jl:3-53)",
         "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
         "cache": "inputs (85 KB per problem) + LDP scratch (166 KB per problem) per step are larger than L2 (126 MB) "
                  "down to 12500 problems per GPU; no explicit flush"}
    if note:
        c["note"] = note
    return c


def timed_steps(step, steps, warmup, barrier, torch):
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def leg_c4(eng, dev, peak, args):
    """BASELINE.json config 4: MPC-style QPs n=120, m=400 (box on every variable + 280 general rows), 50 000-batch, cold
    and warm-started from the active set of a neighbour whose linear term differs by 5 % (an MPC step)."""
    import numpy as np
    import torch
    import daqp_b200
    from daqp_b200.problems import SEED_BASE, generate_g1_torch, torch_to_batch
    n, m, ms, N = C4["n"], C4["m"], C4["ms"], args.c4_problems
    t = generate_g1_torch(N, n, m, ms, C4["n_active"], C4["kappa"], seed=SEED_BASE + 4, device=dev)
    diag = eng.alloc_diag(N, n, m, dev)
    sampler = ClockSampler(dev.index or 0).start()
    out = {}

    def run(sense, d=None):
        return eng.solve_batch_device(t["H"], t["f"], t["A"], t["bupper"], t["blower"], sense, ms=ms, diag=d)

    r = run(None, diag); torch.cuda.synchronize()
    assert bool((r["exitflag"] == 1).all()) and float((r["x"] - t["xref"]).abs().max()) < 1e-5
    counts = diag["counts"].cpu().numpy()
    eng.stats(reset=True)
    ms_cold = timed_steps(lambda: run(None), 2, 1, torch.cuda.synchronize, torch) / 2
    st = eng.stats(reset=True)
    alg = algorithmic_bytes(counts, n, m)
    solve_ms = st["solve_ms"] / 3
    out["cold"] = {"value": N / ms_cold * 1e3, "unit": UNIT, "ms_per_batch": ms_cold, "mean_iterations": float(r["iter"].double().mean()),
                   "setup_ms": st["setup_ms"] / 3, "solve_ms": solve_ms, "resident_problems_per_sm": st["warps_per_sm"],
                   "roofline": {"bound": "hbm", "kernel": "ldp_solve_kernel (team mode)", "achieved": alg / (solve_ms * 1e-3) / 1e9,
                                "peak": peak, "unit": "GB/s", "frac": alg / (solve_ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_qp": alg / N, "scans_per_qp": float(counts[:, 0].mean())}}
    g = torch.Generator(device=dev); g.manual_seed(44)
    fn_ = t["f"] * (1 + 0.05 * torch.randn(t["f"].shape, dtype=torch.float64, device=dev, generator=g))
    rn = eng.solve_batch_device(t["H"], fn_, t["A"], t["bupper"], t["blower"], None, ms=ms)
    torch.cuda.synchronize()
    sense = torch.zeros((N, m), dtype=torch.int32, device=dev)
    sense[rn["lam"] > 1e-12] = 1
    sense[rn["lam"] < -1e-12] = 3
    del rn, fn_
    rw = run(sense, diag); torch.cuda.synchronize()
    assert bool((rw["exitflag"] == 1).all()) and float((rw["x"] - t["xref"]).abs().max()) < 1e-5
    eng.stats(reset=True)
    ms_warm = timed_steps(lambda: run(sense), 2, 1, torch.cuda.synchronize, torch) / 2
    st = eng.stats(reset=True)
    out["warm"] = {"value": N / ms_warm * 1e3, "unit": UNIT, "ms_per_batch": ms_warm, "mean_iterations": float(rw["iter"].double().mean()),
                   "setup_ms": st["setup_ms"] / 3, "solve_ms": st["solve_ms"] / 3,
                   "warm_start": "sense bits = optimal active set of a neighbour whose f differs by 5 % N(0,1)"}
    out["clocks"] = sampler.stop()
    # the reference on a sample of the same problems, 1 thread, cold and warm; flags and iteration counts cross-checked
    if not args.no_cpu:
        S = min(N, args.c4_cpu_sample)
        drv, kind = cpu_driver()
        b = torch_to_batch(t, n, m, ms, 0, S)
        sc = drv.solve_packed(b, nthreads=1)
        assert np.array_equal(sc.exitflag, r["exitflag"][:S].cpu().numpy()) and np.array_equal(sc.iter, r["iter"][:S].cpu().numpy()), \
            "C4 cold: flags / iteration counts differ from the CPU reference"
        b.sense[:] = sense[:S].cpu().numpy()
        sw = drv.solve_packed(b, nthreads=1, use_sense=True)
        assert np.array_equal(sw.exitflag, rw["exitflag"][:S].cpu().numpy()) and np.array_equal(sw.iter, rw["iter"][:S].cpu().numpy()), \
            "C4 warm: flags / iteration counts differ from the CPU reference"
        out["cpu_baseline"] = {"cold": S / sc.seconds, "warm": S / sw.seconds, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": f"first {S} problems of the batch, daqp_quadprog per problem, 1 thread; exit flags and "
                                         "iteration counts equal to the GPU's"}
    out["config"] = {"workload": f"C4: {N} MPC-style QPs n=120 m=400 ms=120 fp64, G1 generator, nActive=96", "api": "daqp_b200_solve_device"}
    del t, r, rw, sense, diag
    torch.cuda.empty_cache()
    return out


def leg_c5(dev, args):
    """BASELINE.json config 5: mixed-size batch, n in {8,16,...,128}, m = 4 n, the number of active constraints at the
    optimum drawn from U{0..n} (divergent iteration counts), 200 000 problems, fp32. Device-resident: the sixteen shape
    groups are dealt, largest estimated cost first, to three engines (own streams and scratch) that run side by side."""
    import torch
    import daqp_b200
    from daqp_b200.problems import generate_g1_torch
    sizes = list(range(8, 129, 8))
    per = args.c5_problems // len(sizes)
    groups = []
    for k, n in enumerate(sizes):
        t = generate_g1_torch(per, n, 4 * n, 0, n, seed=500 + k, device=dev, random_nactive=True)
        groups.append({k2: (v.float().contiguous() if v.dtype == torch.float64 else v) for k2, v in t.items()})
        del t
    torch.cuda.empty_cache()
    lanes = [daqp_b200.Engine(dev.index or 0) for _ in range(3)]
    streams = [torch.cuda.Stream(device=dev) for _ in lanes]
    order = sorted(range(len(groups)), key=lambda i: -(sizes[i] ** 2) * 4 * sizes[i])

    def run():
        outs = [None] * len(groups)
        torch.cuda.synchronize()
        for j, gi in enumerate(order):
            gp = groups[gi]
            with torch.cuda.stream(streams[j % len(lanes)]):
                outs[gi] = lanes[j % len(lanes)].solve_batch_device_f32(gp["H"], gp["f"], gp["A"], gp["bupper"], gp["blower"], None, ms=0)
        torch.cuda.synchronize()
        return outs

    rs = run()
    t0 = time.perf_counter()
    for _ in range(2):
        rs = run()
    dt = (time.perf_counter() - t0) / 2
    Ntot = per * len(sizes)
    opt = sum(int((r["exitflag"] == 1).sum()) for r in rs)
    err = max(float(((r["x"] - gp["xref"]).abs().max(dim=1).values / (1 + gp["xref"].abs().max(dim=1).values))[r["exitflag"] == 1].max())
              for r, gp in zip(rs, groups))
    its = torch.cat([r["iter"] for r in rs]).double()
    for e in lanes:
        e.close()
    del groups, rs
    torch.cuda.empty_cache()
    return {"value": Ntot / dt, "unit": UNIT, "ms_per_batch": 1e3 * dt, "problems": Ntot, "dtype": "f32",
            "optimal_fraction": opt / Ntot, "max_rel_x_err_vs_constructed_optimum": err,
            "iterations_min_mean_max": [float(its.min()), float(its.mean()), float(its.max())],
            "config": {"workload": f"C5: {Ntot} QPs, n in 8..128 step 8 ({per} each), m = 4 n, nActive ~ U{{0..n}}, fp32",
                       "schedule": "16 shape groups, largest n^2 m first, on 3 engines / streams side by side (wall clock)",
                       "api": "daqp_b200_solve_device_f32"}}


def leg_aos_latency(hn, n, m, ms, args):
    """The literal drop-in entries: daqp_quadprog_batch (array of the reference's structs, host memory, copies included)
    on the first problems of the C3 batch, and the latency of a single daqp_quadprog call at C1 (n=10, m=20)."""
    import numpy as np
    import ctypes as C
    import daqp_b200 as d
    from daqp_b200.problems import generate_g1
    L = d.lib()
    N = min(args.aos_problems, hn["H"].shape[0])
    qps = (d.DAQPProblem * N)(); res = (d.DAQPResult * N)()
    x = np.empty((N, n)); lam = np.empty((N, m))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for i in range(N):
        qps[i] = d.DAQPProblem(n, m, ms, dp(hn["H"][i]), dp(hn["f"][i]), dp(hn["A"][i]), dp(hn["bupper"][i]), dp(hn["blower"][i]),
                               None, None, 0, 0)
        res[i] = d.DAQPResult(dp(x[i]), dp(lam[i]), 0, 0, 0, 0, 0, 0, 0)
    st = d.default_settings()
    assert L.daqp_quadprog_batch(N, qps, res, C.byref(st)) == 0
    t0 = time.perf_counter()
    assert L.daqp_quadprog_batch(N, qps, res, C.byref(st)) == 0
    dt = time.perf_counter() - t0
    assert all(res[i].exitflag == 1 for i in range(0, N, max(1, N // 64)))
    aos = {"value": N / dt, "unit": UNIT, "problems": N, "api": "daqp_quadprog_batch (array of DAQPProblem, pageable host memory, "
           "packing + copies + results inside the timed call)"}
    b = generate_g1(1, 10, 20, 0, 8, seed=11)
    x1 = np.empty(10); l1 = np.empty(20)
    qp = d.DAQPProblem(10, 20, 0, dp(b.H[0]), dp(b.f[0]), dp(b.A[0]), dp(b.bupper[0]), dp(b.blower[0]), None, None, 0, 0)
    r1 = d.DAQPResult(dp(x1), dp(l1), 0, 0, 0, 0, 0, 0, 0)
    for _ in range(20):
        L.daqp_quadprog(C.byref(r1), C.byref(qp), C.byref(st))
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); L.daqp_quadprog(C.byref(r1), C.byref(qp), C.byref(st)); ts.append(time.perf_counter() - t0)
    lat = {"median_us": 1e6 * statistics.median(ts), "p90_us": 1e6 * sorted(ts)[180], "config": "C1: one QP n=10 m=20 through daqp_quadprog",
           "exitflag": int(r1.exitflag), "iterations": int(r1.iter)}
    try:
        from oracle import harness
        if harness.have_ref():
            ref = C.CDLL(os.path.join(harness.REF_DIR, "libdaqp_ref.so"))
            tr = []
            for _ in range(200):
                t0 = time.perf_counter(); ref.daqp_quadprog(C.byref(r1), C.byref(qp), C.byref(st)); tr.append(time.perf_counter() - t0)
            lat["reference_cpu_median_us"] = 1e6 * statistics.median(tr)
    except Exception:
        pass
    return aos, lat


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=100_000, help="size of the ONE batch that is split over the GPUs (default: C3)")
    ap.add_argument("--weak-problems", type=int, default=100_000, help="problems per GPU of the weak-scaling leg (N > 1)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--ref-sample", type=int, default=None)
    ap.add_argument("--c4-problems", type=int, default=C4["N"])
    ap.add_argument("--c4-cpu-sample", type=int, default=384)
    ap.add_argument("--c5-problems", type=int, default=200_000)
    ap.add_argument("--aos-problems", type=int, default=20_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workspace", action="store_true", help="skip the persistent-workspace (MPC step) legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the C4 / C5 / AoS / latency legs")
    ap.add_argument("--no-weak", action="store_true")
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
    from daqp_b200.sharding import partition
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: daqp_b200 has no CPU path")
    build.build()
    numa = bind_numa(local)  # before anything is pinned
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    n, m, ms = CFG["n"], CFG["m"], CFG["ms"]
    lo, hi = partition(args.problems, world)[rank]
    P = hi - lo
    # every rank owns its block of the batch (born sharded): same construction, rank-specific seed
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

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        import torch.distributed as dist
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def gather_floats(v: float) -> list:
        if world == 1:
            return [v]
        import torch.distributed as dist
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        outl = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(outl, tt)
        return [float(o.item()) for o in outl]

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
    sampler = ClockSampler(local).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_rank = e0.elapsed_time(e1)
    st = eng.stats(reset=True)  # kernel times and launch counts of the TIMED steps only
    # the timed region can be shorter than nvidia-smi's first sample (16 ms per step at 8 GPUs): keep the same load running,
    # untimed, until the sampler has seen it
    t_probe = time.perf_counter()
    while len(sampler.rows) < 8 and time.perf_counter() - t_probe < 3.0:
        step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["note"] = "sampled every 100 ms over the timed steps and the identical untimed steps that follow them"
    barrier()
    eng.stats(reset=True)
    ms_total = max_over_ranks(ms_rank)
    value = args.problems * args.steps / (ms_total * 1e-3)
    per_rank_ms = gather_floats(ms_rank / args.steps)

    # ---- roofline of the dominant kernel (ldp_solve_kernel), from its own CUDA-event time on the launch stream
    peak, peak_src = peaks()
    alg = algorithmic_bytes(counts, n, m)
    solve_ms = st["solve_ms"] / max(1, st["solve_launches"])
    setup_ms = st["setup_ms"] / max(1, st["solve_launches"])
    achieved = alg / (solve_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "ldp_solve_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "model": "streaming model of SURVEY 8d: every feasibility scan streams the matrix (algorithmic bytes); the kernel "
                         "keeps the screening copy L2-resident, so what binds it in fact is issue slots / dependent latency",
                "binding_resource": "issue/latency",
                "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_per_qp": alg / P,
                "kernel_ms_per_launch": solve_ms, "setup_kernels_ms_per_launch": setup_ms,
                "solve_share_of_step": st["solve_ms"] / (st["solve_ms"] + st["setup_ms"]),
                "resident_problems_per_sm": st["warps_per_sm"], "mean_iterations": iters_mean,
                "scans_per_qp": float(counts[:, 0].mean()), "adds_per_qp": float(counts[:, 1].mean()),
                "removes_per_qp": float(counts[:, 2].mean())}
    mp = os.path.join(ROOT, "profiles", "ncu_solve_r02.json")
    if os.path.exists(mp):  # one `ncu --set full` pass of this binary (scripts/gpu_prof.sh), per QP so that it scales with the launch
        try:
            mj = json.load(open(mp))
            dram = mj["dram_bytes_per_qp"] * P
            roofline["traffic"] = dram
            roofline["dram_frac"] = dram / (solve_ms * 1e-3) / 1e9 / peak
            roofline["l2_gbs"] = mj["lts_bytes_per_qp"] * P / (solve_ms * 1e-3) / 1e9
            roofline["issue_active_pct"] = mj.get("issue_active_pct")
            roofline["warp_instructions_per_iteration"] = mj.get("warp_instructions_per_iteration")
            roofline["ncu_source"] = mj.get("source")
        except Exception as ex:
            roofline["ncu_source"] = "unreadable: " + repr(ex)[:80]

    # ---- weak scaling (N > 1): 100 000 problems per GPU, the round-1 headline, kept next to the strong number
    weak = None
    if world > 1 and not args.no_weak:
        Pw = args.weak_problems
        tw = generate_g1_torch(Pw, n, m, ms, CFG["n_active"], CFG["kappa"], seed=SEED_BASE + 3 + 1000 * rank, device=dev)
        ow = None

        def wstep():
            nonlocal ow
            ow = eng.solve_batch_device(tw["H"], tw["f"], tw["A"], tw["bupper"], tw["blower"], None, ms=ms, out=ow)

        ksw = max(1, min(args.steps, 3))
        msw = max_over_ranks(timed_steps(wstep, ksw, 2, barrier, torch))
        assert bool((ow["exitflag"] == 1).all())
        weak = {"value": world * Pw * ksw / (msw * 1e-3), "unit": UNIT, "problems_per_gpu": Pw, "steps": ksw, "ms_per_step": msw / ksw}
        del tw, ow
        torch.cuda.empty_cache()

    # ---- end to end through the host C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    wsp = None
    shw = None
    hn = None
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
            width = max(b - a for a, b in partition(args.problems, world))
            xg = torch.zeros((width, n), dtype=torch.float64, device=dev)
            gather_buf = [torch.empty_like(xg) for _ in range(world)] if rank == 0 else None
        solve_secs = [0.0]

        def e2e_step():
            t0 = time.perf_counter()
            eng.solve_batch(hn["H"], hn["f"], hn["A"], hn["bupper"], hn["blower"], None, ms=ms, out=res)
            solve_secs[0] += time.perf_counter() - t0
            if world > 1:
                xg[:P].copy_(torch.from_numpy(res.x), non_blocking=True)
                dist.gather(xg, gather_buf, dst=0)

        e2e_step()
        assert (res.exitflag == 1).all() and np.abs(res.x - t["xref"].cpu().numpy()).max() < 1e-6
        ksteps = args.e2e_steps or max(1, min(args.steps, 3))
        barrier()
        solve_secs[0] = 0.0
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # the host link alone: every rank copies its pinned H block to the device at the same time, nothing else running --
        # what the box's host-to-device fabric gives N concurrent ranks (the e2e step cannot beat it)
        dst = torch.empty_like(t["H"])
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            dst.copy_(h["H"], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        copy_gbs = 2 * h["H"].numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dst
        h2d = sum(v.numel() * v.element_size() for v in h.values())
        d2h = res.x.nbytes + res.lam.nbytes + res.fval.nbytes + res.exitflag.nbytes + res.iter.nbytes
        h2d_all = sum(gather_floats(float(h2d)))
        d2h_all = sum(gather_floats(float(d2h)))
        e2e = {"value": args.problems * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
               "steps": ksteps, "api": "daqp_b200_solve_packed (host C ABI, pinned buffers, chunked copy/solve overlap)",
               "h2d_gbs_per_rank": [round(g, 2) for g in gather_floats(h2d * ksteps / solve_secs[0] / 1e9)],
               "h2d_copy_only_gbs_per_rank": [round(g, 2) for g in gather_floats(copy_gbs)],
               "numa": numa}
        # ---- persistent workspace: setup once, then update(f, b) + warm solve per "MPC step".
        # Reported next to the headline, not part of it: an extra object in the same line.
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
            # ---- shared workspace (one controller, many states): G matrix sets, P / G problems per set that differ only
            # in f and the bounds; every step is a COLD solve of all P problems with f / bounds copied from pinned host memory.
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

    # ---- the reference CPU solver on a bounded sample of the SAME problems (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        S = min(P, args.cpu_sample)
        b = torch_to_batch(t, n, m, ms, 0, S)
        drv, kind = cpu_driver()
        s = drv.solve_packed(b, nthreads=1)
        x_gpu = out["x"][:S].cpu().numpy()
        assert np.array_equal(s.exitflag, out["exitflag"][:S].cpu().numpy())
        assert np.array_equal(s.iter, out["iter"][:S].cpu().numpy()), "iteration counts differ from the CPU reference"
        assert np.abs(s.x - x_gpu).max() < 1e-9 * (1 + np.abs(s.x).max())
        cpu = {"value": S / s.seconds, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {S} problems of the timed batch, daqp_quadprog per problem, 1 thread; "
                         "exit flags, iteration counts and x checked against the GPU results",
               "host_cpus": os.cpu_count()}

    # ---- the other BASELINE.json configurations and the literal drop-in entries (rank 0, N=1 only; each leg guarded)
    extra = {}
    if rank == 0 and world == 1 and not args.no_configs:
        if hn is not None:
            try:
                extra["aos"], extra["latency"] = leg_aos_latency(hn, n, m, ms, args)
            except Exception as ex:
                extra["aos"] = {"error": repr(ex)[:200]}
        del t, out, diag, hn
        torch.cuda.empty_cache()
        for name, fn in (("c4", lambda: leg_c4(eng, dev, peak, args)), ("c5", lambda: leg_c5(dev, args))):
            try:
                extra[name] = fn()
            except Exception as ex:
                extra[name] = {"error": repr(ex)[:300]}
                torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, args.problems, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": st["setup_launches"] + st["solve_launches"], "clocks": clocks,
                "per_rank_ms_per_step": [round(v, 3) for v in per_rank_ms], "weak": weak,
                "workspace": wsp, "shared_workspace": shw,
                "parity": {"max_abs_x_err_vs_constructed_optimum": err, "all_optimal": True,
                           "active_set_differs_from_construction": as_mismatch}}
        line.update(extra)
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()

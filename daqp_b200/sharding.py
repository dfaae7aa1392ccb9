"""Batch sharding across GPUs.

Problems are independent, so the data path has NO collective: every rank solves a contiguous block of the batch on
its own GPU. torch.distributed is used only as transport around the solve -- scatter of the packed problem slabs
from the rank that owns the host data, gather of the small result arrays back to it. With ``backend="nccl"`` the
buffers are device tensors (NVLink / NVSwitch); with ``backend="gloo"`` the same code runs on CPU tensors, which is
how the partitioning and the scatter / gather plumbing are tested without GPUs (tests/test_sharding.py).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def partition(N: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced blocks: the first ``N % world`` ranks get one extra problem."""
    base, extra = divmod(N, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def partition_by_cost(costs: Sequence[float], world: int) -> list[np.ndarray]:
    """Mixed-size batches: longest-processing-time-first on an estimated cost (e.g. n^2 m) so that every rank gets
    the same mix. Returns, per rank, the indices it owns (ascending)."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(costs), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return [np.nonzero(owner == r)[0] for r in range(world)]


def scatter_solve_gather(arrays: dict | None, n: int, m: int, ms: int, solve_local: Callable[[dict], dict], src: int = 0,
                         device=None) -> dict | None:
    """Rank ``src`` holds the whole batch (dict of torch tensors H, A, bupper, blower and optionally f and sense -- int32
    warm-start / equality / soft bits -- with leading dim N); every rank receives its block of EVERY array present, runs
    ``solve_local`` on it and the results (x, lam, fval, exitflag, iter, in their own dtypes) are gathered back on
    ``src``. Other ranks pass ``arrays=None`` and get ``None`` back. (n, m, ms are kept for the callers' convenience:
    the shapes travel with the tensors.)"""
    return scatter_apply_gather(arrays, solve_local, src=src, device=device)


def scatter_apply_gather(arrays: dict | None, apply_local: Callable[[dict], dict], src: int = 0, device=None) -> dict | None:
    """Generic form of the same plumbing for the other batched entry points (polyhedra for ``minrep_batch``, iterates for
    ``init_active_batch``): rank ``src`` holds a dict of torch tensors that share their leading dimension N; every rank
    receives its contiguous block of each, runs ``apply_local`` on the block and the tensors it returns (leading
    dimension = the block's length) are gathered back on ``src``. Other ranks pass ``arrays=None`` and get ``None``."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        meta = [{k: (tuple(v.shape), v.dtype) for k, v in arrays.items()}]
    dist.broadcast_object_list(meta, src=src)
    meta = meta[0]
    N = next(iter(meta.values()))[0][0]
    blocks = partition(N, world)
    lo, hi = blocks[rank]
    width = max(b - a for a, b in blocks)
    dev = device if device is not None else (next(iter(arrays.values())).device if rank == src else torch.device("cpu"))
    local = {}
    for key, (shape, dtype) in meta.items():
        buf = torch.empty((width,) + shape[1:], dtype=dtype, device=dev)
        padded = None
        if rank == src:
            padded = []
            for a, b in blocks:  # scatter needs equally sized chunks: pad the short ones, receivers trim
                c = arrays[key][a:b].contiguous().to(dev)
                padded.append(torch.cat([c, c.new_zeros((width - c.shape[0],) + shape[1:])]) if c.shape[0] < width else c)
        dist.scatter(buf, padded, src=src)
        local[key] = buf[: hi - lo].clone()
    res = apply_local(local)
    out = {} if rank == src else None
    for key in sorted(res):
        t = res[key]
        pad = t.new_zeros((width,) + tuple(t.shape[1:]))
        pad[: t.shape[0]] = t
        gathered = [torch.empty_like(pad) for _ in range(world)] if rank == src else None
        dist.gather(pad, gathered, dst=src)
        if rank == src:
            out[key] = torch.cat([g[: b - a] for g, (a, b) in zip(gathered, blocks)])
    return out

"""CPU suite: batch partitioning and the scatter -> solve -> gather plumbing with world_size 2 over gloo.
The local "solve" is the oracle here (no GPU in this container); on the GPU box the same plumbing carries the CUDA
path (tests/test_gpu_parity.py::test_sharded_two_ranks_nccl when two GPUs are present)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_balanced():
    from daqp_b200.sharding import partition
    for N in (0, 1, 7, 100_000, 100_003):
        for world in (1, 2, 4, 8):
            blocks = partition(N, world)
            assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == N
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_partition_by_cost_mixed_sizes():
    from daqp_b200.sharding import partition_by_cost
    rng = np.random.default_rng(0)
    n = rng.choice(np.arange(8, 129, 8), 5000)
    cost = n.astype(float) ** 2 * (4 * n)
    parts = partition_by_cost(cost, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(5000))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from daqp_b200.problems import QPBatch, generate_g1
    from daqp_b200.sharding import scatter_solve_gather
    from oracle import harness
    n, m, ms = 10, 24, 3
    arrays = None
    if rank == 0:
        b = generate_g1(37, n, m, ms, 7, seed=77)  # 37: uneven split
        arrays = {k: torch.from_numpy(getattr(b, k)) for k in ("H", "f", "A", "bupper", "blower")}

    def solve_local(loc):
        lb = QPBatch(n, m, ms, *(loc[k].numpy() for k in ("H", "f", "A", "bupper", "blower")),
                     np.zeros((loc["H"].shape[0], m), np.int32))
        s = harness.OracleLib().solve_packed(lb)
        return {"x": torch.from_numpy(s.x), "lam": torch.from_numpy(s.lam), "fval": torch.from_numpy(s.fval),
                "exitflag": torch.from_numpy(s.exitflag), "iter": torch.from_numpy(s.iter)}

    out = scatter_solve_gather(arrays, n, m, ms, solve_local, src=0)
    if rank == 0:
        whole = harness.OracleLib().solve_packed(b)
        ok = (np.array_equal(out["x"].numpy(), whole.x) and np.array_equal(out["iter"].numpy(), whole.iter)
              and np.array_equal(out["exitflag"].numpy(), whole.exitflag) and np.array_equal(out["lam"].numpy(), whole.lam))
        open(tmp, "w").write("ok" if ok else "mismatch")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_solve_gather_world2_gloo(oracle_libs, tmp_path):
    import torch.multiprocessing as mp
    marker = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, 29517, marker), nprocs=2, join=True)
    assert open(marker).read() == "ok"


def _worker_sense(rank, world, port, tmp):
    """sense (int32 warm-start / equality bits) is scattered with its block; f may be absent."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from daqp_b200.problems import QPBatch, generate_g1
    from daqp_b200.sharding import scatter_solve_gather
    from oracle import harness
    n, m, ms = 8, 20, 2
    arrays = None
    if rank == 0:
        b = generate_g1(23, n, m, ms, 6, seed=78)
        cold = harness.OracleLib().solve_packed(b)
        b.sense[cold.lam > 1e-12] = 1
        b.sense[cold.lam < -1e-12] = 3
        arrays = {k: torch.from_numpy(getattr(b, k)) for k in ("H", "f", "A", "bupper", "blower", "sense")}

    def solve_local(loc):
        assert loc["sense"].dtype == torch.int32
        lb = QPBatch(n, m, ms, *(loc[k].numpy() for k in ("H", "f", "A", "bupper", "blower")), loc["sense"].numpy())
        s = harness.OracleLib().solve_packed(lb, use_sense=True)
        return {"x": torch.from_numpy(s.x), "iter": torch.from_numpy(s.iter), "exitflag": torch.from_numpy(s.exitflag)}

    out = scatter_solve_gather(arrays, n, m, ms, solve_local, src=0)
    if rank == 0:
        ok = bool((out["iter"] == 1).all()) and out["iter"].dtype == torch.int32 and np.abs(out["x"].numpy() - cold.x).max() < 1e-9
        open(tmp, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_solve_gather_carries_sense_world2_gloo(oracle_libs, tmp_path):
    import torch.multiprocessing as mp
    marker = str(tmp_path / "result_sense.txt")
    mp.spawn(_worker_sense, args=(2, 29519, marker), nprocs=2, join=True)
    assert open(marker).read() == "ok"


def _worker_minrep(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from daqp_b200.problems import generate_polyhedra
    from daqp_b200.sharding import scatter_apply_gather
    from oracle import harness
    arrays = None
    if rank == 0:
        A, b = generate_polyhedra(11, 4, 18, 0, seed=5)  # 11 polyhedra: uneven split
        arrays = {"A": torch.from_numpy(A), "b": torch.from_numpy(b)}

    def minrep_local(loc):  # on the GPU box: Engine.minrep_batch_device(loc["A"], loc["b"])
        o = harness.OracleLib()
        red = np.stack([o.minrep(loc["A"][q].numpy(), loc["b"][q].numpy()) for q in range(loc["A"].shape[0])])
        return {"is_redundant": torch.from_numpy(red)}

    out = scatter_apply_gather(arrays, minrep_local, src=0)
    if rank == 0:
        o = harness.OracleLib()
        whole = np.stack([o.minrep(A[q], b[q]) for q in range(A.shape[0])])
        open(tmp, "w").write("ok" if np.array_equal(out["is_redundant"].numpy(), whole) else "mismatch")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_minrep_gather_world2_gloo(oracle_libs, tmp_path):
    """The generic scatter -> apply -> gather plumbing with the polyhedra of the minimal-representation path."""
    import torch.multiprocessing as mp
    marker = str(tmp_path / "result_minrep.txt")
    mp.spawn(_worker_minrep, args=(2, 29523, marker), nprocs=2, join=True)
    assert open(marker).read() == "ok"

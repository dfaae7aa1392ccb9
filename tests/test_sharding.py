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

"""Multi-process (world_size 2, gloo, CPU) test of the instance-sharding host logic: the
partition is a disjoint cover, every rank ends with the identical residual table, results gather to rank 0.  The
solver is a stub (no GPU here); on the GPU box the same code path runs with NCCL and the CUDA solver (bench.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_inst, out_dir):
    import torch.distributed as dist
    import itensorqtt_jl_b200 as q
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def solve_one(i):
            return {"inst": i, "owner": rank}, (0.5 ** i, 100 + i, float(rank))
        table, results = q.batch_solve(n_inst, solve_one, gather_results=True)
        np.save(os.path.join(out_dir, f"table_{rank}.npy"), table)
        np.save(os.path.join(out_dir, f"owned_{rank}.npy"), np.array(sorted(results.keys())))
    finally:
        dist.destroy_process_group()


def test_shard_instances_is_a_disjoint_cover():
    import itensorqtt_jl_b200 as q
    for n, w in ((256, 8), (5, 2), (3, 4), (0, 2), (7, 1)):
        parts = [q.shard_instances(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        q.shard_instances(4, 2, 2)


@pytest.mark.parametrize("n_inst", [5, 1])
def test_batch_solve_two_ranks_gloo(tmp_path, n_inst):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_inst, str(tmp_path)), nprocs=world, join=True)
    t0 = np.load(tmp_path / "table_0.npy")
    t1 = np.load(tmp_path / "table_1.npy")
    assert t0.shape == (n_inst, 3) and np.array_equal(t0, t1)
    for i in range(n_inst):
        assert t0[i, 0] == 0.5 ** i and t0[i, 1] == 100 + i and t0[i, 2] == i % world   # owner = i mod world
    assert list(np.load(tmp_path / "owned_0.npy")) == list(range(n_inst))                # gathered to rank 0
    assert list(np.load(tmp_path / "owned_1.npy")) == list(range(1, n_inst, 2))


def test_batch_solve_single_process():
    import itensorqtt_jl_b200 as q
    table, results = q.batch_solve(4, lambda i: (i * i, (float(i), 2.0)))
    assert table.shape == (4, 2) and list(table[:, 0]) == [0.0, 1.0, 2.0, 3.0]
    assert results == {0: 0, 1: 1, 2: 4, 3: 9}


def test_batch_solve_streams_threads():
    import itensorqtt_jl_b200 as q
    seen = []
    def solve_one(i, slot):
        seen.append((i, slot))
        return i, (float(i), float(slot))
    table, results = q.batch_solve(7, solve_one, streams=3)
    assert list(table[:, 0]) == [float(i) for i in range(7)]
    assert sorted(i for i, _ in seen) == list(range(7)) and all(slot == i % 3 for i, slot in seen)

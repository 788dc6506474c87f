"""Host-side multi-GPU logic on CPU: frame sharding arithmetic and the gather of unequal
row blocks, exercised with world_size = 2 over the gloo backend (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manifold_based_optical_flow_method_b200 import distributed as mdist


def test_shard_ranges_cover_all_frames():
    for n in (0, 1, 7, 999, 1000, 7999):
        for world in (1, 2, 4, 8):
            counts = mdist.shard_counts(n, world)
            assert sum(counts) == n and max(counts) - min(counts) <= 1
            edges = [mdist.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, width, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert mdist.world_size() == world and mdist.rank() == rank
        counts = mdist.shard_counts(n_frames, world)
        k0, k1 = mdist.shard_range(n_frames, world, rank)
        full = torch.arange(n_frames * width, dtype=torch.float64).reshape(n_frames, width)
        local = full[k0:k1].clone()
        got_all = mdist.gather_rows(local, counts)
        assert torch.equal(got_all, full)
        got_root = mdist.gather_rows(local, counts, root=0)
        if rank == 0:
            assert torch.equal(got_root, full)
        else:
            assert got_root is None
        # host delivery through shared memory: every rank writes its rows, everyone sees all of them
        assert mdist.same_host()
        shared = mdist.shared_host_rows(n_frames, width)
        assert shared.shape == (n_frames, width) and shared.dtype == np.float64
        shared[k0:k1] = full[k0:k1].numpy()
        dist.barrier()
        assert np.array_equal(shared, full.numpy())
        dist.barrier()
        empty = mdist.shared_host_rows(0, width)
        assert empty.shape == (0, width)
        del shared, empty
        # pooled result buffers: reused once the caller has dropped every view, never while one is held
        pool = mdist.SharedResultPool()
        assert pool.same_host()
        e1 = pool.acquire(n_frames, width, (k0, k1), receiver=(rank == 0), register=False)
        e1["mine"][...] = full[k0:k1].numpy()
        dist.barrier()
        V1 = pool.view(e1)
        assert np.array_equal(V1, full.numpy())
        rows = list(V1)                                   # what compute_velocity_field hands to the caller
        del V1
        e2 = pool.acquire(n_frames, width, (k0, k1), receiver=(rank == 0), register=False)
        assert e2 is not e1                               # rank 0 still holds rows of the first buffer
        del rows
        e3 = pool.acquire(n_frames, width, (k0, k1), receiver=(rank == 0), register=False)
        assert e3 is e1 or e3 is e2                       # both free now: no third buffer
        assert len(pool.entries) == 2
        pool.close()
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8, 1])
def test_gather_rows_gloo_world2(tmp_path, n_frames):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, 5, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))

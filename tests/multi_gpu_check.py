"""Multi-GPU check, run under torchrun on the GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank calls the reference-shaped compute_velocity_field with the full signal; frames are
sharded by rank, solved, and delivered (shared host memory by default, NCCL gather as the alternative).  The result must equal the oracle to the
parity tolerance and -- because every frame's arithmetic is lane-private and its reductions
are deterministic -- be bit-identical to what a single GPU computes for the same frames.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof  # noqa: E402
from manifold_based_optical_flow_method_b200 import distributed as mdist  # noqa: E402
from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import mof_oracle  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    world, rank = dist.get_world_size(), dist.get_rank()
    coords, tris, normals, areas = synthetic.icosphere(4)
    T = 71                                   # 70 frames: uneven shards for world = 4 / 8
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=3)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    V_k, _ = cof.compute_velocity_field(world, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    assert len(V_k) == T - 1 and cof.last_solve_info.converged and len(cof.last_solve_info.iterations) == T - 1
    V = np.asarray(V_k)
    # every rank holds the same gathered field
    h = torch.from_numpy(V).to(f"cuda:{local}")
    ref = h.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(h, ref)
    # the rank's own shard solved locally without torch.distributed: bit-identical
    k0, k1 = mdist.shard_range(T - 1, world, rank)
    I_dev = torch.from_numpy(I).to(f"cuda:{local}")
    V_loc, _ = cof.solve_on_device(a2, I_dev, I_dev, t_k, 0.01, 0, T - 1)
    assert torch.equal(V_loc[k0:k1], h[k0:k1]), "sharded result differs from the single-GPU result"
    assert torch.equal(V_loc, h), "sharding changed a frame's bits"
    # both host transports (shared host memory / NCCL gather + drain), to every rank and to rank 0 only
    for transport in ("shm", "nccl"):
        out, info = mdist.compute_velocity_field_sharded(a2, T - 1, t_k, 0.01, I, I, gather="all", transport=transport)
        assert np.array_equal(out, V), transport
        assert len(info.iterations) == T - 1 and info.converged
        out, info = mdist.compute_velocity_field_sharded(a2, T - 1, t_k, 0.01, I, I, gather="root", transport=transport)
        if rank == 0:
            assert np.array_equal(out, V), transport
            assert len(info.iterations) == T - 1
        else:
            assert out is None
    # equal shards (64 frames): the pipelined NCCL path
    out, _ = mdist.compute_velocity_field_sharded(a2, 64, t_k, 0.01, I, I, gather="root", transport="nccl")
    if rank == 0:
        assert np.array_equal(out, V[:64])
    # results that stay on the device
    out, _ = mdist.compute_velocity_field_sharded(a2, T - 1, t_k, 0.01, I, I, gather="all", to_host=False)
    assert torch.equal(out, h)
    # config 5: S5 wave speed with the frames sharded over the GPUs (time-derivative halo, no data-path collective
    # besides the gather of the result): bit-identical to the whole trial on one GPU, both derivative modes
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    surf = synthetic.SurfaceMesh(coords, tris, normals, areas)
    phases = synthetic.wrapped_phase(coords, t_k, seed=5, omega=300.0)
    for data, fn, phase in ((phases, s5.wave_velocity_phase, True), (I, s5.wave_velocity_amplitude, False)):
        w_sharded = fn(surf, data, 1 / 512.0, T, e)
        op5 = s5._operator(coords, tris, areas, np.asarray(e, dtype=np.float64).reshape(-1, 2, 3))
        d = torch.from_numpy(np.ascontiguousarray(data)).to(f"cuda:{local}")
        _, w_one = s5.wave_speed_device(op5, d, 0, T, 0, T, 1 / 512.0, phase)
        assert np.array_equal(w_sharded, w_one.cpu().numpy(), equal_nan=True), "sharded wave speed differs"
    if rank == 0:
        a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
        for k in (0, 17, 35, 69):
            Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
            rel = np.linalg.norm(V[k] - Vo) / np.linalg.norm(Vo)
            assert rel <= 1e-8, (k, rel)
        print(f"multi_gpu_check OK: world={world} frames={T - 1} shards={mdist.shard_counts(T - 1, world)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

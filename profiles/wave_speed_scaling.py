"""Config 5 scaling series (BASELINE.json configs[4]: S5 wave speed on a ~160k-vertex mesh over the GPUs of one box).

    python profiles/wave_speed_scaling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/wave_speed_scaling.py --gpus N                             # N GPUs

A trial of N x 1000 phase frames (weak scaling: 1000 frames per GPU) on the pial-like ico7 mesh is sharded by frame
range; every rank holds its rows plus the time-derivative halo in HBM (S5_compute_wave_v.halo_rows) and calls
mof_wave_speed on them -- no collective on the data path.  Timed with CUDA events between barriers, max over ranks;
rank 0 prints one JSON line (frames/s of the whole job; per-call and stencil-only rates).
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifold_based_optical_flow_method_b200 import _lib, synthetic  # noqa: E402
from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5  # noqa: E402
from manifold_based_optical_flow_method_b200.distributed import shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU")
    ap.add_argument("--level", type=int, default=7)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    coords, tris, normals, areas = synthetic.pial_like(args.level)
    N = len(coords)
    T = world * args.frames
    k0, k1 = shard_range(T, world, rank)
    a, b = s5.halo_rows(k0, k1, T, True)
    t_k = synthetic.time_axis(T, 512.0)
    phases = synthetic.wrapped_phase(coords, t_k[a:b], seed=1, omega=500.0)
    e = np.zeros((N, 2, 3))
    from oracle import mof_oracle
    e = mof_oracle.orthonormal_basis(normals)
    op = s5._operator(coords, tris, areas, e)
    d = torch.from_numpy(np.ascontiguousarray(phases)).to(dev)
    lib = _lib.load()
    ms5 = op.struct()
    work = torch.empty((int(lib.mof_wave_work_doubles(ctypes.byref(ms5), b - a, 0, 1)),), dtype=torch.float64, device=dev)
    out = torch.empty((k1 - k0, N), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run():
        return s5.wave_speed_device(op, d, a, T, k0 - a, k1 - k0, 1 / 512.0, True, work=work, wave_out=out)[1]

    for _ in range(3):
        w = run()
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(args.steps):
        w = run()
    e1.record()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(args.steps):
        _lib.check(lib.mof_wave_stencil(ctypes.byref(ms5), b - a, k0 - a, k1 - k0, a, T, 1 / 512.0, 1, None, out.data_ptr(), work.data_ptr(), st))
    e2.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps, e1.elapsed_time(e2) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    call_ms, sten_ms = (float(x) for x in ms.cpu())
    if rank == 0:
        print(json.dumps({
            "metric": "S5 wave-speed frames/sec @160k-vertex mesh (config 5)", "value": T / (call_ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "ms_per_step": call_ms, "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pial-like ico{args.level} mesh ({N} vertices), {args.frames} wrapped-phase frames per GPU, frames "
                                   "sharded by range with a time-derivative halo, inputs resident in HBM"},
            "rows_kernel_only": {"frames_per_s": T / (sten_ms * 1e-3), "ms": sten_ms,
                             "hbm_GBps_on_16N_bytes": 16.0 * N * T / (sten_ms * 1e-3) / 1e9 / world, "note": "per GPU"},
            "finite_fraction": float(torch.isfinite(w).double().mean())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

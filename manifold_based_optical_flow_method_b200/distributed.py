"""Frame sharding across the GPUs of one box (one process per GPU, torchrun).

Frames are independent units (worker k reads only I_k[k], I_k_2[k+1], t_k[k], t_k[k+1];
utils/compute_optical_flow.py:162-176), so rank r solves a contiguous range of frames with
no data-path collective; the only exchange is the gather of the per-frame fields that
replaces the reference's ``AsyncResult.get()`` loop (:190-191): one NCCL all-gather over
NVLink of the (frames_r, 2N) fp64 shards.  The helpers are device-agnostic so the host
logic is testable on CPU with the gloo backend (tests/test_distributed_cpu.py).
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world_size():
    dist = _dist()
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    dist = _dist()
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(n_frames, world, r):
    """Contiguous split: rank r gets frames [k0, k1); sizes differ by at most one and the
    first ``n_frames % world`` ranks take the extra frame."""
    base, extra = divmod(int(n_frames), int(world))
    k0 = r * base + min(r, extra)
    return k0, k0 + base + (1 if r < extra else 0)


def shard_counts(n_frames, world):
    return [shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world)]


def gather_rows(local, counts, group=None):
    """All-gather row blocks of unequal height: ``local`` is this rank's (counts[rank], C)
    tensor (CUDA -> NCCL over NVLink, CPU -> gloo); returns the (sum(counts), C) tensor on
    every rank.  Shards are padded to the tallest one so a single collective moves them."""
    import torch
    dist = _dist()
    world = len(counts)
    if world == 1:
        return local
    cmax = max(counts)
    tail = tuple(local.shape[1:])
    if local.shape[0] != cmax:
        pad = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    else:
        pad = local.contiguous()
    out = torch.empty((world, cmax) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view((world * cmax,) + tail), pad, group=group)
    if all(c == cmax for c in counts):
        return out.view((world * cmax,) + tail)
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def compute_velocity_field_sharded(op, n_frames, t_k, lambda_, I_k, I_k_2, to_host=True):
    """Rank-local solve of frames shard_range(...) + all-gather.  -> (V (n_frames, 2N), SolveInfo)
    V is a numpy array (to_host) or a device tensor."""
    import torch
    from . import compute_optical_flow as cof
    from .solver import SolveInfo
    world, r = world_size(), rank()
    counts = shard_counts(n_frames, world)
    k0, k1 = shard_range(n_frames, world, r)
    N = op.n_vertices
    if k1 > k0:
        # one-frame input halo: frame k1-1 reads I_k_2[k1]
        I_dev, I2_dev = cof._upload_signals(op, I_k, I_k_2, k1 - k0, first=k0)
        V_loc, info = cof.solve_on_device(op, I_dev, I2_dev, t_k[k0:k1 + 1], lambda_, 0, k1 - k0)
        rep = np.stack([info.iterations.astype(np.float64), info.relres, info.status.astype(np.float64)], axis=1)
    else:
        V_loc = torch.empty((0, 2 * N), dtype=torch.float64, device=op.device)
        rep = np.zeros((0, 3))
    V_all = gather_rows(V_loc, counts)
    rep_all = gather_rows(torch.from_numpy(rep).to(op.device), counts).cpu().numpy()
    info = SolveInfo(rep_all[:, 0].astype(np.int32), rep_all[:, 1].copy(), rep_all[:, 2].astype(np.int32))
    return (V_all.cpu().numpy() if to_host else V_all), info

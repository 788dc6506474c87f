"""Frame sharding across the GPUs of one box (one process per GPU, torchrun).

Frames are independent units (worker k reads only I_k[k], I_k_2[k+1], t_k[k], t_k[k+1];
utils/compute_optical_flow.py:162-176), so rank r solves a contiguous range of frames with
no data-path collective; the only exchange is the delivery of the per-frame fields that
replaces the reference's ``AsyncResult.get()`` loop (:190-191):

* results wanted on the device: one NCCL all-gather (or gather to rank 0) over NVLink of the
  (frames_r, 2N) fp64 shards;
* results wanted on the host (what the reference's caller gets): the same gather, batch by
  batch on a side stream, with rank 0 copying the gathered batch straight into pinned host
  memory while the next batch is solved; or (``transport="shm"``) every rank drains its own
  shard through its own PCIe link into one array in shared host memory that all ranks of the
  host map (``shared_host_rows``).

The helpers are device-agnostic so the host logic is testable on CPU with the gloo backend
(tests/test_distributed_cpu.py).
"""
import mmap
import os
import socket
import struct
import uuid

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


_numa_bound = {}


def bind_to_gpu_numa_node(device_index=None):
    """Pin this process (its threads and, through first touch, the host buffers it allocates afterwards) to the
    CPUs of the NUMA node its GPU hangs off, so that a rank's H2D / D2H traffic does not cross the socket
    interconnect.  torchrun does not place ranks; with 8 ranks moving 31 GB per step through the host it matters.
    Best effort: returns the node id, or None when the topology cannot be read (then nothing changes)."""
    try:
        import torch
        if device_index is None:
            device_index = torch.cuda.current_device()
        if device_index in _numa_bound:
            return _numa_bound[device_index]
        props = torch.cuda.get_device_properties(device_index)
        addr = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            raise OSError("no NUMA information")
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            raise OSError("node has no allowed CPU")
        os.sched_setaffinity(0, allowed)
        _numa_bound[device_index] = node
        return node
    except Exception:
        _numa_bound[device_index] = None
        return None


def shard_range(n_frames, world, r):
    """Contiguous split: rank r gets frames [k0, k1); sizes differ by at most one and the
    first ``n_frames % world`` ranks take the extra frame."""
    base, extra = divmod(int(n_frames), int(world))
    k0 = r * base + min(r, extra)
    return k0, k0 + base + (1 if r < extra else 0)


def shard_counts(n_frames, world):
    return [shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world)]


def gather_rows(local, counts, group=None, root=None):
    """Gather row blocks of unequal height: ``local`` is this rank's (counts[rank], C) tensor
    (CUDA -> NCCL over NVLink, CPU -> gloo).  root=None: all-gather, every rank gets the
    (sum(counts), C) tensor; root=r: only rank r gets it (others get None).  Shards are
    padded to the tallest one so a single collective moves them."""
    import torch
    dist = _dist()
    world = len(counts)
    if world == 1:
        return local
    me = dist.get_rank(group)
    cmax = max(counts)
    tail = tuple(local.shape[1:])
    if local.shape[0] != cmax:
        pad = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    else:
        pad = local.contiguous()
    if root is None:
        out = torch.empty((world, cmax) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view((world * cmax,) + tail), pad, group=group)
    else:
        out = torch.empty((world, cmax) + tail, dtype=local.dtype, device=local.device) if me == root else None
        dist.gather(pad, list(out.unbind(0)) if me == root else None, dst=root, group=group)
        if me != root:
            return None
    if all(c == cmax for c in counts):
        return out.view((world * cmax,) + tail)
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def same_host(group=None):
    """True when every rank of the group runs on this host (collective call)."""
    dist = _dist()
    world = dist.get_world_size(group)
    names = [None] * world
    dist.all_gather_object(names, (socket.gethostname(), os.stat("/proc/self/ns/ipc").st_ino), group=group)
    return all(n == names[0] for n in names)


def shared_host_rows(rows, width, group=None):
    """(rows, width) float64 array in anonymous shared memory (a memfd created by rank 0 and
    handed to the other ranks over an abstract-namespace unix socket), mapped by every rank of
    the group; all ranks must be on one host and call this together.  The memory goes away with
    the last array that refers to it -- nothing is left in /dev/shm, whatever way a rank dies."""
    dist = _dist()
    world, me = dist.get_world_size(group), dist.get_rank(group)
    nbytes = max(int(rows) * int(width) * 8, mmap.PAGESIZE)
    token = [None]
    if me == 0:
        fd = os.memfd_create("mof_rows")
        os.ftruncate(fd, nbytes)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        token[0] = f"\0mof-rows-{os.getpid()}-{uuid.uuid4().hex}"
        srv.bind(token[0])
        srv.listen(world)
    dist.broadcast_object_list(token, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if me == 0:
        served = 0
        while served < world - 1:
            conn, _ = srv.accept()
            # abstract sockets have no file permissions: only hand the memory to processes of this user
            cred = conn.getsockopt(socket.SOL_SOCKET, socket.SO_PEERCRED, struct.calcsize("3i"))
            if struct.unpack("3i", cred)[1] == os.getuid():
                socket.send_fds(conn, [b"m"], [fd])
                served += 1
            conn.close()
        srv.close()
    else:
        conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        conn.connect(token[0])
        _, fds, _, _ = socket.recv_fds(conn, 1, 1)
        conn.close()
        fd = fds[0]
    mm = mmap.mmap(fd, nbytes)                       # MAP_SHARED; the mapping outlives the descriptor
    os.close(fd)
    return np.frombuffer(mm, dtype=np.float64, count=int(rows) * int(width)).reshape(int(rows), int(width))


class SharedResultPool:
    """Result buffers of the multi-GPU host delivery: (rows, width) float64 arrays in shared memory that every
    rank of the host maps (shared_host_rows), kept across calls.  When a buffer is created every rank touches
    its own rows (the page faults of fresh shared memory happen once, in parallel, outside any later timed
    path) and registers them with CUDA (cudaHostRegister), so that from then on a rank's drain is one direct
    DMA per batch over its own PCIe link.  A buffer is handed out again only after the caller has let go of
    every array that refers to it (reference count of the mapping's root array, decided by the receiving
    rank and broadcast), so results a caller still holds are never overwritten."""

    def __init__(self):
        self.entries = []
        self._same_host = {}

    def same_host(self, group=None):
        key = id(group)
        if key not in self._same_host:
            self._same_host[key] = same_host(group)
        return self._same_host[key]

    @staticmethod
    def _in_use(entry):
        import sys
        return sys.getrefcount(entry["root"]) > entry["baseline"]

    def acquire(self, rows, width, my_rows, receiver, register=True, group=None):
        """Collective.  -> entry dict: ``array`` (rows, width) view of the whole buffer (make a fresh view per
        hand-out with ``view()``), ``mine`` numpy view of rows my_rows[0]:my_rows[1], ``mine_t`` torch CPU tensor
        over the same memory (pinned when ``pinned`` is True)."""
        import sys
        import torch
        dist = _dist()
        rows, width = int(rows), int(width)
        pick = [-1]
        if receiver:
            for i, e in enumerate(self.entries):
                if e["shape"] == (rows, width) and e["my_rows"] == tuple(my_rows) and not self._in_use(e):
                    pick[0] = i
                    break
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(pick, src=src, group=group)
        if pick[0] >= 0:
            return self.entries[pick[0]]
        # receivers drop buffers of other shapes that nobody holds any more (a new mesh, a new frame count)
        for e in [e for e in self.entries if e["shape"] != (rows, width) and not self._in_use(e)]:
            self._release(e)
        arr = shared_host_rows(rows, width, group=group)
        root = arr.base if isinstance(arr.base, np.ndarray) else arr
        a, b = int(my_rows[0]), int(my_rows[1])
        mine = arr[a:b]
        mine[...] = 0.0                                   # first touch of this rank's pages
        mine_t = torch.from_numpy(mine)
        pinned = False
        if register and b > a and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(mine.ctypes.data, mine.nbytes, 0)
            pinned = int(rc) == 0
        entry = {"shape": (rows, width), "my_rows": (a, b), "root": root, "mine": mine, "mine_t": mine_t,
                 "pinned": pinned, "ptr": mine.ctypes.data}
        del arr, root, mine
        entry["baseline"] = sys.getrefcount(entry["root"])   # the pool's own references; more = a caller holds a view
        self.entries.append(entry)
        dist.barrier(group=group)                        # every rank has faulted and registered its rows
        return entry

    @staticmethod
    def view(entry):
        """A fresh (rows, width) array over the buffer for the caller; while it (or anything derived from it)
        lives, the buffer is not reused."""
        return entry["root"].reshape(entry["shape"])

    def _release(self, entry):
        import torch
        if entry["pinned"]:
            torch.cuda.cudart().cudaHostUnregister(entry["ptr"])
            entry["pinned"] = False
        self.entries.remove(entry)

    def close(self):
        for e in list(self.entries):
            self._release(e)


_result_pool = SharedResultPool()


def solve_shard_and_gather(op, I_shard, I2_shard, t_k_shard, lambda_, counts, gather="all", to_host=True, transport="auto"):
    """Multi-GPU entry point: this rank solves the frames of its shard -- ``I_shard`` /
    ``I2_shard`` hold rows k0 .. k1 (one-frame halo: frame k1-1 reads I2[k1]) and
    ``t_k_shard`` the matching k1-k0+1 time stamps -- then the per-frame fields are
    delivered.  gather: "all" (every rank gets all frames), "root" (rank 0 only; other ranks
    get None) or "none" (each rank keeps its shard).  -> (V, SolveInfo); V is a numpy array
    if to_host else a device tensor.  transport (host delivery with gather "all" / "root"):
    "shm" -- no collective on the data path: every rank drains its own shard over its own PCIe link, one
    direct DMA per batch, into its rows of a pooled shared host array that all ranks of the host map and
    that was faulted in and registered with CUDA when it was created (SharedResultPool); with "all" every
    rank returns the SAME memory; needs all ranks on one host;
    "nccl" -- every finished batch is gathered over NVLink on a side stream and drained by rank 0 (straight
    into pinned memory when the block can be pinned) or by every rank;
    "auto" -- "shm" when all ranks share the host, else "nccl".  (Round 1 funnelled everything through rank 0's
    link: 0.79 end-to-end efficiency at 8 GPUs; fresh unregistered shared memory had lost to it at 2 GPUs only
    because of first-touch faults inside the timed path.)"""
    import torch
    from . import compute_optical_flow as cof
    from .solver import SolveInfo
    r = rank()
    n_loc = counts[r] if len(counts) > 1 else len(t_k_shard) - 1
    N = op.n_vertices
    world = len(counts)
    width = 2 * N
    if transport not in ("auto", "shm", "nccl"):
        raise ValueError("transport must be 'auto', 'shm' or 'nccl'")
    if to_host and world > 1 and cof.settings.get("numa_bind", True):
        bind_to_gpu_numa_node(op.device.index)
    if to_host and world > 1 and gather in ("all", "root") and transport in ("shm", "auto"):
        if _result_pool.same_host():
            return _solve_shard_to_shared_host(op, I_shard, I2_shard, t_k_shard, lambda_, counts, gather, r, n_loc, width)
        if transport == "shm":
            raise RuntimeError("transport='shm' needs every rank on the same host")
    # Host delivery over NCCL with equal shards: gather every finished batch over NCCL on a side stream and
    # drain it to the host while the next batch is being solved.
    pipelined = to_host and world > 1 and gather in ("all", "root") and len(set(counts)) == 1 and n_loc > 0
    V_host = None
    if pipelined:
        dist = _dist()
        root = 0 if gather == "root" else None
        receives = root is None or r == root
        solver = cof._solver(op)
        # rank 0 alone receiving: its result goes to pinned memory by direct DMA (with "all" every
        # rank would pin world x the data, so that case keeps pageable memory and the staged drain)
        from .solver import try_pinned_rows
        block = try_pinned_rows(solver.torch, sum(counts), width) \
            if receives and gather == "root" and cof.settings["pinned_results"] else None
        pinned = block is not None
        if pinned:
            V_pin, V_host = block
        elif receives:
            V_host = np.empty((sum(counts), width), dtype=np.float64)
        drain = solver.drain(width)
        if receives and gather == "root" and cof.settings["pinned_results"] and not pinned:
            import warnings
            warnings.warn(f"result of {sum(counts) * width * 8 / 2**30:.1f} GiB was not pinned (above the cap or refused by "
                          "the host): delivering into pageable memory through the staged drain")

        def on_batch(k0, k1, Vd):
            def collect():
                if root is None:
                    out = torch.empty((world, k1 - k0, width), dtype=Vd.dtype, device=Vd.device)
                    dist.all_gather_into_tensor(out.view(world * (k1 - k0), width), Vd.contiguous())
                    return out
                out = torch.empty((world, k1 - k0, width), dtype=Vd.dtype, device=Vd.device) if receives else None
                dist.gather(Vd.contiguous(), list(out.unbind(0)) if receives else None, dst=root)
                return out
            out = drain.on_side_stream(collect)
            if pinned:
                for q in range(world):
                    drain.submit_pinned(out[q], V_pin[q * n_loc + k0:q * n_loc + k1])
            elif receives:
                for q in range(world):
                    drain.submit(out[q], V_host[q * n_loc + k0:q * n_loc + k1])
    else:
        on_batch = None

    if n_loc > 0:
        I_dev, I2_dev, upload = cof._upload_for_solve(op, I_shard, I2_shard, n_loc)
        V_loc, info = cof.solve_on_device(op, I_dev, I2_dev, list(t_k_shard), lambda_, 0, n_loc, on_batch=on_batch, upload=upload)
        rep = np.stack([info.iterations.astype(np.float64), info.relres, info.status.astype(np.float64)], axis=1)
    else:
        V_loc = torch.empty((0, width), dtype=torch.float64, device=op.device)
        rep = np.zeros((0, 3))
    rep_dev = torch.from_numpy(rep).to(op.device)
    if pipelined:
        drain.finish()
        torch.cuda.current_stream(op.device).wait_stream(drain.stream)
        rep_all = gather_rows(rep_dev, counts, root=0 if gather == "root" else None)
        if rep_all is None:
            return None, SolveInfo(info.iterations, info.relres, info.status)
        rep_np = rep_all.cpu().numpy()
        return V_host, SolveInfo(rep_np[:, 0].astype(np.int32), rep_np[:, 1].copy(), rep_np[:, 2].astype(np.int32))
    if gather == "none" or world == 1:
        V_all, rep_all = V_loc, rep_dev
    else:
        root = 0 if gather == "root" else None
        V_all = gather_rows(V_loc, counts, root=root)
        rep_all = gather_rows(rep_dev, counts, root=root)
    if V_all is None:
        return None, (SolveInfo(info.iterations, info.relres, info.status) if n_loc > 0 else None)
    rep_np = rep_all.cpu().numpy()
    info = SolveInfo(rep_np[:, 0].astype(np.int32), rep_np[:, 1].copy(), rep_np[:, 2].astype(np.int32))
    return (V_all.cpu().numpy() if to_host else V_all), info


def _solve_shard_to_shared_host(op, I_shard, I2_shard, t_k_shard, lambda_, counts, gather, r, n_loc, width):
    """Host delivery through pooled shared memory: each rank's finished batches go straight into its own
    rows of the shared array (one DMA per batch on a side stream, over the rank's own PCIe link) while the
    next batch is being solved; the per-frame solve report is the only thing that crosses NCCL."""
    import torch
    from . import compute_optical_flow as cof
    from .solver import SolveInfo
    start = int(np.sum(counts[:r]))
    receiver = gather == "all" or r == 0
    entry = _result_pool.acquire(sum(counts), width, (start, start + n_loc), receiver)
    if n_loc > 0:
        solver = cof._solver(op)
        drain = solver.drain(width)
        I_dev, I2_dev, upload = cof._upload_for_solve(op, I_shard, I2_shard, n_loc)
        if entry["pinned"]:
            mine_t = entry["mine_t"]
            on_batch = lambda k0, k1, Vd: drain.submit_pinned(Vd, mine_t[k0:k1])
        else:                                            # registration refused: staged copies into the same rows
            mine = entry["mine"]
            on_batch = lambda k0, k1, Vd: drain.submit(Vd, mine[k0:k1])
        _, info = cof.solve_on_device(op, I_dev, I2_dev, list(t_k_shard), lambda_, 0, n_loc, on_batch=on_batch, upload=upload)
        drain.finish()
        rep = np.stack([info.iterations.astype(np.float64), info.relres, info.status.astype(np.float64)], axis=1)
    else:
        info, rep = None, np.zeros((0, 3))
    # the report gather also orders every rank's writes before the receivers' return
    rep_all = gather_rows(torch.from_numpy(rep).to(op.device), counts, root=0 if gather == "root" else None)
    if rep_all is None:
        return None, (SolveInfo(info.iterations, info.relres, info.status) if info is not None else None)
    rep_np = rep_all.cpu().numpy()
    return _result_pool.view(entry), SolveInfo(rep_np[:, 0].astype(np.int32), rep_np[:, 1].copy(), rep_np[:, 2].astype(np.int32))


def compute_velocity_field_sharded(op, n_frames, t_k, lambda_, I_k, I_k_2, gather="all", to_host=True, transport="auto"):
    """compute_velocity_field under torchrun: every rank holds the full (T,N) signal like
    the reference's processes do, solves frames shard_range(...) and delivers all of them.
    -> (V (n_frames, 2N), SolveInfo)"""
    world, r = world_size(), rank()
    counts = shard_counts(n_frames, world)
    k0, k1 = shard_range(n_frames, world, r)
    same = I_k_2 is I_k
    I_sh = I_k[k0:k1 + 1]
    I2_sh = I_sh if same else I_k_2[k0:k1 + 1]
    return solve_shard_and_gather(op, I_sh, I2_sh, t_k[k0:k1 + 1], lambda_, counts, gather=gather, to_host=to_host,
                                  transport=transport)

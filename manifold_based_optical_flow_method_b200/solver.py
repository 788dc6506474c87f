"""Frame-batch driver: pack -> assemble (K1) -> batched PCG (K2/K3: SSOR on the level-scheduled or the
block-multicolour ordering, or block Jacobi) -> unpack, with the device->host drain of finished batches overlapped with the next solve.

Host orchestration only (PyTorch provides device memory and streams); all arithmetic is in
libmof_b200.so.  Replaces the multiprocessing fan-out of compute_velocity_field
(utils/compute_optical_flow.py:152-194) and the per-frame worker (:100-149).
"""
import ctypes
import dataclasses

import numpy as np

from . import _lib
from ._lib import GROUP

DEFAULT_TOL = 1e-12          # ||b - A x|| / ||b||, north-star parity setting
DEFAULT_MAX_ITER = 20000
DEFAULT_CHECK_EVERY = 32
DEFAULT_MAX_RESTARTS = 3
DEFAULT_PRECOND = "ssor_level"   # "ssor_level": SSOR in Eisenstat's form on the level-scheduled natural (Cuthill-McKee)
                             # ordering; "ssor": the same on the block-multicolour ordering; "jacobi": 2x2 block Jacobi
DEFAULT_OMEGA = 1.4          # SSOR relaxation factor of the block-multicolour ordering (372 iterations at ico7)
DEFAULT_OMEGA_LEVEL = 1.9    # ... of the level-scheduled natural ordering (B200, ico7: 143 iterations; 148 at 1.85, 192 at 1.7)
SSOR_KINDS = ("ssor", "ssor_level")
REORDER_OF = {"jacobi": 1, "ssor": 2, "ssor_level": 3}
DEFAULT_BATCH_GROUPS = 32    # 32 x 32 = 1024 frames per launch (~66 GB at 164k vertices)
DRAIN_STAGE_ROWS = 256       # rows per pinned staging buffer of the device->host pipeline
DEFAULT_STREAMS = 1          # concurrent solve streams (2 fills launch tails: +2 % measured, but blurs per-kernel timing)


@dataclasses.dataclass
class SolveInfo:
    """Per-frame solver report (the reference's spsolve returns nothing comparable)."""
    iterations: np.ndarray      # (n_frames,) int32
    relres: np.ndarray          # (n_frames,) true ||b - A x|| / ||b||
    status: np.ndarray          # (n_frames,) _lib.STATUS_*
    seconds: float = 0.0
    path: tuple = ()            # mof_pcg_last_path of the last batch: (path, grid, CTAs per SM, fallback reason)

    @property
    def converged(self):
        return bool(np.all((self.status == _lib.STATUS_CONVERGED) | (self.status == _lib.STATUS_ZERO_RHS)))


class UnconvergedError(RuntimeError):
    def __init__(self, info, frames):
        self.info, self.frames = info, frames
        super().__init__(
            f"{len(frames)} frame(s) did not converge (first: frame {frames[0]}, status {int(info.status[frames[0]])}, "
            f"relres {info.relres[frames[0]]:.3e}, {int(info.iterations[frames[0]])} iterations)")


class FrameBatch:
    """Device buffers of one batch of ``n_groups`` x 32 frames (mof_batch_dev)."""

    def __init__(self, op, n_groups, with_t=True):
        torch = _lib.require_cuda()
        lib = _lib.load()
        self.op, self.n_groups = op, int(n_groups)
        N, nb, G, W = op.n_vertices, op.n_blocks, self.n_groups, GROUP
        dev = op.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.n_tiles = int(lib.mof_num_tiles(N))
        self.It = torch.empty((G, N, W), **f64)
        self.dIt = torch.empty((G, N, W), **f64)
        self.vals = torch.empty((G, nb, 4, W), **f64)
        self.rhs = torch.empty((G, N, 2, W), **f64)
        self.minv = torch.empty((G, N, 3, W), **f64)
        self.x = torch.empty((G, N, 2, W), **f64)
        self.r = torch.empty((G, N, 2, W), **f64)
        self.z = torch.empty((G, N, 2, W), **f64)
        self.p = torch.empty((G, N, 2, W), **f64)
        self.ap = torch.empty((G, N, 2, W), **f64)
        self.t = torch.empty((G, N, 2, W), **f64) if with_t else None
        self.partial = torch.zeros((G, self.n_tiles, 2, W), **f64)
        self.scal = torch.zeros((G, _lib.SCAL_SLOTS, W), **f64)
        self.state = torch.zeros((int(lib.mof_state_ints(G)),), dtype=torch.int32, device=dev)
        # per-(group, row) sweep stamps of the persistent level-scheduled kernel
        self.ready = torch.zeros((G, N), dtype=torch.int32, device=dev) if (with_t and op.pattern.n_levels > 0) else None
        self.n_frames = 0

    def struct(self, n_frames=None, n_groups=None):
        """Descriptor for the first ``n_groups`` groups (every buffer is group-major, so a
        prefix of the allocation is a valid smaller batch)."""
        if n_frames is not None:
            self.n_frames = int(n_frames)
        G = self.n_groups if n_groups is None else int(n_groups)
        assert 0 < G <= self.n_groups and self.n_frames <= G * GROUP
        return _lib.BatchDev(
            G, self.n_frames, self.It.data_ptr(), self.dIt.data_ptr(), self.vals.data_ptr(),
            self.rhs.data_ptr(), self.minv.data_ptr(), self.x.data_ptr(), self.r.data_ptr(), self.z.data_ptr(),
            self.p.data_ptr(), self.ap.data_ptr(), self.t.data_ptr() if self.t is not None else None,
            self.partial.data_ptr(), self.scal.data_ptr(), self.state.data_ptr(),
            self.ready.data_ptr() if self.ready is not None else None)

    @staticmethod
    def bytes_per_group(n_vertices, n_blocks):
        return 8 * GROUP * (4 * n_blocks + n_vertices * (2 + 2 * 7 + 3)) + 4 * n_vertices


def upload_chunks(n_rows, step):
    """Row ranges [r0, r1) in which a signal of ``n_rows`` rows is copied so that the frames [c*step, (c+1)*step) of
    chunk c are complete when copy c is (frame k reads rows k and k+1): the first copy carries one row more."""
    out, r0 = [], 0
    while r0 < n_rows:
        r1 = min(int(n_rows), step + 1 if r0 == 0 else r0 + step)
        out.append((r0, r1))
        r0 = r1
    return out


class SignalUpload:
    """A (rows, N) device tensor that is still being filled from host memory: chunks of rows are copied on a side
    stream, in order, and ``arrivals`` lists (rows_complete, event) pairs -- rows [0, rows_complete) are on the device
    once the event has fired.  VelocitySolver.solve_frames packs and assembles every 32-frame group chunk as soon as
    its rows are there, so the host-to-device copy of a signal hides behind K1 instead of preceding it."""

    CHUNK_GROUPS = 4                     # 128 frames (~170 MB at 164k vertices) per copy

    def __init__(self, torch, device, src, n_rows, n_cols):
        self.tensor = torch.empty((int(n_rows), int(n_cols)), dtype=torch.float64, device=device)
        self.src = src                       # the host rows must outlive the copies in flight
        self.arrivals = []
        self.stream = torch.cuda.Stream(device=device)
        self.stream.wait_stream(torch.cuda.current_stream(device))
        step = self.CHUNK_GROUPS * GROUP
        with torch.cuda.stream(self.stream):
            for r0, r1 in upload_chunks(n_rows, step):
                self.tensor[r0:r1].copy_(src[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.arrivals.append((r1, ev))
        self.tensor.record_stream(self.stream)

    def wait_rows(self, torch, device, rows):
        """Make the current stream wait until rows [0, rows) have arrived."""
        for have, ev in self.arrivals:
            if have >= rows:
                torch.cuda.current_stream(device).wait_event(ev)
                return
        if self.arrivals:
            torch.cuda.current_stream(device).wait_event(self.arrivals[-1][1])


class VelocitySolver:
    """Solves batches of frames on one GPU.  Buffers are allocated once and reused."""

    def __init__(self, op, batch_groups=None, tol=DEFAULT_TOL, max_iter=DEFAULT_MAX_ITER,
                 check_every=DEFAULT_CHECK_EVERY, max_restarts=DEFAULT_MAX_RESTARTS, precond=None, omega=None,
                 n_streams=DEFAULT_STREAMS):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.op = op
        if precond is None:
            precond = "ssor_level" if op.pattern.n_levels > 0 else ("ssor" if op.pattern.n_colors > 0 else "jacobi")
        if precond not in SSOR_KINDS + ("jacobi",):
            raise ValueError(f"precond must be 'ssor', 'ssor_level' or 'jacobi', got {precond!r}")
        if precond == "ssor" and op.pattern.n_colors == 0:
            raise ValueError("the SSOR preconditioner needs a mesh built with the block-multicolour ordering (reorder=2)")
        if precond == "ssor_level" and op.pattern.n_levels == 0:
            raise ValueError("the level-scheduled SSOR preconditioner needs a mesh built with reorder=3")
        self.precond = precond
        self.ssor = precond in SSOR_KINDS
        if omega is None:
            omega = DEFAULT_OMEGA_LEVEL if precond == "ssor_level" else DEFAULT_OMEGA
        self.omega = float(omega) if self.ssor else 0.0
        self.tol, self.max_iter, self.check_every, self.max_restarts = tol, max_iter, check_every, max_restarts
        if batch_groups is None:
            free, _total = self.torch.cuda.mem_get_info(op.device)
            per_group = FrameBatch.bytes_per_group(op.n_vertices, op.n_blocks)
            batch_groups = max(1, min(DEFAULT_BATCH_GROUPS, int(0.6 * free) // max(per_group, 1)))
        self.batch_groups = int(batch_groups)
        self.n_streams = max(1, int(n_streams))
        self._batch = None
        self._lanes = None           # per-stream (FrameBatch, torch Stream, PcgProfile) of the concurrent path
        self._pool = None
        self._drain = None
        self._drain_finalizer = None
        self.profile = None          # set to _lib.PcgProfile() to accumulate sampled kernel timings
        self.aux_launches = 0        # pack / assemble / unpack launches issued so far

    def batch(self, n_groups):
        if self._batch is None or self._batch.n_groups < n_groups:
            self._batch = None
            self._batch = FrameBatch(self.op, n_groups, with_t=self.ssor)
        return self._batch

    def _solve_concurrent(self, ranges, lanes, groups_per_batch, I_dev, I2_dev, dt_dev, lambda_, V_dev, on_batch):
        """Batches are dealt round-robin to ``lanes`` host threads, each with its own CUDA stream and
        FrameBatch.  Every PCG iteration is a chain of short dependent launches (one per patch
        colour); with two independent chains in flight the SMs idling in one chain's launch tail run
        the other chain's CTAs, and one chain's convergence poll never drains the GPU."""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        torch, op = self.torch, self.op
        if self._lanes is None or len(self._lanes) < lanes or self._lanes[0][0].n_groups < groups_per_batch:
            self._lanes = None
            self._lanes = [(FrameBatch(op, groups_per_batch, with_t=self.ssor),
                            torch.cuda.Stream(device=op.device)) for _ in range(lanes)]
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.n_streams)
        main = torch.cuda.current_stream(op.device)
        infos = [None] * len(ranges)
        turn = threading.Condition()
        next_q = [0]                 # callbacks run in batch order on every rank (they may issue collectives)
        profiles = [_lib.PcgProfile() if self.profile is not None else None for _ in range(lanes)]

        failed = [False]

        def work(lane):
            batch, stream = self._lanes[lane]
            try:
                with torch.cuda.device(op.device), torch.cuda.stream(stream):
                    stream.wait_stream(main)
                    for q in range(lane, len(ranges), lanes):
                        k0, k1 = ranges[q]
                        infos[q] = self.solve_batch(I_dev[k0:k1], I2_dev[k0 + 1:k1 + 1], dt_dev[k0:k1], lambda_,
                                                    V_dev[k0:k1], batch=batch, profile=profiles[lane])
                        with turn:
                            turn.wait_for(lambda: next_q[0] == q or failed[0])
                            if failed[0]:
                                return stream
                            if on_batch is not None:
                                on_batch(k0, k1, V_dev[k0:k1])
                            next_q[0] = q + 1
                            turn.notify_all()
            except BaseException:
                with turn:
                    failed[0] = True
                    turn.notify_all()
                raise
            return stream

        futures = [self._pool.submit(work, lane) for lane in range(lanes)]
        errors = []
        for f in futures:
            try:
                main.wait_stream(f.result())
            except Exception as exc:            # every worker is joined before re-raising
                errors.append(exc)
        if errors:
            raise errors[0]
        if self.profile is not None:
            for p in profiles:
                _lib.add_profile(self.profile, p)
        return infos

    def drain(self, width, rows=None):
        """Cached HostDrain with pinned staging buffers of ``rows`` x width doubles."""
        rows = int(rows or DRAIN_STAGE_ROWS)
        if self._drain is None or not self._drain.fits(rows, width):
            if self._drain is not None:
                self._drain.close()
            self._drain = HostDrain(self.torch, self.op.device, rows, width)
            if self._drain_finalizer is not None:
                self._drain_finalizer.detach()
            import weakref
            self._drain_finalizer = weakref.finalize(self, HostDrain.close, self._drain)
        return self._drain

    def close(self):
        """Release the host-side helpers (drain thread, pinned staging, stream pool)."""
        if self._drain is not None:
            self._drain.close()
            self._drain = None
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def assemble(self, batch, I_now, I_next, dt, lambda_, n_frames):
        """pack + K1 for the first n_frames rows of I_now / I_next (device, (>=n_frames, N))."""
        torch, lib, op = self.torch, self.lib, self.op
        st = torch.cuda.current_stream(op.device).cuda_stream
        ms, bs = op.struct(), batch.struct(n_frames, (int(n_frames) + GROUP - 1) // GROUP)
        assert I_now.stride(1) == 1 and I_next.stride(1) == 1
        assert I_now.stride(0) == I_next.stride(0)
        _lib.check(lib.mof_pack_frames(ctypes.byref(ms), ctypes.byref(bs), I_now.data_ptr(), I_next.data_ptr(),
                                       I_now.stride(0), dt.data_ptr(), st))
        _lib.check(lib.mof_assemble_batch(ctypes.byref(ms), ctypes.byref(bs), float(lambda_), self.omega, st))
        return ms, bs

    def assemble_streamed(self, batch, upload, row0, dt, lambda_, n_frames):
        """pack + K1 chunk by chunk while ``upload`` (a SignalUpload whose row ``row0`` is frame 0 of this batch, the same
        tensor serving as I_now and, one row later, as I_next) is still arriving.  Same kernels on the same data as
        ``assemble``: every chunk is a view of the batch (the buffers are group-major) whose launches wait for its rows."""
        torch, lib, op = self.torch, self.lib, self.op
        st = torch.cuda.current_stream(op.device).cuda_stream
        ms = op.struct()
        I = upload.tensor
        N, nb, W = op.n_vertices, op.n_blocks, GROUP
        step = SignalUpload.CHUNK_GROUPS * GROUP
        for f0 in range(0, n_frames, step):
            f1 = min(n_frames, f0 + step)
            upload.wait_rows(torch, op.device, row0 + f1 + 1)
            g0, gc = f0 // GROUP, -(-(f1 - f0) // GROUP)
            b = batch
            view = _lib.BatchDev(
                gc, f1 - f0, b.It.data_ptr() + 8 * g0 * N * W, b.dIt.data_ptr() + 8 * g0 * N * W,
                b.vals.data_ptr() + 8 * g0 * nb * 4 * W, b.rhs.data_ptr() + 8 * g0 * N * 2 * W,
                b.minv.data_ptr() + 8 * g0 * N * 3 * W, None, None, None, None, None, None, None, None, None, None)
            _lib.check(lib.mof_pack_frames(ctypes.byref(ms), ctypes.byref(view), I[row0 + f0].data_ptr(), I[row0 + f0 + 1].data_ptr(),
                                           I.stride(0), dt[f0:].data_ptr(), st))
            _lib.check(lib.mof_assemble_batch(ctypes.byref(ms), ctypes.byref(view), float(lambda_), self.omega, st))
        return ms, batch.struct(n_frames, (int(n_frames) + GROUP - 1) // GROUP)

    def solve_batch(self, I_now, I_next, dt, lambda_, V_out, batch=None, profile=None, upload=None, upload_row0=0):
        """One batch: frames = rows of I_now.  V_out: device (n_frames, 2N) view.  -> SolveInfo.
        upload: a SignalUpload still in flight whose tensor I_now / I_next are views of (then K1 runs chunk by chunk)."""
        torch, lib, op = self.torch, self.lib, self.op
        n_frames = int(I_now.shape[0])
        G = (n_frames + GROUP - 1) // GROUP
        if batch is None:
            batch = self.batch(G)
        if profile is None:
            profile = self.profile
        st = torch.cuda.current_stream(op.device).cuda_stream
        if upload is not None:
            ms, bs = self.assemble_streamed(batch, upload, upload_row0, dt, lambda_, n_frames)
        else:
            ms, bs = self.assemble(batch, I_now, I_next, dt, lambda_, n_frames)
        iters = np.zeros(G * GROUP, np.int32)
        relres = np.zeros(G * GROUP, np.float64)
        status = np.zeros(G * GROUP, np.int32)
        _lib.check(lib.mof_pcg_solve_batch(ctypes.byref(ms), ctypes.byref(bs), float(self.tol), self.omega, int(self.max_iter),
                                           int(self.check_every), int(self.max_restarts), iters.ctypes.data,
                                           relres.ctypes.data, status.ctypes.data,
                                           ctypes.byref(profile) if profile is not None else None, st),
                   allow_positive=True)
        path = (ctypes.c_int32 * 4)()
        lib.mof_pcg_last_path(path)
        self.aux_launches += 4                    # pack, assemble (diagonal blocks, off-diagonal blocks), unpack
        assert V_out.stride(1) == 1
        _lib.check(lib.mof_unpack_solution(ctypes.byref(ms), ctypes.byref(bs), V_out.data_ptr(), V_out.stride(0), st))
        return SolveInfo(iters[:n_frames], relres[:n_frames], status[:n_frames], path=tuple(path))

    def solve_frames(self, I_dev, I2_dev, dt_dev, lambda_, V_dev=None, on_batch=None, upload=None):
        """All frames k = 0 .. n-1 with (I_dev[k], I2_dev[k+1]) (compute_optical_flow.py:174-175).
        I_dev, I2_dev: device (>= n+1, N) float64 (may be the same tensor); dt_dev: device (n,).
        on_batch(k0, k1, V_dev[k0:k1]) is called after each batch has been queued on the stream
        (used to drain results to the host while the next batch is being solved).
        -> (V_dev (n, 2N) device, SolveInfo)"""
        torch, op = self.torch, self.op
        n = int(dt_dev.shape[0])
        if V_dev is None:
            V_dev = torch.empty((n, 2 * op.n_vertices), dtype=torch.float64, device=op.device)
        groups = -(-n // GROUP)
        lanes = min(self.n_streams, max(1, groups // 2))          # concurrency only pays with >= 2 groups per stream
        per = max(1, min(self.batch_groups // lanes, -(-groups // lanes))) if lanes > 1 else self.batch_groups
        step = per * GROUP
        ranges = [(k0, min(n, k0 + step)) for k0 in range(0, n, step)]
        if upload is not None and not (I_dev is I2_dev and (lanes == 1 or len(ranges) == 1)):
            upload.wait_rows(torch, op.device, n + 1)              # concurrent lanes / two signals: wait for the whole copy
            upload = None
        if lanes == 1 or len(ranges) == 1:
            infos = []
            for k0, k1 in ranges:
                infos.append(self.solve_batch(I_dev[k0:k1], I2_dev[k0 + 1:k1 + 1], dt_dev[k0:k1], lambda_, V_dev[k0:k1],
                                              upload=upload, upload_row0=k0))
                if on_batch is not None:
                    on_batch(k0, k1, V_dev[k0:k1])
        else:
            infos = self._solve_concurrent(ranges, lanes, step // GROUP, I_dev, I2_dev, dt_dev, lambda_, V_dev, on_batch)
        if infos:
            info = SolveInfo(np.concatenate([i.iterations for i in infos]), np.concatenate([i.relres for i in infos]),
                             np.concatenate([i.status for i in infos]), path=infos[-1].path)
        else:
            info = SolveInfo(np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32))
        return V_dev, info


def pinned_rows(torch, rows, width):
    """(rows, width) float64 result buffer in pinned host memory -> (tensor, numpy view).  The
    block comes from torch's caching host allocator: once the caller lets go of the previous
    result (the numpy view keeps the tensor alive) the next call reuses it, so in steady state
    there is neither a cudaHostAlloc nor a page fault in the delivery path."""
    t = torch.empty((int(rows), int(width)), dtype=torch.float64, pin_memory=True)
    return t, t.numpy()


PINNED_RESULT_MAX_BYTES = 16 << 30   # larger results (e.g. rank 0 collecting 8 GPUs x 1000 frames) stay pageable


def try_pinned_rows(torch, rows, width):
    """pinned_rows, or None when the block is larger than PINNED_RESULT_MAX_BYTES or the host refuses to pin it
    (the caller then delivers into pageable memory through the staged drain)."""
    if int(rows) * int(width) * 8 > PINNED_RESULT_MAX_BYTES:
        return None
    try:
        return pinned_rows(torch, rows, width)
    except RuntimeError:
        return None


class HostDrain:
    """Device -> host pipeline that overlaps the D2H copy of finished batches with the solve
    of the next one: rows are copied on a side stream into a ring of pinned staging buffers
    and a worker thread moves them into the caller's (pageable) numpy array.  The solver's
    host thread sits inside libmof_b200 (GIL released) meanwhile."""

    def __init__(self, torch, device, max_rows, width, n_stages=4, copy_threads=None):
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.torch, self.device = torch, device
        self.max_rows, self.width = int(max_rows), int(width)
        self.n_stages = n_stages
        self.stages = None                # pinned staging ring, copy pool and worker thread: created by the first
        self.free = None                  # submit(); the direct-DMA path (submit_pinned) never needs them
        self.stream = torch.cuda.Stream(device=device)
        self.queue = queue.Queue()
        if copy_threads is None:          # staging -> destination memcpy is the slow half of the drain
            import os
            copy_threads = max(4, min(12, (os.cpu_count() or 8) - 2))
        self.pool = None
        self.copy_threads = copy_threads
        self.count = 0
        self.error = None
        self.thread = None
        self.closed = False

    def _start_staging(self):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        torch = self.torch
        self.stages = [torch.empty((self.max_rows, self.width), dtype=torch.float64).pin_memory() for _ in range(self.n_stages)]
        self.free = [threading.Event() for _ in range(self.n_stages)]
        for f in self.free:
            f.set()
        self.pool = ThreadPoolExecutor(self.copy_threads)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def close(self):
        """Stop the worker thread and the copy pool and drop the pinned staging buffers (a replaced or
        garbage-collected drain must not keep ~2.7 GB of pinned memory and a thread alive)."""
        if self.closed:
            return
        self.closed = True
        if self.thread is not None:
            self.queue.put(None)
            self.thread.join()
            self.thread = None
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
        self.stages = None

    def fits(self, rows, width):
        return rows <= self.max_rows and width == self.width

    def on_side_stream(self, fn):
        """Run ``fn()`` with the side stream current, after everything queued so far on the
        compute stream (used for the per-batch NCCL gather).  Returns fn's result."""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            return fn()

    def submit(self, src_dev, dst_host):
        """Queue rows ``src_dev`` (device (r, width), produced on the current stream or on the
        side stream) for delivery into ``dst_host`` (numpy (r, width)); long blocks are cut into
        staging-buffer-sized pieces."""
        rows = int(src_dev.shape[0])
        if rows > self.max_rows:
            for a in range(0, rows, self.max_rows):
                self.submit(src_dev[a:a + self.max_rows], dst_host[a:a + self.max_rows])
            return
        torch = self.torch
        if self.closed:
            raise RuntimeError("HostDrain is closed")
        if self.stages is None:
            self._start_staging()
        i = self.count % len(self.stages)
        self.count += 1
        self.free[i].wait()
        self.free[i].clear()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            r = int(src_dev.shape[0])
            self.stages[i][:r].copy_(src_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.queue.put((i, r, dst_host, done, src_dev))

    def submit_pinned(self, src_dev, dst_pinned):
        """Rows ``src_dev`` straight into ``dst_pinned`` (a pinned host tensor view of the same shape)
        with one asynchronous copy on the side stream: no staging buffer and no host memcpy.
        finish() waits for it."""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dst_pinned.copy_(src_dev, non_blocking=True)
        src_dev.record_stream(self.stream)
        self.direct_pending = True

    def _run(self):
        while True:
            item = self.queue.get()
            if item is None:
                self.queue.task_done()
                return
            i, r, dst, done, _keep = item
            try:
                done.synchronize()
                src = self.stages[i][:r].numpy()
                parts = max(1, min(self.copy_threads, r))
                edges = [r * q // parts for q in range(parts + 1)]
                list(self.pool.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(edges[:-1], edges[1:])))
            except Exception as exc:   # surfaced by finish()
                self.error = exc
            finally:
                self.free[i].set()
                self.queue.task_done()

    def finish(self):
        if self.thread is not None:
            self.queue.join()
        if getattr(self, "direct_pending", False):
            self.stream.synchronize()
            self.direct_pending = False
        if self.error is not None:
            err, self.error = self.error, None
            raise err


def frame_dt(t_k, k0, k1):
    """dt[k] = t_k[k+1] - t_k[k] evaluated in fp64 exactly like the reference's
    ``t_k[k + 1] - t_k[k]`` on Python floats (compute_optical_flow.py:125)."""
    return np.array([t_k[k + 1] - t_k[k] for k in range(k0, k1)], dtype=np.float64)

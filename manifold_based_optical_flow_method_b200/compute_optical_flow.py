"""Drop-in for the reference's ``utils/compute_optical_flow.py`` (hot-path functions).

Same function names and positional signatures as the reference:

    compute_geometrical_quantities(coordinates, normals, triangles, areas)
        -> (a2, grad_w, e, integral_wi_wj, execution_time)            # reference :27-97
    compute_velocity_field(processes_num, time_steps, a2, grad_w, e, integral_wi_wj,
                           triangles, t_k, areas, lambda_, I_k, I_k_2)
        -> (V_k, execution_time)                                       # reference :152-194
    worker(k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k_k, I_k_kplus1)
        -> V                                                           # reference :100-149
    reshape_and_save_data(data, file_path)                             # reference :314-320

Differences a caller can observe (all documented in INTEGRATION.md):
  * ``a2`` is a ``MeshOperator`` handle (device-resident geometry + a2 block values) rather
    than a scipy lil_matrix; ``a2.tocsr()`` gives the reference's matrix.
  * ``processes_num`` (the reference's Pool size) is ignored: frames are batched on the
    GPU of this process; under torchrun (torch.distributed initialised) they are also
    sharded across ranks and gathered with NCCL (see distributed.py).
  * the linear systems are solved by preconditioned CG (SSOR on a level-scheduled natural ordering by
    default; settings["precond"] = "ssor" for the block-multicolour ordering, "jacobi" for 2x2 block
    Jacobi) to ||b-Ax||/||b|| <= 1e-12 instead
    of SuperLU; frames that fail to converge raise ``UnconvergedError`` (the reference
    would return NaNs with a MatrixRankWarning) unless ``allow_unconverged`` is set.
  * float32 mesh arrays are promoted to float64 (the reference computes grad_w in
    float32 when pyvista hands it float32 points).
There is no CPU fallback: without a CUDA device / the compiled library these functions raise.
"""
import time

import numpy as np

from . import _lib
from .mesh import MeshOperator
from .solver import (DEFAULT_MAX_ITER, DEFAULT_OMEGA, DEFAULT_OMEGA_LEVEL, DEFAULT_PRECOND, DEFAULT_TOL, REORDER_OF,
                     SolveInfo, UnconvergedError, VelocitySolver, frame_dt)

# solver settings (module-level so the reference's positional signatures stay untouched)
settings = {
    "tol": DEFAULT_TOL,
    "max_iter": DEFAULT_MAX_ITER,
    "precond": DEFAULT_PRECOND,    # "ssor": block-multicolour SSOR (Eisenstat form); "ssor_level": the same on the
                                   # level-scheduled natural ordering; "jacobi": 2x2 block Jacobi
    "omega": None,                 # SSOR relaxation factor; None = 1.4 ("ssor") / 1.9 ("ssor_level")
    "batch_groups": None,          # None = sized from free device memory (<= 32 groups of 32 frames)
    "streams": None,               # None = solver default (1; 2 = batches on two concurrent streams)
    "allow_unconverged": False,
    "pinned_results": True,        # host results land in pinned memory by direct DMA (False: pageable numpy via staging)
    "device": None,                # None = current CUDA device
    "streamed_upload": True,       # host signal copied in chunks on a side stream, K1 runs chunk by chunk as rows arrive
    "numa_bind": True,             # multi-GPU host delivery: pin each rank to the CPUs of its GPU's NUMA node (best effort)
}

last_solve_info = None             # SolveInfo of the most recent compute_velocity_field / worker call
_solvers = {}


def compute_geometrical_quantities(coordinates, normals, triangles, areas):
    """Reference :27-97.  -> (a2 handle, grad_w (F,3,3), e (N,2,3), integral_wi_wj (F,2), seconds)"""
    start = time.time()
    # the vertex numbering is chosen for the preconditioner: colour-major patches for the SSOR
    # sweeps, Cuthill-McKee (smallest gather window) for block Jacobi
    op = MeshOperator(coordinates, normals, triangles, areas, device=settings["device"],
                      reorder=REORDER_OF[settings["precond"]])
    execution_time = time.time() - start
    return op, op.grad_w, op.e, op.integral_wi_wj, execution_time


def _operator(a2, triangles):
    if isinstance(a2, MeshOperator):
        return a2
    raise TypeError(
        "a2 must be the handle returned by this module's compute_geometrical_quantities (a MeshOperator); "
        "to reuse a matrix computed by the reference build one with MeshOperator.from_reference_a2(...)")


def _solver(op):
    key = id(op)
    precond = "ssor" if op.pattern.n_colors > 0 else ("ssor_level" if op.pattern.n_levels > 0 else "jacobi")
    if settings["precond"] == "jacobi":
        precond = "jacobi"          # block Jacobi runs on any numbering
    omega = settings["omega"]
    if omega is None:
        omega = DEFAULT_OMEGA_LEVEL if precond == "ssor_level" else DEFAULT_OMEGA
    omega = float(omega)
    s = _solvers.get(key)
    streams = settings["streams"]
    if s is None or s.op is not op or s.precond != precond or (precond != "jacobi" and s.omega != omega) \
            or (settings["batch_groups"] is not None and s.batch_groups != settings["batch_groups"]) \
            or (streams is not None and s.n_streams != streams):
        for old in _solvers.values():     # one mesh at a time keeps device memory bounded; stop the old drain thread
            old.close()
        _solvers.clear()
        kw = {} if streams is None else {"n_streams": streams}
        s = _solvers[key] = VelocitySolver(op, batch_groups=settings["batch_groups"], precond=precond, omega=omega, **kw)
    s.tol, s.max_iter = settings["tol"], settings["max_iter"]
    return s


def _check_converged(info, first_frame=0):
    bad = np.nonzero((info.status != _lib.STATUS_CONVERGED) & (info.status != _lib.STATUS_ZERO_RHS))[0]
    if len(bad) and not settings["allow_unconverged"]:
        raise UnconvergedError(info, [int(b) + first_frame for b in bad])


def _upload_for_solve(op, I_k, I_k_2, n, first=0):
    """_upload_signals for a solve that follows at once -> (I_dev, I2_dev, upload).  With settings["streamed_upload"]
    and both signals being the same contiguous float64 host array (what S3 passes, S3:115-116) the copy is issued
    in chunks on a side stream and ``upload`` is the ``SignalUpload`` in flight: the solver assembles every chunk of
    frames as its rows arrive.  Otherwise a blocking upload and ``upload`` = None."""
    torch = _lib.require_cuda()
    N = op.n_vertices
    if settings["streamed_upload"] and I_k_2 is I_k and not isinstance(I_k, (list, tuple)) and n > 0:
        src = None
        if torch.is_tensor(I_k):
            if I_k.device.type == "cpu" and I_k.dtype == torch.float64 and I_k.ndim == 2 and I_k.shape[1] == N \
                    and I_k.shape[0] >= first + n + 1 and I_k[first:first + n + 1].is_contiguous():
                src = I_k[first:first + n + 1]
        else:
            a = np.asarray(I_k)
            if a.ndim == 2 and a.shape[1] == N and a.shape[0] >= first + n + 1 and a.dtype == np.float64 \
                    and a[first:first + n + 1].flags.c_contiguous:
                src = torch.from_numpy(a[first:first + n + 1])
        if src is not None:
            from .solver import SignalUpload
            up = SignalUpload(torch, op.device, src, n + 1, N)
            return up.tensor, up.tensor, up
    I_dev, I2_dev = _upload_signals(op, I_k, I_k_2, n, first)
    return I_dev, I2_dev, None


def _upload_signals(op, I_k, I_k_2, n, first=0):
    """Rows first .. first+n of the (T,N) signals -> device float64 tensors whose row 0 is
    frame ``first``.  Frame k reads I_k[k] and I_k_2[k+1] (:174-175)."""
    torch = _lib.require_cuda()
    N = op.n_vertices

    def rows(a, lo, hi, name):
        if torch.is_tensor(a):                       # torch input (CPU or already on a GPU): no numpy detour
            if a.ndim != 2 or a.shape[1] != N or a.shape[0] < hi:
                raise ValueError(f"{name} must have shape (>= {hi}, {N}), got {tuple(a.shape)}")
            return a[lo:hi].to(device=op.device, dtype=torch.float64).contiguous()
        if isinstance(a, (list, tuple)):
            a = np.asarray(a[lo:hi], dtype=np.float64)
            lo, hi = 0, len(a)
        a = np.asarray(a)
        if a.ndim != 2 or a.shape[1] != N or a.shape[0] < hi:
            raise ValueError(f"{name} must have shape (>= {hi}, {N}), got {a.shape}")
        return torch.from_numpy(np.ascontiguousarray(a[lo:hi], dtype=np.float64)).to(op.device)

    if I_k_2 is I_k:
        I_dev = rows(I_k, first, first + n + 1, "I_k")
        return I_dev, I_dev
    # row 0 of the second tensor (I_k_2[first]) is never read; it only keeps row k+1 aligned
    return rows(I_k, first, first + n, "I_k"), rows(I_k_2, first, first + n + 1, "I_k_2")


def solve_on_device(op, I_dev, I2_dev, t_k, lambda_, k0, k1, V_dev=None, on_batch=None, upload=None):
    """Frames k0..k1-1 from device-resident signals (rows are absolute frame indices).
    -> (V_dev (k1-k0, 2N) device tensor, SolveInfo).  Used by bench.py (inputs resident in
    HBM) and by distributed.py.  upload: the SignalUpload still filling I_dev (k0 must be 0), see _upload_signals."""
    torch = _lib.require_cuda()
    s = _solver(op)
    dt = torch.from_numpy(frame_dt(t_k, k0, k1)).to(op.device)
    if upload is not None and k0 == 0:
        return s.solve_frames(I_dev, I2_dev, dt, lambda_, V_dev, on_batch=on_batch, upload=upload)
    if upload is not None:
        upload.wait_rows(torch, op.device, k1 + 1)
    return s.solve_frames(I_dev[k0:k1 + 1], I2_dev[k0:k1 + 1], dt, lambda_, V_dev, on_batch=on_batch)


def solve_to_host(op, I_dev, I2_dev, t_k, lambda_, n, upload=None):
    """solve_on_device for frames 0..n-1 with the results drained to a host array batch by
    batch, overlapped with the solve of the next batch.  -> (V (n, 2N) numpy, SolveInfo)"""
    s = _solver(op)
    drain = s.drain(2 * op.n_vertices)
    from .solver import try_pinned_rows
    pinned = try_pinned_rows(s.torch, n, 2 * op.n_vertices) if settings["pinned_results"] else None
    if pinned is not None:
        V_t, V = pinned
        on_batch = lambda k0, k1, Vd: drain.submit_pinned(Vd, V_t[k0:k1])
    else:
        V = np.empty((n, 2 * op.n_vertices), dtype=np.float64)
        on_batch = lambda k0, k1, Vd: drain.submit(Vd, V[k0:k1])
    _, info = solve_on_device(op, I_dev, I2_dev, t_k, lambda_, 0, n, on_batch=on_batch, upload=upload)
    drain.finish()
    return V, info


def tune_omega(a2, triangles, t_k, lambda_, I_k, I_k_2, candidates=(1.7, 1.8, 1.9), n_frames=32, apply=True):
    """Pick the SSOR relaxation factor for a data set: solve its first ``n_frames`` frames (one 32-frame group)
    with every candidate and keep the one with the fewest iterations (the iteration count of this solver
    depends on the input: ~143 at omega 1.9 for smooth travelling waves at ico7, fewer at ~1.7-1.8 for
    wrapped-phase input).  No counterpart in the reference (spsolve has no parameters); the
    result does not change what is computed, only how fast.  -> (omega, {candidate: mean iterations});
    with ``apply`` the choice is stored in ``settings["omega"]``."""
    torch = _lib.require_cuda()
    op = _operator(a2, triangles)
    n = min(int(n_frames), len(t_k) - 1)
    if n <= 0:
        raise ValueError("need at least two frames")
    I_dev, I2_dev = _upload_signals(op, I_k, I_k_2, n)
    old = settings["omega"]
    report = {}
    try:
        for w in candidates:
            settings["omega"] = float(w)
            s = _solver(op)
            if not s.ssor:
                raise ValueError("tune_omega needs an SSOR preconditioner (settings['precond'] 'ssor_level' or 'ssor')")
            _, info = solve_on_device(op, I_dev, I2_dev, list(t_k[:n + 1]), lambda_, 0, n)
            report[float(w)] = float(np.mean(info.iterations)) if info.converged else float("inf")
    finally:
        settings["omega"] = old
    best = min(report, key=report.get)
    if apply:
        settings["omega"] = best
    return best, report


def compute_velocity_field(processes_num, time_steps, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas,
                           lambda_, I_k, I_k_2):
    """Reference :152-194.  Frame k (k = 0 .. time_steps-2) uses (I_k[k], I_k_2[k+1]) (:174-175).
    -> (V_k: list of time_steps-1 arrays (2N,), V_k[k][i + N*alpha]; seconds)"""
    global last_solve_info
    torch = _lib.require_cuda()
    op = _operator(a2, triangles)
    op.use_geometry(grad_w, e, integral_wi_wj, areas)
    start_time = time.time()
    n = int(time_steps) - 1
    N = op.n_vertices
    if n <= 0:
        return [], 0.0
    from . import distributed
    if distributed.world_size() > 1:
        V, info = distributed.compute_velocity_field_sharded(op, n, t_k, lambda_, I_k, I_k_2)
    else:
        I_dev, I2_dev, upload = _upload_for_solve(op, I_k, I_k_2, n)
        V, info = solve_to_host(op, I_dev, I2_dev, t_k, lambda_, n, upload=upload)
    execution_time = time.time() - start_time
    info.seconds = execution_time
    last_solve_info = info
    _check_converged(info)
    return [V[k] for k in range(n)], execution_time


def worker(k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k_k, I_k_kplus1):
    """Reference :100-149: one frame.  -> V (2N,) float64."""
    global last_solve_info
    torch = _lib.require_cuda()
    op = _operator(a2, triangles)
    op.use_geometry(grad_w, e, integral_wi_wj, areas)
    s = _solver(op)
    I_now = torch.from_numpy(np.ascontiguousarray(np.asarray(I_k_k, dtype=np.float64)).reshape(1, -1)).to(op.device)
    I_next = torch.from_numpy(np.ascontiguousarray(np.asarray(I_k_kplus1, dtype=np.float64)).reshape(1, -1)).to(op.device)
    dt = torch.from_numpy(frame_dt(t_k, k, k + 1)).to(op.device)
    V_dev = torch.empty((1, 2 * op.n_vertices), dtype=torch.float64, device=op.device)
    info = s.solve_batch(I_now, I_next, dt, lambda_, V_dev)
    last_solve_info = info
    _check_converged(info, first_frame=k)
    return V_dev[0].cpu().numpy()


def reshape_and_save_data(data, file_path):
    """Reference :314-320: pd.DataFrame(data.reshape(n, -1)).to_csv(file_path).  Same bytes on disk
    (header ",0,1,...", Python-repr floats), written by the multi-threaded C++ writer of
    libmof_b200 (csrc/csvio.cpp): a (999, 327684) field file is 6.5 GB of text."""
    import ctypes
    if isinstance(data, list):
        data = np.array(data)
    reshaped = np.ascontiguousarray(np.asarray(data).reshape(np.asarray(data).shape[0], -1), dtype=np.float64)
    _lib.check(_lib.load().mof_csv_write(str(file_path).encode(), reshaped.ctypes.data, reshaped.shape[0],
                                         reshaped.shape[1], 0))


def load_potentials(csv_path):
    """Reference :203-207: pd.read_csv(csv_path, sep=',', header='infer', index_col=0).values, read by
    the multi-threaded C++ parser (correctly rounded doubles = pandas float_precision='round_trip')."""
    import ctypes
    lib = _lib.load()
    rows, cols = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.mof_csv_dims(str(csv_path).encode(), ctypes.byref(rows), ctypes.byref(cols)))
    out = np.empty((rows.value, cols.value), dtype=np.float64)
    _lib.check(lib.mof_csv_read(str(csv_path).encode(), out.ctypes.data, rows.value, cols.value, 0))
    return out


__all__ = ["compute_geometrical_quantities", "compute_velocity_field", "worker", "reshape_and_save_data",
           "load_potentials", "solve_on_device", "settings", "SolveInfo", "UnconvergedError", "MeshOperator"]

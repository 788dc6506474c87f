"""Drop-in for the numerical functions of the reference's ``S5_compute_wave_v.py`` ("next" row 1
of SURVEY.md section 8f): wave speed = |d/dt| / |surface gradient| of a phase (or amplitude) map.

    wave_velocity_phase(surface, phases, dt, time_steps, e)          # reference :79-123
    wave_velocity_amplitude(surface, potentials, dt, time_steps, e)  # reference :14-58
    compute_grad_M_I(coordinates, triangles, potentials, surface, areas)   # reference :136-171
    compute_temporal_gradient_phase(data, dt)                         # reference :60-77
    angle_subtract(f1, f2, angleFlag=True)                            # reference :224-233

``surface`` is anything with the pyvista members the reference uses (``points``, ``faces``,
``compute_cell_sizes(...)['Area']``) -- a real ``pyvista.PolyData`` or ``synthetic.SurfaceMesh``.
The per-vertex face lists come from the mesh pattern (ascending face index) instead of
``surface.point_cell_ids``.  csrc/wave.cu folds the per-face gradients, the area-weighted vertex average, the
tangent projection and the basis coefficients -- all linear in the signal -- into two coefficients per
(vertex, ring vertex) pair once per call, transposes the signal into the frame-minor layout, and runs one kernel
that evaluates a sparse row product per vertex and frame (lane = frame), the (wrapped) time derivative and the
division, and writes the (T, N) result directly; no CPU fallback.  Under torchrun the frames are sharded over the GPUs with a
time-derivative halo (``wave_speed_device`` / the public functions do it transparently), no collective on
the data path.
"""
import ctypes

import numpy as np

from . import _lib
from .mesh import MeshOperator

_ops = {}


def _surface_arrays(surface):
    coordinates = np.asarray(surface.points, dtype=np.float64)
    triangles = np.asarray(surface.faces).reshape(-1, 4)[:, 1:]
    areas = np.asarray(surface.compute_cell_sizes(length=False, volume=False)['Area'], dtype=np.float64)
    return coordinates, triangles, areas


def _operator(coordinates, triangles, areas, e=None):
    """Mesh handle cached per (coordinates, triangles) buffers; normals only define e, which S5
    receives from the caller, so the handle is built with dummy normals and e is uploaded."""
    import hashlib
    digest = hashlib.blake2b(digest_size=16)             # content hash: equal sums do not mean equal meshes
    digest.update(np.ascontiguousarray(coordinates, dtype=np.float64).tobytes())
    digest.update(np.ascontiguousarray(triangles, dtype=np.int64).tobytes())
    key = (coordinates.shape, triangles.shape, digest.hexdigest())
    op = _ops.get(key)
    if op is None:
        _ops.clear()
        normals = np.zeros_like(coordinates)
        normals[:, 2] = 1.0
        # reference vertex order (reorder=0): the (T, N) <-> frame-minor transposes of csrc/wave.cu are then fully
        # coalesced, and the ring gathers move whole 256-byte lines whatever the numbering
        op = _ops[key] = MeshOperator(coordinates, normals, triangles, areas, reorder=0)
    op.use_geometry(None, e, None, areas)
    return op


def halo_rows(k0, k1, T, phase_mode):
    """Rows [a, b) of a T-frame trial that a call producing frames [k0, k1) must be given: one frame either
    side for the central difference, the three end frames where np.gradient's one-sided formula applies."""
    a, b = max(k0 - 1, 0), min(k1 + 1, T)
    if not phase_mode and k1 > k0:
        if k0 == 0:
            b = max(b, min(3, T))
        if k1 == T:
            a = min(a, max(T - 3, 0))
    return a, b


def wave_speed_device(op, d_rows, t_first, T_trial, out0, n_out, dt, phase_mode, want_grad=False, want_wave=True,
                      work=None, grad_out=None, wave_out=None):
    """mof_wave_speed on device-resident rows ``d_rows`` ((n_rows, N) float64, reference vertex order; row 0 is
    frame ``t_first`` of a ``T_trial``-frame trial) -> (grad (n_out, N, 3) or None, wave (n_out, N) or None) for
    rows out0 .. out0+n_out-1.  bench.py and the sharded path call this.  ``work`` / ``grad_out`` / ``wave_out``:
    optional preallocated scratch and result tensors (a repeated caller then allocates nothing per call)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n_rows, N = int(d_rows.shape[0]), int(d_rows.shape[1])
    if N != op.n_vertices:
        raise ValueError(f"data must have shape (T, {op.n_vertices}), got {tuple(d_rows.shape)}")
    assert d_rows.stride(1) == 1
    ms = op.struct()
    need = int(lib.mof_wave_work_doubles(ctypes.byref(ms), n_rows, int(want_grad), int(want_wave)))
    if work is None or work.numel() < need:
        work = torch.empty((need,), dtype=torch.float64, device=op.device)
    def result(given, shape):
        if given is None:
            return torch.empty(shape, dtype=torch.float64, device=op.device)
        if tuple(given.shape) != shape or given.dtype != torch.float64 or not given.is_contiguous() or given.device != op.device:
            raise ValueError(f"result tensor must be a contiguous float64 {shape} on {op.device}")
        return given
    grad = result(grad_out, (int(n_out), N, 3)) if want_grad else None
    wave = result(wave_out, (int(n_out), N)) if want_wave else None
    st = torch.cuda.current_stream(op.device).cuda_stream
    _lib.check(lib.mof_wave_speed(ctypes.byref(ms), n_rows, int(out0), int(n_out), int(t_first), int(T_trial),
                                  d_rows.data_ptr(), d_rows.stride(0), float(dt), 1 if phase_mode else 0,
                                  grad.data_ptr() if want_grad else None, wave.data_ptr() if want_wave else None,
                                  work.data_ptr(), st))
    return grad, wave


def _run(op, data, dt, phase_mode, want_grad, want_wave):
    """All frames of a trial.  One GPU: one call.  Under torchrun: every rank computes a contiguous range of
    frames from its rows plus the halo (halo_rows) and the ranges are gathered (S5:79-123 is a loop over
    independent frames once the time derivative has its neighbours)."""
    torch = _lib.require_cuda()
    from . import distributed
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
    T, N = data.shape
    if N != op.n_vertices:
        raise ValueError(f"data must have shape (T, {op.n_vertices}), got {data.shape}")
    world, r = distributed.world_size(), distributed.rank()
    if world == 1:
        d = torch.from_numpy(data).to(op.device)
        grad, wave = wave_speed_device(op, d, 0, T, 0, T, dt, phase_mode, want_grad, want_wave)
        return (grad.cpu().numpy() if want_grad else None), (wave.cpu().numpy() if want_wave else None)
    counts = distributed.shard_counts(T, world)
    k0, k1 = distributed.shard_range(T, world, r)
    a, b = halo_rows(k0, k1, T, phase_mode)
    if k1 > k0:
        d = torch.from_numpy(data[a:b]).to(op.device)
        grad, wave = wave_speed_device(op, d, a, T, k0 - a, k1 - k0, dt, phase_mode, want_grad, want_wave)
    else:
        grad = torch.empty((0, N, 3), dtype=torch.float64, device=op.device) if want_grad else None
        wave = torch.empty((0, N), dtype=torch.float64, device=op.device) if want_wave else None
    out_g = distributed.gather_rows(grad.reshape(k1 - k0, 3 * N), counts).reshape(T, N, 3).cpu().numpy() if want_grad else None
    out_w = distributed.gather_rows(wave, counts).cpu().numpy() if want_wave else None
    return out_g, out_w


def compute_grad_M_I(coordinates, triangles, potentials, surface, areas):
    """Reference :136-171 -> grad_point (time_steps, point_num, 3)."""
    op = _operator(np.asarray(coordinates, dtype=np.float64), np.asarray(triangles), np.asarray(areas, dtype=np.float64))
    return _run(op, potentials, 1.0, True, True, False)[0]


def wave_velocity_phase(surface, phases, dt, time_steps, e):
    """Reference :79-123 -> wave_velocity (time_steps, point_num), signed, rad/s per unit length."""
    coordinates, triangles, areas = _surface_arrays(surface)
    op = _operator(coordinates, triangles, areas, np.asarray(e, dtype=np.float64).reshape(-1, 2, 3))
    return _run(op, np.asarray(phases)[:time_steps], dt, True, False, True)[1]


def wave_velocity_amplitude(surface, potentials, dt, time_steps, e):
    """Reference :14-58 (np.gradient(edge_order=2) time derivative)."""
    coordinates, triangles, areas = _surface_arrays(surface)
    op = _operator(coordinates, triangles, areas, np.asarray(e, dtype=np.float64).reshape(-1, 2, 3))
    return _run(op, np.asarray(potentials)[:time_steps], dt, False, False, True)[1]


def angle_subtract(f1, f2, angleFlag=True):
    """Reference :224-233 (host helper; the kernel applies the same formula per element)."""
    if angleFlag:
        return np.mod(f1 - f2 + np.pi, 2 * np.pi) - np.pi
    return f1 - f2


def compute_temporal_gradient_phase(data, dt):
    """Reference :60-77; computed by the kernel on a flat two-vertex mesh would be overkill, so
    this thin helper evaluates the same three formulas with numpy for callers that want the
    derivative alone (wave_velocity_phase fuses it into the kernel)."""
    data = np.asarray(data, dtype=np.float64)
    g = np.zeros_like(data)
    g[0] = angle_subtract(data[1], data[0]) / dt
    g[1:-1] = angle_subtract(data[2:], data[:-2]) / (2 * dt)
    g[-1] = angle_subtract(data[-1], data[-2]) / dt
    return g

"""Drop-in for the numerical core of the reference's ``S2_interpolate.py`` and
``S2_interpolate_phases.py`` ("next" row 4 of SURVEY.md section 8f): the radial-basis interpolation
of electrode signals onto the surface vertices that produces the (T, N) input of the optical-flow
solve.

    interpolation(surface, ieeg_data_array, coordinates, start_sample, end_sample, save_path, ifsave)   # S2:22-53
    rbf_interpolate_device(data, coordinates, vertices, phase=False) -> torch tensor (T, N) on the GPU
    rbf_interpolate(data, coordinates, vertices, phase=False)        -> numpy (T, N)
    rbf_epsilon(coordinates)                                          # scipy's default shape parameter

The reference builds ``scipy.interpolate.Rbf(x, y, z, frame)`` for every frame (same m x m system
solved T times on the CPU, then a dense N x m kernel matrix per frame: 0.16 s per frame at 164k
vertices and 128 electrodes).  Here the system is factorised once on the GPU, all frames are solved
against it, and one GEMM-shaped CUDA kernel evaluates every frame (csrc/rbf.cu); the result can
stay in HBM and go straight into ``compute_optical_flow.solve_on_device`` without ever crossing
PCIe.  Same defaults as ``Rbf``: multiquadric, smooth 0, epsilon from the electrodes' bounding box.

``surface``: a path (read with pyvista, if installed), or anything with ``points``, or an (N, 3)
array.  Unlike the reference, ``interpolation`` also returns the interpolated array.  A singular
system (two electrodes at the same position) raises ``numpy.linalg.LinAlgError`` like scipy.
No CPU fallback.
"""
import numpy as np

from . import _lib


def rbf_epsilon(coordinates):
    """Default ``epsilon`` of scipy.interpolate.Rbf: (prod(non-zero bounding-box edges) / m) ** (1 / n_edges)."""
    c = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    edges = c.max(axis=0) - c.min(axis=0)
    edges = edges[np.nonzero(edges)]
    if edges.size == 0:
        raise ValueError("all electrodes coincide")
    return float(np.power(np.prod(edges) / len(c), 1.0 / edges.size))


def _vertices_of(surface):
    if isinstance(surface, (str, bytes)) or hasattr(surface, "__fspath__"):
        try:
            import pyvista as pv
        except ImportError as exc:                                   # pragma: no cover - pyvista absent offline
            raise RuntimeError("reading a surface file needs pyvista; pass the surface object or its points instead") from exc
        surface = pv.read(surface)
    pts = getattr(surface, "points", surface)
    return np.ascontiguousarray(np.asarray(pts, dtype=np.float64)).reshape(-1, 3)


def rbf_interpolate_device(data, coordinates, vertices, phase=False, epsilon=None, out=None):
    """data (T, m) real -- or complex, e.g. exp(1j * electrode phase) -- at the m electrodes
    ``coordinates`` (m, 3) -> device tensor (T, N) of the interpolant at ``vertices`` (N, 3);
    ``phase=True`` returns its angle in (-pi, pi] (S2_interpolate_phases.py:51-52).  ``vertices`` may
    already be a CUDA tensor."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    c = np.ascontiguousarray(np.asarray(coordinates, dtype=np.float64)).reshape(-1, 3)
    m = len(c)
    d = np.asarray(data)
    if d.ndim != 2 or d.shape[1] != m:
        raise ValueError(f"data must have shape (T, {m}), got {d.shape}")
    if phase or np.iscomplexobj(d):
        if not phase:
            raise ValueError("complex data needs phase=True (the reference takes np.angle of the complex interpolant)")
        d = np.asarray(d, dtype=np.complex128)
        rhs = np.concatenate([d.real, d.imag])                      # rows 0..T-1 real, T..2T-1 imaginary
    else:
        rhs = np.asarray(d, dtype=np.float64)
    T = d.shape[0]
    eps = rbf_epsilon(c) if epsilon is None else float(epsilon)
    if torch.is_tensor(vertices):
        v_d = vertices.to(device=dev, dtype=torch.float64).contiguous().reshape(-1, 3)
    else:
        v_d = torch.from_numpy(_vertices_of(vertices)).to(dev)
    N = int(v_d.shape[0])
    if out is None:
        out = torch.empty((T, N), dtype=torch.float64, device=dev)
    elif tuple(out.shape) != (T, N) or out.dtype != torch.float64 or not out.is_cuda or out.stride(1) != 1:
        raise ValueError("out must be a float64 CUDA tensor of shape (T, N) with unit column stride")
    if T == 0:
        return out
    c_d = torch.from_numpy(c).to(dev)
    rhs_d = torch.from_numpy(np.ascontiguousarray(rhs)).to(dev)
    n_rhs = int(rhs_d.shape[0])
    lu = torch.empty((m, m), dtype=torch.float64, device=dev)
    piv = torch.empty((m,), dtype=torch.int32, device=dev)
    w = torch.empty((m, n_rhs), dtype=torch.float64, device=dev)
    info = torch.zeros((1,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.mof_rbf_fit(m, c_d.data_ptr(), eps, n_rhs, rhs_d.data_ptr(), m, lu.data_ptr(), piv.data_ptr(),
                               w.data_ptr(), info.data_ptr(), st))
    bad = int(info.item())
    if bad:
        raise np.linalg.LinAlgError(f"RBF matrix is singular (zero pivot {bad - 1}): duplicated electrode positions?")
    _lib.check(lib.mof_rbf_evaluate(N, m, T, v_d.data_ptr(), c_d.data_ptr(), eps, w.data_ptr(), 1 if phase else 0,
                                    out.data_ptr(), out.stride(0), st))
    return out


def rbf_interpolate(data, coordinates, vertices, phase=False, epsilon=None):
    return rbf_interpolate_device(data, coordinates, vertices, phase, epsilon).cpu().numpy()


def _save(values, save_path):
    from .compute_optical_flow import reshape_and_save_data          # pandas-dialect CSV, C++ writer
    reshape_and_save_data(values, save_path)


def interpolation(surface_path, ieeg_data_array, coordinates, start_sample, end_sample, save_path, ifsave):
    """Reference S2_interpolate.py:22-53: potentials (time, electrodes) -> (end-start, N) on the
    surface, saved as CSV when ``ifsave``.  Returns the array (the reference returns None)."""
    data = np.asarray(ieeg_data_array)[start_sample:end_sample]
    values = rbf_interpolate(data, coordinates, surface_path, phase=False)
    print(f"interpolated shape (t, vertices): {values.shape}")
    if ifsave is True:
        _save(values, save_path)
    return values


__all__ = ["interpolation", "rbf_interpolate", "rbf_interpolate_device", "rbf_epsilon"]

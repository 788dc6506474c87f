"""Drop-in for the numerical core of the reference's ``S7_winding_line.py`` ("next" row 3 of
SURVEY.md section 8f): over how many topological rings around a singular point does the velocity
field keep winding number +1 (node / focus) or -1 (saddle)?

    calculate_winding_numbers(surf, singularity_points, V_now, e, points, max_level=25)   # reference :120-165
    winding_numbers(...)  -> WindingResult     # all frames in one launch, flat arrays

``surf`` is anything with ``faces`` in the pyvista layout ([3, a, b, c, ...]) or ``triangles``
-- a real ``pyvista.PolyData`` or ``synthetic.SurfaceMesh``; the two things the reference asks of
it (``find_closest_point``, ``point_neighbors_levels``) are computed on the GPU from ``points`` and
the triangle list: nearest vertex by Euclidean distance (lowest index on ties) and breadth-first
topological rings.  One CUDA kernel (csrc/winding.cu), one CTA per singular point; no CPU fallback.

Differences from the reference: where the mesh has fewer than ``max_level`` rings around a point
whose rings all pass, the reference raises IndexError on the empty ring; here the count stops.  A
ring of more than 1024 vertices raises ``RuntimeError`` (25 rings of a cortical mesh hold a few
hundred).
"""
from dataclasses import dataclass

import numpy as np

from . import _lib
from .find_singularity_point import mesh_adjacency

RING_CAPACITY = 1024
_adjacency = {}


def _rings(triangles, n_vertices):
    """1-ring CSR of the mesh, cached for the last mesh seen (it is rebuilt from the triangle list
    with numpy: ~1 s at 328k faces, far more than the kernel)."""
    tri = np.ascontiguousarray(np.asarray(triangles), dtype=np.int64)
    key = (n_vertices, tri.shape, int(tri.sum()), int((tri[:, 0] * 3 + tri[:, 1] * 5 + tri[:, 2] * 7).sum()))
    hit = _adjacency.get(key)
    if hit is None:
        _adjacency.clear()
        ring_ptr, ring_idx, _ = mesh_adjacency(tri, n_vertices)
        hit = _adjacency[key] = (ring_ptr, ring_idx)
    return hit


@dataclass
class WindingResult:
    closest: np.ndarray     # (n,) int32  vertex the rings are grown from
    counts: np.ndarray      # (n,) int64  number of consecutive rings accepted (reference: winding_numbers_counts)
    types: np.ndarray       # (n,) int64  +1 / -1 fixed by the first ring, 0 if it is neither
    winding: np.ndarray     # (n, max_level) winding numbers evaluated, NaN after the stop


def _triangles_of(surf):
    tri = getattr(surf, "triangles", None)
    if tri is None:
        tri = np.asarray(surf.faces).reshape(-1, 4)[:, 1:]
    return np.ascontiguousarray(np.asarray(tri), dtype=np.int64)


def winding_numbers(triangles, coordinates, singularity_points, frame_of_point, V_k_coord, e, max_level=25):
    """All points of all frames in one launch.  singularity_points (n,3); frame_of_point (n,)
    indexes V_k_coord (n_frames, N, 3) (a single (N,3) field is frame 0); e (N,2,3)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    coords = np.ascontiguousarray(np.asarray(coordinates, dtype=np.float64)).reshape(-1, 3)
    N = len(coords)
    V = np.asarray(V_k_coord, dtype=np.float64)
    if V.ndim == 2:
        V = V[None]
    V = np.ascontiguousarray(V[:, :, :3])
    if V.shape[1] != N:
        raise ValueError(f"velocity fields must have shape (n_frames, {N}, 3), got {V.shape}")
    e_np = np.ascontiguousarray(np.asarray(e, dtype=np.float64)).reshape(N, 2, 3)
    pts = np.ascontiguousarray(np.asarray(singularity_points, dtype=np.float64)).reshape(-1, 3)
    n = len(pts)
    fop = np.array(np.broadcast_to(np.asarray(frame_of_point, dtype=np.int32), (n,)))
    if n and (fop.min() < 0 or fop.max() >= V.shape[0]):
        raise ValueError("frame_of_point out of range")
    max_level = int(max_level)
    if max_level < 1:
        raise ValueError("max_level must be positive")
    if n == 0:
        return WindingResult(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, max_level)))
    ring_ptr, ring_idx = _rings(triangles, N)
    up = lambda a: torch.from_numpy(np.array(a, copy=True, order='C')).to(dev)
    coords_d, V_d, e_d, rp_d, ri_d, pts_d, fop_d = up(coords), up(V), up(e_np), up(ring_ptr), up(ring_idx), up(pts), up(fop)
    closest = torch.empty((n,), dtype=torch.int32, device=dev)
    counts = torch.empty((n,), dtype=torch.int32, device=dev)
    types = torch.empty((n,), dtype=torch.int8, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    wind = torch.empty((n, max_level), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.mof_winding_numbers(N, V.shape[0], coords_d.data_ptr(), V_d.data_ptr(), e_d.data_ptr(), rp_d.data_ptr(),
                                       ri_d.data_ptr(), n, pts_d.data_ptr(), fop_d.data_ptr(), max_level, closest.data_ptr(),
                                       counts.data_ptr(), types.data_ptr(), wind.data_ptr(), status.data_ptr(), st))
    status_h = status.cpu().numpy()
    if status_h.any():
        bad = int(np.nonzero(status_h)[0][0])
        raise RuntimeError(f"winding numbers: a ring around point {bad} holds more than {RING_CAPACITY} vertices")
    return WindingResult(closest.cpu().numpy(), counts.cpu().numpy().astype(np.int64), types.cpu().numpy().astype(np.int64),
                         wind.cpu().numpy())


def calculate_winding_numbers(surf, singularity_points, V_now, e, points, max_level=25):
    """Reference :120-165 -> (winding_numbers_counts, types): one count per singular point, and
    the +1 / -1 type of every point whose first ring qualifies (in point order, like the
    reference's list, which skips the others)."""
    r = winding_numbers(_triangles_of(surf), points, singularity_points, 0, V_now, e, max_level)
    return [int(c) for c in r.counts], [int(t) for t in r.types if t != 0]


__all__ = ["calculate_winding_numbers", "winding_numbers", "WindingResult", "RING_CAPACITY"]

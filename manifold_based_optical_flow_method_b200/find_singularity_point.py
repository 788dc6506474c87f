"""Drop-in for the hot-path functions of the reference's ``utils/find_singularity_point.py``.

    process_V_k(V_k, e) -> V_k_coord                                        # reference :28-69
    find_singularity_points(coordinates, triangles, V_now, eps)
        -> (singularity_vertices, singularity_interiors, v_length_max)      # reference :140-189
    find_singularity_points_for_all_Vk(V_k_coord, coordinates, triangles, eps)
        -> list (per frame) of point coordinates                            # reference :530-558
    find_singularity_points_and_classify_for_all_Vk(V_k_coord, coordinates, triangles, eps, surface, e)
        -> (points per frame, "Node" / "Focus" / "Saddle" / "Indeterminate" per point)   # reference :561-605

plus the batched form the kernels natively produce, ``detect_singularities`` (all frames in
one call, flat index arrays + per-face Poincare index).

Observable differences: ``process_V_k`` returns one (T-1, N, 3) ndarray instead of nested
lists of (3,) arrays (S3 converts to ndarray right away, S3...:130); the interior-zero test
solves the reference's least-squares system in closed form inside the face plane, so (lam, mu)
agree to ~1e-13 rather than bitwise; a face whose projected vertex velocities are exactly
collinear (singular 2x2 system) is rejected, where numpy's lstsq would return a minimum-norm
answer.  No CPU fallback.
"""
import dataclasses

import numpy as np

from . import _lib

FRAMES_PER_CALL = 256


def _torch_dev(device=None):
    torch = _lib.require_cuda()
    return torch, torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


def tangent_to_xyz_device(V_dev, e_dev, want_speed=True, want_vmax=True):
    """K4 on device tensors: V_dev (n, 2N), e_dev (N,2,3) reference order.
    -> (Vxyz (n,N,3), speed (n,N) | None, vmax (n,) | None)"""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n, N = int(V_dev.shape[0]), int(e_dev.shape[0])
    dev = V_dev.device
    Vxyz = torch.empty((n, N, 3), dtype=torch.float64, device=dev)
    speed = torch.empty((n, N), dtype=torch.float64, device=dev) if want_speed else None
    vmax = torch.empty((n,), dtype=torch.float64, device=dev) if want_vmax else None
    st = torch.cuda.current_stream(dev).cuda_stream
    for k0 in range(0, n, 32768):
        k1 = min(n, k0 + 32768)
        _lib.check(lib.mof_tangent_to_xyz(
            N, k1 - k0, V_dev[k0:k1].data_ptr(), V_dev.stride(0), e_dev.data_ptr(), Vxyz[k0:k1].data_ptr(),
            speed[k0:k1].data_ptr() if want_speed else None, vmax[k0:k1].data_ptr() if want_vmax else None, st))
    return Vxyz, speed, vmax


def process_V_k(V_k, e):
    """Reference :28-69: V_k_coord[k][i] = V_k[k][i]*e[i][0] + V_k[k][i+N]*e[i][1].
    V_k: sequence of (2N,) arrays (or (T-1, 2N) array); e: (N,2,3) or (N,6).  -> ndarray (T-1, N, 3)"""
    torch, dev = _torch_dev()
    e_np = np.ascontiguousarray(np.asarray(e, dtype=np.float64)).reshape(-1, 2, 3)
    V_np = np.ascontiguousarray(np.asarray(V_k, dtype=np.float64))
    N = e_np.shape[0]
    if V_np.ndim != 2 or V_np.shape[1] != 2 * N:
        raise ValueError(f"V_k must have shape (frames, {2 * N}), got {V_np.shape}")
    e_dev = torch.from_numpy(e_np).to(dev)
    out = np.empty((V_np.shape[0], N, 3), dtype=np.float64)
    for k0 in range(0, V_np.shape[0], FRAMES_PER_CALL):
        k1 = min(V_np.shape[0], k0 + FRAMES_PER_CALL)
        Vxyz, _, _ = tangent_to_xyz_device(torch.from_numpy(V_np[k0:k1]).to(dev), e_dev, want_speed=False, want_vmax=False)
        out[k0:k1] = Vxyz.cpu().numpy()
    return out


def speed_magnitude(V_k, e):
    """V_c of S3_compute_v_and_detection_singularity.py:130-132 straight from the tangent
    coefficients (fused into K4).  -> ndarray (T-1, N)"""
    torch, dev = _torch_dev()
    e_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(e, dtype=np.float64)).reshape(-1, 2, 3)).to(dev)
    V_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(V_k, dtype=np.float64))).to(dev)
    _, speed, _ = tangent_to_xyz_device(V_dev, e_dev, want_speed=True, want_vmax=False)
    return speed.cpu().numpy()


@dataclasses.dataclass
class Singularities:
    """Flat per-batch result of K5; frame k owns vertex entries
    vertex_offsets[k]:vertex_offsets[k+1] and face entries face_offsets[k]:face_offsets[k+1]."""
    vertex_offsets: np.ndarray    # (n+1,) int64
    vertex_idx: np.ndarray        # (sum nv,) int32, ascending inside a frame
    face_offsets: np.ndarray      # (n+1,) int64
    face_idx: np.ndarray          # (sum nf,) int32, ascending inside a frame
    lam_mu: np.ndarray            # (sum nf, 2)
    P: np.ndarray                 # (sum nf, 3)  lam*A + mu*B + (1-lam-mu)*C
    index: np.ndarray             # (sum nf,) int8 per-face Poincare index (+1 node/focus, -1 saddle)
    v_length_max: np.ndarray      # (n,)

    def frame(self, k):
        v = slice(self.vertex_offsets[k], self.vertex_offsets[k + 1])
        f = slice(self.face_offsets[k], self.face_offsets[k + 1])
        return self.vertex_idx[v], self.face_idx[f], self.lam_mu[f], self.P[f], self.index[f]


def detect_singularities_device(coords_dev, tri_dev, Vxyz_dev, eps, vmax_dev=None):
    """K5 on device tensors: coords (N,3) f64, tri (F,3) int32, Vxyz (n,N,3) f64 -> Singularities"""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = Vxyz_dev.device
    n, N, F = int(Vxyz_dev.shape[0]), int(coords_dev.shape[0]), int(tri_dev.shape[0])
    st = torch.cuda.current_stream(dev).cuda_stream
    if vmax_dev is None:
        vmax_dev = torch.empty((n,), dtype=torch.float64, device=dev)
        _lib.check(lib.mof_vmax(N, n, Vxyz_dev.data_ptr(), vmax_dev.data_ptr(), st))
    nvc, nfc = -(-N // _lib.DETECT_CHUNK), -(-F // _lib.DETECT_CHUNK)
    vflag = torch.empty((n, N), dtype=torch.uint8, device=dev)
    fflag = torch.empty((n, max(F, 1)), dtype=torch.uint8, device=dev)
    vcnt = torch.zeros((n, nvc), dtype=torch.int32, device=dev)
    fcnt = torch.zeros((n, max(nfc, 1)), dtype=torch.int32, device=dev)
    totals = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    _lib.check(lib.mof_singularity_flags(N, F, n, coords_dev.data_ptr(), tri_dev.data_ptr(), Vxyz_dev.data_ptr(),
                                         vmax_dev.data_ptr(), float(eps), vflag.data_ptr(), fflag.data_ptr(),
                                         vcnt.data_ptr(), fcnt.data_ptr(), totals.data_ptr(), st))
    tot = totals.cpu().numpy().astype(np.int64)            # sync: exact output sizes
    voff = np.concatenate([[0], np.cumsum(tot[:, 0])])
    foff = np.concatenate([[0], np.cumsum(tot[:, 1])])
    nv, nf = int(voff[-1]), int(foff[-1])
    vertex_idx = torch.empty((max(nv, 1),), dtype=torch.int32, device=dev)
    face_idx = torch.empty((max(nf, 1),), dtype=torch.int32, device=dev)
    lam_mu = torch.empty((max(nf, 1), 2), dtype=torch.float64, device=dev)
    P = torch.empty((max(nf, 1), 3), dtype=torch.float64, device=dev)
    index = torch.empty((max(nf, 1),), dtype=torch.int8, device=dev)
    voff_dev = torch.from_numpy(voff[:-1].copy()).to(dev)
    foff_dev = torch.from_numpy(foff[:-1].copy()).to(dev)
    _lib.check(lib.mof_singularity_compact(N, F, n, coords_dev.data_ptr(), tri_dev.data_ptr(), Vxyz_dev.data_ptr(),
                                           vmax_dev.data_ptr(), vflag.data_ptr(), fflag.data_ptr(), vcnt.data_ptr(),
                                           fcnt.data_ptr(), voff_dev.data_ptr(), foff_dev.data_ptr(),
                                           vertex_idx.data_ptr(), face_idx.data_ptr(), lam_mu.data_ptr(), P.data_ptr(),
                                           index.data_ptr(), st))
    return Singularities(voff, vertex_idx[:nv].cpu().numpy(), foff, face_idx[:nf].cpu().numpy(), lam_mu[:nf].cpu().numpy(),
                         P[:nf].cpu().numpy(), index[:nf].cpu().numpy(), vmax_dev.cpu().numpy())


def _concat(parts):
    if len(parts) == 1:
        return parts[0]
    voff = [np.zeros(1, np.int64)]
    foff = [np.zeros(1, np.int64)]
    for p in parts:
        voff.append(p.vertex_offsets[1:] + voff[-1][-1])
        foff.append(p.face_offsets[1:] + foff[-1][-1])
    cat = lambda name: np.concatenate([getattr(p, name) for p in parts], axis=0)
    return Singularities(np.concatenate(voff), cat("vertex_idx"), np.concatenate(foff), cat("face_idx"), cat("lam_mu"),
                         cat("P"), cat("index"), cat("v_length_max"))


def detect_singularities(V_k_coord, coordinates, triangles, eps):
    """All frames of a (n, N, 3) field in batched kernel calls -> Singularities."""
    torch, dev = _torch_dev()
    coords = np.ascontiguousarray(np.asarray(coordinates, dtype=np.float64))
    tri = np.ascontiguousarray(np.asarray(triangles), dtype=np.int32)
    V = np.asarray(V_k_coord, dtype=np.float64)
    if V.ndim == 2:
        V = V[None]
    V = np.ascontiguousarray(V[:, :, :3])
    coords_dev, tri_dev = torch.from_numpy(coords).to(dev), torch.from_numpy(tri).to(dev)
    parts = []
    for k0 in range(0, V.shape[0], FRAMES_PER_CALL):
        Vd = torch.from_numpy(V[k0:k0 + FRAMES_PER_CALL]).to(dev)
        parts.append(detect_singularities_device(coords_dev, tri_dev, Vd, eps))
    return _concat(parts)


def find_singularity_points(coordinates, triangles, V_now, eps):
    """Reference :140-189, same return structure:
    ([[i, coord], ...], [[face_idx, P_coord, triangle, [lam, mu, 1-lam-mu], [A, B, C]], ...], v_length_max)"""
    coordinates = np.asarray(coordinates)
    triangles = np.asarray(triangles)
    s = detect_singularities(np.asarray(V_now, dtype=np.float64)[None], coordinates, triangles, eps)
    return _as_reference_lists(s, 0, coordinates, triangles)


def _as_reference_lists(s, k, coordinates, triangles):
    vi, fi, lm, P, _ = s.frame(k)
    singularity_vertices = [[int(i), coordinates[i]] for i in vi]
    singularity_interiors = []
    for q, t in enumerate(fi):
        tri = triangles[t]
        lam, mu = float(lm[q, 0]), float(lm[q, 1])
        singularity_interiors.append([int(t), P[q].copy(), tri, [lam, mu, 1 - lam - mu],
                                      [coordinates[tri[0]], coordinates[tri[1]], coordinates[tri[2]]]])
    return singularity_vertices, singularity_interiors, float(s.v_length_max[k])


def find_singularity_points_for_all_Vk(V_k_coord, coordinates, triangles, eps):
    """Reference :530-558: per frame, coordinates of singular vertices then interior points."""
    coordinates = np.asarray(coordinates)
    s = detect_singularities(V_k_coord, coordinates, triangles, eps)
    out = []
    for k in range(len(s.v_length_max)):
        vi, fi, lm, P, _ = s.frame(k)
        out.append([coordinates[i] for i in vi] + [P[q].copy() for q in range(len(fi))])
    return out


CLASS_NAMES = ("Node", "Focus", "Saddle", "Indeterminate")      # classify_critical_point, reference :463-498


@dataclasses.dataclass
class Classified:
    """Critical points of a batch of frames with the reference's 2x2 "Jacobian" and class; within a
    frame the order is the reference's: singular vertices first, then interior points."""
    singularities: "Singularities"
    offsets: np.ndarray           # (n+1,) start of each frame in the arrays below
    points: np.ndarray            # (m,3)
    jacobians: np.ndarray         # (m,2,2)
    codes: np.ndarray             # (m,) index into CLASS_NAMES


def mesh_adjacency(triangles, n_vertices):
    """1-ring neighbour lists (CSR, ascending; pyvista point_neighbors, reference :375) and the face
    across each edge AB / BC / CA of every face (-1 on the boundary; the smallest other face if an
    edge is shared by more than two), as int32 arrays."""
    t = np.ascontiguousarray(np.asarray(triangles), dtype=np.int64)
    F = len(t)
    pairs = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]], t[:, [1, 0]], t[:, [2, 1]], t[:, [0, 2]]])
    pairs = np.unique(pairs, axis=0)
    ring_ptr = np.concatenate([[0], np.cumsum(np.bincount(pairs[:, 0], minlength=n_vertices))]).astype(np.int32)
    ring_idx = pairs[:, 1].astype(np.int32)
    a = np.concatenate([t[:, 0], t[:, 1], t[:, 2]])
    b = np.concatenate([t[:, 1], t[:, 2], t[:, 0]])
    key = np.minimum(a, b) * n_vertices + np.maximum(a, b)
    face = np.tile(np.arange(F), 3)
    slot = np.repeat(np.arange(3), F)
    order = np.lexsort((face, key))
    ks, fs = key[order], face[order]
    nbr = -np.ones((F, 3), dtype=np.int32)
    start = np.concatenate([[True], ks[1:] != ks[:-1]])
    group_first = np.maximum.accumulate(np.where(start, np.arange(len(ks)), 0))
    group_size = np.diff(np.concatenate([np.nonzero(start)[0], [len(ks)]]))
    size_of = np.repeat(group_size, group_size)
    first_face = fs[group_first]
    second_face = np.where(size_of > 1, fs[np.minimum(group_first + 1, len(ks) - 1)], -1)
    other = np.where(fs == first_face, second_face, first_face)       # smallest other face of the edge
    other = np.where(size_of > 1, other, -1)
    nbr[fs, slot[order]] = other
    return ring_ptr, ring_idx, nbr


def classify_singularities(V_k_coord, coordinates, triangles, eps, e):
    """Detection (K5) + Jacobian classification (K5b) for all frames -> Classified."""
    torch, dev = _torch_dev()
    lib = _lib.load()
    coords = np.ascontiguousarray(np.asarray(coordinates, dtype=np.float64))
    tri = np.ascontiguousarray(np.asarray(triangles), dtype=np.int32)
    V = np.asarray(V_k_coord, dtype=np.float64)
    if V.ndim == 2:
        V = V[None]
    V = np.ascontiguousarray(V[:, :, :3])
    N, F = len(coords), len(tri)
    e_np = np.ascontiguousarray(np.asarray(e, dtype=np.float64)).reshape(-1, 2, 3)
    ring_ptr, ring_idx, nbr = mesh_adjacency(tri, N)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    coords_d, tri_d, e_d = up(coords), up(tri), up(e_np)
    rp_d, ri_d, nb_d = up(ring_ptr), up(ring_idx), up(nbr)
    st = torch.cuda.current_stream(dev).cuda_stream
    parts, jv, cv, jf, cf = [], [], [], [], []
    for k0 in range(0, V.shape[0], FRAMES_PER_CALL):
        Vd = up(V[k0:k0 + FRAMES_PER_CALL])
        n = int(Vd.shape[0])
        vmax = torch.empty((n,), dtype=torch.float64, device=dev)
        _lib.check(lib.mof_vmax(N, n, Vd.data_ptr(), vmax.data_ptr(), st))
        s = detect_singularities_device(coords_d, tri_d, Vd, eps, vmax)
        nv, nf = len(s.vertex_idx), len(s.face_idx)
        jac_v = torch.empty((max(nv, 1), 4), dtype=torch.float64, device=dev)
        cls_v = torch.empty((max(nv, 1),), dtype=torch.int8, device=dev)
        jac_f = torch.empty((max(nf, 1), 4), dtype=torch.float64, device=dev)
        cls_f = torch.empty((max(nf, 1),), dtype=torch.int8, device=dev)
        voff, foff = up(s.vertex_offsets.astype(np.int64)), up(s.face_offsets.astype(np.int64))
        vi = up(s.vertex_idx) if nv else cls_v
        fi = up(s.face_idx) if nf else cls_f
        Pd = up(s.P) if nf else jac_f
        _lib.check(lib.mof_classify_singularities(
            N, F, n, coords_d.data_ptr(), tri_d.data_ptr(), Vd.data_ptr(), vmax.data_ptr(), e_d.data_ptr(),
            rp_d.data_ptr(), ri_d.data_ptr(), nb_d.data_ptr(), voff.data_ptr(), foff.data_ptr(), nv, nf,
            vi.data_ptr(), fi.data_ptr(), Pd.data_ptr(), jac_v.data_ptr(), cls_v.data_ptr(), jac_f.data_ptr(),
            cls_f.data_ptr(), st))
        parts.append(s)
        jv.append(jac_v[:nv].cpu().numpy()); cv.append(cls_v[:nv].cpu().numpy())
        jf.append(jac_f[:nf].cpu().numpy()); cf.append(cls_f[:nf].cpu().numpy())
    s = _concat(parts)
    jv, cv, jf, cf = np.concatenate(jv), np.concatenate(cv), np.concatenate(jf), np.concatenate(cf)
    pts, jac, codes, offsets = [], [], [], [0]
    for k in range(len(s.v_length_max)):
        v = slice(s.vertex_offsets[k], s.vertex_offsets[k + 1])
        f = slice(s.face_offsets[k], s.face_offsets[k + 1])
        pts += [coords[s.vertex_idx[v]], s.P[f]]
        jac += [jv[v], jf[f]]
        codes += [cv[v], cf[f]]
        offsets.append(offsets[-1] + (v.stop - v.start) + (f.stop - f.start))
    return Classified(s, np.asarray(offsets, dtype=np.int64), np.concatenate(pts).reshape(-1, 3),
                      np.concatenate(jac).reshape(-1, 2, 2), np.concatenate(codes).astype(np.int64))


def find_singularity_points_and_classify_for_all_Vk(V_k_coord, coordinates, triangles, eps, surface, e):
    """Reference :561-605.  ``surface`` is accepted for signature compatibility; the adjacency the
    reference takes from it (point_neighbors, find_cells_intersecting_line) is derived from
    ``triangles``.  -> (singularity_points, classification): per frame a list of coordinates and a
    list of "Node" / "Focus" / "Saddle" / "Indeterminate"."""
    c = classify_singularities(V_k_coord, coordinates, triangles, eps, e)
    pts, cls = [], []
    for k in range(len(c.offsets) - 1):
        sl = slice(c.offsets[k], c.offsets[k + 1])
        pts.append([p.copy() for p in c.points[sl]])
        cls.append([CLASS_NAMES[q] for q in c.codes[sl]])
    return pts, cls


def classify_critical_point(jacobian_matrix):
    """Reference :463-498 for one 2x2 matrix (host-side convenience; the per-point classes of a whole
    recording come from ``classify_singularities`` on the GPU with the same rule)."""
    J = np.asarray(jacobian_matrix, dtype=np.float64).reshape(2, 2)
    trace = J[0, 0] + J[1, 1]
    det = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
    if det > 0:
        return "Node" if trace ** 2 > 4 * det else "Focus"
    return "Saddle" if det < 0 else "Indeterminate"


def analyze_classification(classification):
    """Reference :501-527: prints the Focus / Saddle / Node totals over all frames (like the reference,
    anything that is neither Focus nor Saddle counts as Node); also returns them."""
    counts = {"Focus": 0, "Saddle": 0, "Node": 0}
    for frame in classification:
        for kind in frame:
            counts[kind if kind in ("Focus", "Saddle") else "Node"] += 1
    print(f"Focus: {counts['Focus']}")
    print(f"Saddle: {counts['Saddle']}")
    print(f"Node: {counts['Node']}")
    return counts


__all__ = ["process_V_k", "speed_magnitude", "find_singularity_points", "find_singularity_points_for_all_Vk",
           "classify_critical_point", "analyze_classification",
           "detect_singularities", "detect_singularities_device", "tangent_to_xyz_device", "Singularities",
           "find_singularity_points_and_classify_for_all_Vk", "classify_singularities", "mesh_adjacency", "Classified",
           "CLASS_NAMES"]

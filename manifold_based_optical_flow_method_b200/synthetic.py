"""Synthetic meshes and signals for tests and benchmarks (numpy only, float64).

The reference takes its mesh arrays from pyvista/VTK
(S3_compute_v_and_detection_singularity.py:79-84): ``points``, ``faces``,
``point_normals`` and ``compute_cell_sizes()['Area']``.  pyvista is not
available offline, so these helpers synthesise the same four arrays the way
VTK defines them:

* normals  = normalised sum of the unit normals of the incident faces,
* areas    = 0.5 * |(B - A) x (C - A)|,

for icospheres (N = 10 * 4**L + 2 vertices), a "pial-like" radially
perturbed icosphere, a two-hemisphere mesh and an open patch (mesh with
boundary, as produced by S1_reconstruct_surface.py:85-98).
"""
import numpy as np

__all__ = [
    "icosphere", "pial_like", "two_hemispheres", "open_patch", "fan_mesh",
    "vertex_normals", "face_areas", "travelling_wave", "wrapped_phase",
    "time_axis", "mesh_for_config",
]


def _icosahedron():
    phi = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([
        [-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0],
        [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
        [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1],
    ], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
    ], dtype=np.int64)
    return v, f


def _subdivide(v, f):
    """Loop-style 1->4 split with midpoints pushed back to the unit sphere."""
    n = len(v)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    key = e[:, 0] * n + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    a = uniq // n
    b = uniq % n
    mid = v[a] + v[b]
    mid /= np.linalg.norm(mid, axis=1, keepdims=True)
    nf = len(f)
    m01 = n + inv[:nf]
    m12 = n + inv[nf:2 * nf]
    m20 = n + inv[2 * nf:]
    v2 = np.concatenate([v, mid], axis=0)
    f2 = np.concatenate([
        np.stack([f[:, 0], m01, m20], axis=1),
        np.stack([f[:, 1], m12, m01], axis=1),
        np.stack([f[:, 2], m20, m12], axis=1),
        np.stack([m01, m12, m20], axis=1),
    ], axis=0)
    return v2, f2


def face_areas(coordinates, triangles):
    a = coordinates[triangles[:, 0]]
    b = coordinates[triangles[:, 1]]
    c = coordinates[triangles[:, 2]]
    return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)


def vertex_normals(coordinates, triangles):
    """VTK point_normals convention: normalised sum of unit face normals."""
    a = coordinates[triangles[:, 0]]
    b = coordinates[triangles[:, 1]]
    c = coordinates[triangles[:, 2]]
    fn = np.cross(b - a, c - a)
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    vn = np.zeros_like(coordinates)
    for m in range(3):
        np.add.at(vn, triangles[:, m], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    return vn


def _finish(v, f):
    f = np.ascontiguousarray(f, dtype=np.int64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return v, f, vertex_normals(v, f), face_areas(v, f)


def icosphere(level, radius=1.0):
    """-> (coordinates (N,3), triangles (F,3) int64, normals (N,3), areas (F,))"""
    v, f = _icosahedron()
    for _ in range(level):
        v, f = _subdivide(v, f)
    return _finish(v * radius, f)


def _harmonic_bumps(xhat, seed):
    """Smooth low-order polynomial harmonics on the unit sphere, max |.| = 1."""
    rng = np.random.default_rng(seed)
    x, y, z = xhat[:, 0], xhat[:, 1], xhat[:, 2]
    basis = np.stack([
        x, y, z, x * y, y * z, z * x, x * x - y * y, 3 * z * z - 1,
        x * (x * x - 3 * y * y), y * (3 * x * x - y * y), z * (x * x - y * y), x * y * z,
        z * (5 * z * z - 3), x * (5 * z * z - 1), y * (5 * z * z - 1),
    ], axis=1)
    coef = rng.standard_normal(basis.shape[1])
    s = basis @ coef
    return s / np.max(np.abs(s))


def pial_like(level=7, radius=80.0, amplitude=0.15, seed=0):
    """Icosphere topology (ico7 = 163,842 v / 327,680 f, the size of a
    FreeSurfer ?h.pial) with a smooth seeded radial perturbation."""
    v, f = _icosahedron()
    for _ in range(level):
        v, f = _subdivide(v, f)
    r = radius * (1.0 + amplitude * _harmonic_bumps(v, seed))
    return _finish(v * r[:, None], f)


def two_hemispheres(level=7, radius=80.0, amplitude=0.15, seed=0):
    """Two pial-like components offset in x (config 4: 2 x ico7 = 327,684 v)."""
    v1, f1, _, _ = pial_like(level, radius, amplitude, seed)
    v2, f2, _, _ = pial_like(level, radius, amplitude, seed + 1)
    off = np.array([2.6 * radius, 0.0, 0.0])
    v = np.concatenate([v1 - off / 2, v2 + off / 2], axis=0)
    f = np.concatenate([f1, f2 + len(v1)], axis=0)
    return _finish(v, f)


def open_patch(n=12, seed=0, size=10.0):
    """Open surface with boundary: a jittered, gently curved n x n grid split
    into triangles with alternating diagonals (irregular valence 2..8)."""
    rng = np.random.default_rng(seed)
    g = np.linspace(-size / 2, size / 2, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    h = size / (n - 1)
    X = X + 0.2 * h * rng.uniform(-1, 1, X.shape)
    Y = Y + 0.2 * h * rng.uniform(-1, 1, Y.shape)
    Z = 0.08 * size * np.sin(2.0 * X / size) * np.cos(1.5 * Y / size) + 0.02 * (X * X + Y * Y) / size
    v = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    idx = np.arange(n * n).reshape(n, n)
    tris = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = idx[i, j], idx[i + 1, j], idx[i + 1, j + 1], idx[i, j + 1]
            if (i + j) % 2 == 0:
                tris += [[a, b, c], [a, c, d]]
            else:
                tris += [[a, b, d], [b, c, d]]
    return _finish(v, np.array(tris, dtype=np.int64))


def fan_mesh(n_rim=40, n_rings=3, seed=0):
    """Open cap with one vertex of valence ``n_rim`` (> 32: a block row longer than a warp) surrounded
    by ``n_rings`` rings of ``n_rim`` vertices each; gently curved and jittered."""
    rng = np.random.default_rng(seed)
    pts = [[0.0, 0.0, 0.0]]
    for r in range(1, n_rings + 1):
        ang = 2 * np.pi * (np.arange(n_rim) + 0.5 * (r % 2)) / n_rim
        rad = r * (1.0 + 0.05 * rng.uniform(-1, 1, n_rim))
        pts += [[rad[k] * np.cos(ang[k]), rad[k] * np.sin(ang[k]), -0.05 * rad[k] ** 2] for k in range(n_rim)]
    v = np.array(pts)
    ring = lambda r, k: 1 + (r - 1) * n_rim + (k % n_rim)
    tris = [[0, ring(1, k), ring(1, k + 1)] for k in range(n_rim)]
    for r in range(1, n_rings):
        for k in range(n_rim):
            a, b = ring(r, k), ring(r, k + 1)
            if r % 2 == 1:
                c, d = ring(r + 1, k), ring(r + 1, k + 1)
                tris += [[a, c, b], [b, c, d]]
            else:
                c, d = ring(r + 1, k - 1), ring(r + 1, k)
                tris += [[a, c, d], [a, d, b]]
    return _finish(v, np.array(tris, dtype=np.int64))


def time_axis(n_frames, sampling_frequency=512.0):
    """t_k exactly as S3 builds it (S3...:87): a list of Python floats i/SF."""
    return [i / sampling_frequency for i in range(n_frames)]


def _wave_vectors(coordinates, seed):
    rng = np.random.default_rng(seed + 1000)
    R = float(np.mean(np.linalg.norm(coordinates - coordinates.mean(axis=0), axis=1)))
    d1 = rng.standard_normal(3)
    d1 /= np.linalg.norm(d1)
    d2 = rng.standard_normal(3)
    d2 /= np.linalg.norm(d2)
    return 3.0 / R * d1, 2.2 / R * d2


def travelling_wave(coordinates, t_k, seed=0, noise=1e-3, omega=40.0, omega2=25.0,
                    frame_offset=0):
    """I(x,t) = sin(k.x - w t) + 0.5 cos(2 k'.x + w' t) + noise*N(0,1)  -> (T, N) float64.

    ``frame_offset`` seeds the noise per absolute frame index so that a shard
    of frames is identical to the same rows of the full signal."""
    k1, k2 = _wave_vectors(coordinates, seed)
    t = np.asarray(t_k, dtype=np.float64)[:, None]
    p1 = (coordinates @ k1)[None, :]
    p2 = (coordinates @ k2)[None, :]
    sig = np.sin(p1 - omega * t) + 0.5 * np.cos(2.0 * p2 + omega2 * t)
    if noise:
        for r in range(sig.shape[0]):
            rng = np.random.default_rng([seed, 77, frame_offset + r])
            sig[r] += noise * rng.standard_normal(sig.shape[1])
    return sig


def wrapped_phase(coordinates, t_k, seed=0, omega=40.0):
    """Config-4 style input: I = angle(exp(i(4 k.x - w t))) in (-pi, pi]
    (value range of S2_interpolate_phases.py:52)."""
    k1, _ = _wave_vectors(coordinates, seed)
    t = np.asarray(t_k, dtype=np.float64)[:, None]
    p1 = (coordinates @ k1)[None, :]
    return np.angle(np.exp(1j * (4.0 * p1 - omega * t)))


def mesh_for_config(name):
    """Named workloads of BASELINE.json ``configs``."""
    if name == "C1":
        return icosphere(5)
    if name in ("C2", "C3"):
        return pial_like(7)
    if name == "C4":
        return two_hemispheres(7)
    raise ValueError(f"unknown config {name!r}")


class SurfaceMesh:
    """Minimal stand-in for the pyvista PolyData members the reference scripts touch
    (S3...:79-84, S5_compute_wave_v.py:18-21,162): ``points``, ``faces`` (flat [3,a,b,c,...]),
    ``point_normals``, ``compute_cell_sizes(...)['Area']`` and ``point_cell_ids(i)``."""

    def __init__(self, coordinates, triangles, normals=None, areas=None):
        self.points = np.asarray(coordinates, dtype=np.float64)
        tri = np.asarray(triangles, dtype=np.int64)
        self.triangles = tri
        self.faces = np.concatenate([np.full((len(tri), 1), 3, dtype=np.int64), tri], axis=1).ravel()
        self.point_normals = vertex_normals(self.points, tri) if normals is None else np.asarray(normals)
        self._areas = face_areas(self.points, tri) if areas is None else np.asarray(areas, dtype=np.float64)
        order = np.argsort(tri.ravel(), kind="stable")
        self._cells = order // 3
        self._ptr = np.concatenate([[0], np.cumsum(np.bincount(tri.ravel(), minlength=len(self.points)))])

    def compute_cell_sizes(self, length=False, volume=False):
        return {"Area": self._areas}

    def point_cell_ids(self, index):
        """ids of the cells using point ``index``, ascending (VTK's order for a PolyData)."""
        return [int(c) for c in self._cells[self._ptr[index]:self._ptr[index + 1]]]

    def point_neighbors(self, index):
        """1-ring of point ``index`` (pyvista: points sharing a cell with it), ascending."""
        out = set()
        for c in self.point_cell_ids(index):
            out.update(int(v) for v in self.triangles[c])
        out.discard(int(index))
        return sorted(out)

    def find_closest_point(self, point):
        """index of the vertex closest to ``point`` (lowest index on ties)"""
        return int(np.argmin(np.linalg.norm(self.points - np.asarray(point, dtype=np.float64), axis=1)))

    def point_neighbors_levels(self, ind, n_levels=1):
        """pyvista's generator of topological rings: level 1 = 1-ring, level k = vertices first
        reached after k edges; yields exactly ``n_levels`` lists (empty once the mesh is exhausted)."""
        neighbors = set(self.point_neighbors(ind))
        yield sorted(neighbors)
        visited = set(neighbors)
        visited.add(int(ind))
        for _ in range(n_levels - 1):
            new = set()
            for n in neighbors:
                new.update(self.point_neighbors(n))
            neighbors = new - visited
            yield sorted(neighbors)
            visited |= neighbors

    def find_cells_intersecting_line(self, pointa, pointb, tolerance=0.0):
        """Stand-in for the VTK cell locator as utils/find_singularity_point.py:435 uses it: the
        segment is always a mesh edge there, and the caller wants "the other triangle on that
        edge", so the cells containing BOTH end points are returned (ascending).  (The real
        locator may also return cells that merely touch an end point; which of them
        ``set(...).pop()`` then picks is not defined by the reference.)"""
        a = int(np.argmin(np.linalg.norm(self.points - np.asarray(pointa), axis=1)))
        b = int(np.argmin(np.linalg.norm(self.points - np.asarray(pointb), axis=1)))
        return sorted(set(self.point_cell_ids(a)) & set(self.point_cell_ids(b)))

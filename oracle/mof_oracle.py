"""CPU oracle: numpy/scipy restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
file; the product package never does (it fails loudly without its CUDA library).

Parity status: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), so parity is pinned by *outputs of the reference itself
run in the build container*: ``tests/golden/make_golden.py`` imports the
unmodified reference (oracle/reference_shim.py), runs it on seeded synthetic
inputs and commits the results as ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against those files (and
against the live reference when /root/reference is present).

Third-party arithmetic on the path (not vendored by the reference, no pinned
version): ``scipy.sparse.linalg.spsolve`` -> SuperLU gssv
(utils/compute_optical_flow.py:147) and ``numpy.linalg.lstsq`` -> LAPACK gelsd
(utils/find_singularity_point.py:128).  Pinned here to the versions in this
image: scipy 1.18.1, numpy 2.3.5.  The oracle calls the same two routines.

Every function cites the reference lines it restates.  The restatement is
vectorised over faces/vertices (the literal reference needs ~380 us per face in
pure Python, i.e. ~165 s per frame at 164k vertices) but keeps the reference's
operation order inside each scalar expression.
"""
import math

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve


# ----------------------------------------------------------------------------
# geometry (once per mesh)
# ----------------------------------------------------------------------------
def orthonormal_basis(normals):
    """utils/compute_optical_flow.py:210-235 for all vertices -> e (N,2,3).

    e1 = (-ny, nx, 0) if nx != 0 or ny != 0 else (0, -nz, ny); e2 = n x e1;
    both normalised."""
    n = np.asarray(normals, dtype=np.float64)
    N = len(n)
    e1 = np.zeros((N, 3))
    gen = (n[:, 0] != 0) | (n[:, 1] != 0)
    e1[gen, 0] = -n[gen, 1]
    e1[gen, 1] = n[gen, 0]
    e1[~gen, 1] = -n[~gen, 2]
    e1[~gen, 2] = n[~gen, 1]
    e2 = np.cross(n, e1)
    with np.errstate(invalid="ignore", divide="ignore"):
        e1 = e1 / np.linalg.norm(e1, axis=1, keepdims=True)
        e2 = e2 / np.linalg.norm(e2, axis=1, keepdims=True)
    return np.stack([e1, e2], axis=1)


def _gradient_w(p_i, p_j, p_k):
    """utils/compute_optical_flow.py:238-255 (note: minus the textbook grad w_i)."""
    vjk = p_k - p_j
    vji = p_i - p_j
    with np.errstate(invalid="ignore", divide="ignore"):
        perp = (np.einsum("fi,fi->f", vji, vjk)[:, None] * vjk
                / np.einsum("fi,fi->f", vjk, vjk)[:, None])
        vih = p_j - p_i + perp
        return vih / np.einsum("fi,fi->f", vih, vih)[:, None]


def gradient_w(coordinates, triangles):
    """utils/compute_optical_flow.py:60-68 -> grad_w (F,3,3); argument orders
    (A;B,C), (B;A,C), (C;A,B)."""
    c = np.asarray(coordinates, dtype=np.float64)
    t = np.asarray(triangles)
    A, B, C = c[t[:, 0]], c[t[:, 1]], c[t[:, 2]]
    return np.stack([_gradient_w(A, B, C), _gradient_w(B, A, C), _gradient_w(C, A, B)], axis=1)


def _global_indices(triangles, N):
    """rows/cols of the 36 local entries of each face: index = vertex + N*alpha
    (utils/compute_optical_flow.py:83-84,133-134).  Shape (F,3,2,3,2)."""
    t = np.asarray(triangles, dtype=np.int64)
    al = np.arange(2, dtype=np.int64) * N
    rows = t[:, :, None, None, None] + al[None, None, :, None, None]
    cols = t[:, None, None, :, None] + al[None, None, None, None, :]
    shape = (len(t), 3, 2, 3, 2)
    return np.broadcast_to(rows, shape), np.broadcast_to(cols, shape)


def _to_csr(rows, cols, vals, N):
    m = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(2 * N, 2 * N))
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return m


def a2_matrix(triangles, areas, e, grad_w):
    """utils/compute_optical_flow.py:78-93 + compute_a2 :258-270.
    a2[(i,al),(j,be)] = sum_T (e_i^al . e_j^be) * (g_i . g_j) * A_T, CSR (2N,2N).
    The reference visits only i<=j and mirrors; the mirrored value is the same
    product, so the full local 6x6 block is emitted here."""
    t = np.asarray(triangles, dtype=np.int64)
    N = len(e)
    ef = e[t]                                        # (F,3,2,3)
    ee = np.einsum("fmax,fnbx->fmanb", ef, ef)       # e_i^al . e_j^be
    gg = np.einsum("fmx,fnx->fmn", grad_w, grad_w)   # g_i . g_j
    vals = ee * gg[:, :, None, :, None] * np.asarray(areas, dtype=np.float64)[:, None, None, None, None]
    rows, cols = _global_indices(t, N)
    return _to_csr(rows, cols, vals, N)


def geometrical_quantities(coordinates, normals, triangles, areas):
    """utils/compute_optical_flow.py:27-97 -> (a2 csr, grad_w, e, integral_wi_wj)."""
    areas = np.asarray(areas, dtype=np.float64)
    e = orthonormal_basis(normals)
    gw = gradient_w(coordinates, triangles)
    integral = np.stack([areas / 6, areas / 12], axis=1)    # :73-75
    a2 = a2_matrix(triangles, areas, e, gw)
    return a2, gw, e, integral


# ----------------------------------------------------------------------------
# per-frame assembly and solve
# ----------------------------------------------------------------------------
def assemble_frame(grad_w, e, integral_wi_wj, triangles, areas, dt, I_now, I_next):
    """worker face loop, utils/compute_optical_flow.py:113-141 -> (a1 csr, f).

    grad_M_I = sum_m I[T_m] * grad_w[T][m]                                (:116-117)
    a1[(i,al),(j,be)] += (G.e_i^al) * (G.e_j^be) * (A/6 if i==j else A/12)  (:131-135, :285)
    f[(i,al)] += (e_i^al . G) * (2 d_i + sum_{x in T\\i} d_x) * A / 12      (:123-126, :307-311)
    with d_x = (I_next[x] - I_now[x]) / dt."""
    t = np.asarray(triangles, dtype=np.int64)
    N = len(e)
    I_now = np.asarray(I_now, dtype=np.float64)
    I_next = np.asarray(I_next, dtype=np.float64)
    areas = np.asarray(areas, dtype=np.float64)
    It = I_now[t]                                           # (F,3)
    G = (It[:, 0, None] * grad_w[:, 0] + It[:, 1, None] * grad_w[:, 1]
         + It[:, 2, None] * grad_w[:, 2])                   # (F,3)
    c = np.einsum("fx,fmax->fma", G, e[t])                  # (F,3,2)
    same = np.eye(3, dtype=bool)[None]
    integ = np.where(same, integral_wi_wj[:, 0, None, None], integral_wi_wj[:, 1, None, None])
    vals = c[:, :, :, None, None] * c[:, None, None, :, :] * integ[:, :, None, :, None]
    rows, cols = _global_indices(t, N)
    a1 = _to_csr(rows, cols, vals, N)
    d = (I_next - I_now) / dt                               # (N,)
    dT = d[t]                                               # (F,3)
    others = np.stack([dT[:, 1] + dT[:, 2], dT[:, 0] + dT[:, 2], dT[:, 0] + dT[:, 1]], axis=1)   # :308-309
    w = 2 * dT + others                                     # (F,3)
    fl = c * w[:, :, None] * areas[:, None, None] / 12      # (F,3,2)
    f = np.zeros(2 * N)
    al = np.arange(2, dtype=np.int64) * N
    np.add.at(f, (t[:, :, None] + al[None, None, :]).ravel(), fl.ravel())
    return a1, f


def system_matrix(a1, a2, lambda_):
    """utils/compute_optical_flow.py:144-146: a = csr(a1 + lambda_*a2)."""
    return sp.csr_matrix(a1 + lambda_ * sp.csr_matrix(a2))


def worker(k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k_k, I_k_kplus1):
    """utils/compute_optical_flow.py:100-149 (same signature, no print)."""
    dt = t_k[k + 1] - t_k[k]                                # :125
    a1, f = assemble_frame(grad_w, e, integral_wi_wj, triangles, areas, dt, I_k_k, I_k_kplus1)
    a = system_matrix(a1, a2, lambda_)
    return spsolve(a, f)                                    # :147 (SuperLU gssv)


def compute_velocity_field(processes_num, time_steps, a2, grad_w, e, integral_wi_wj, triangles,
                           t_k, areas, lambda_, I_k, I_k_2):
    """utils/compute_optical_flow.py:152-194: frame k uses (I_k[k], I_k_2[k+1]).
    processes_num <= 1 runs serially; otherwise a multiprocessing Pool like the
    reference (:157-179)."""
    import time
    start = time.time()
    a2 = sp.csr_matrix(a2)
    args = [(k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k[k], I_k_2[k + 1])
            for k in range(time_steps - 1)]
    if processes_num and processes_num > 1 and len(args) > 1:
        import multiprocessing
        ctx = multiprocessing.get_context("fork")
        with ctx.Pool(processes_num) as pool:
            V_k = pool.starmap(worker, args, chunksize=1)
    else:
        V_k = [worker(*a) for a in args]
    return V_k, time.time() - start


# ----------------------------------------------------------------------------
# tangent coefficients -> xyz, singularity detection
# ----------------------------------------------------------------------------
def process_V_k(V_k, e):
    """utils/find_singularity_point.py:28-69 -> array (T-1, N, 3):
    V_xyz[k,i] = V_k[k][i]*e[i,0] + V_k[k][i+N]*e[i,1]."""
    V = np.asarray(V_k, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    N = len(e)
    return V[:, :N, None] * e[None, :, 0, :] + V[:, N:, None] * e[None, :, 1, :]


def speed_magnitude(V_k_coord):
    """S3_compute_v_and_detection_singularity.py:130-132."""
    V = np.asarray(V_k_coord)
    return np.sqrt(np.sum(V[:, :, :3] ** 2, axis=2))


def _face_plane_terms(coordinates, triangles, Vn):
    t = np.asarray(triangles, dtype=np.int64)
    A, B, C = coordinates[t[:, 0]], coordinates[t[:, 1]], coordinates[t[:, 2]]
    n = np.cross(B - A, C - A)                                       # fsp:113
    with np.errstate(invalid="ignore", divide="ignore"):
        n = n / np.linalg.norm(n, axis=1, keepdims=True)             # fsp:114
    VA, VB, VC = Vn[t[:, 0]], Vn[t[:, 1]], Vn[t[:, 2]]
    VAp = VA - np.einsum("fi,fi->f", VA, n)[:, None] * n             # fsp:117-119
    VBp = VB - np.einsum("fi,fi->f", VB, n)[:, None] * n
    VCp = VC - np.einsum("fi,fi->f", VC, n)[:, None] * n
    return A, B, C, n, VAp, VBp, VCp


def find_singularity_points(coordinates, triangles, V_now, eps, margin=1e-6):
    """utils/find_singularity_point.py:140-189.

    -> (vertex_idx (nv,), face_idx (nf,), lam_mu (nf,2), P_coord (nf,3), v_length_max)

    The accept decision for a face is the reference's own: np.linalg.lstsq on
    M = [VA_p - VC_p | VB_p - VC_p], rhs -VC_p (:122-130).  To avoid F Python-level
    lstsq calls, faces whose closed-form (normal-equation) solution is outside the
    triangle by more than ``margin`` *and* whose 2x2 Gram matrix is well conditioned
    are rejected without calling lstsq; every other face goes through lstsq."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    V_now = np.asarray(V_now, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64)
    V_length = np.sqrt(V_now[:, 0] ** 2 + V_now[:, 1] ** 2 + V_now[:, 2] ** 2)   # :161
    vmax = np.max(V_length)                                                         # :162
    Vn = V_now / vmax
    vflag = np.linalg.norm(Vn, axis=1) <= eps                                       # :86,:166
    vertex_idx = np.nonzero(vflag)[0]
    skip = vflag[t].any(axis=1)                                                     # :171
    A, B, C, n, VAp, VBp, VCp = _face_plane_terms(coordinates, t, Vn)
    M0 = VAp - VCp
    M1 = VBp - VCp
    g00 = np.einsum("fi,fi->f", M0, M0)
    g01 = np.einsum("fi,fi->f", M0, M1)
    g11 = np.einsum("fi,fi->f", M1, M1)
    r0 = -np.einsum("fi,fi->f", M0, VCp)
    r1 = -np.einsum("fi,fi->f", M1, VCp)
    det = g00 * g11 - g01 * g01
    with np.errstate(invalid="ignore", divide="ignore"):
        lam = (r0 * g11 - r1 * g01) / det
        mu = (g00 * r1 - g01 * r0) / det
        well = det > 1e-8 * g00 * g11
    clearly_out = well & ((lam < -margin) | (mu < -margin) | (lam + mu > 1 + margin))
    cand = np.nonzero(~skip & ~clearly_out)[0]
    face_idx, lam_mu = [], []
    for fidx in cand:
        Mf = np.column_stack((M0[fidx], M1[fidx]))                                  # :122
        try:
            l, m = np.linalg.lstsq(Mf, -VCp[fidx], rcond=None)[0]                   # :128
        except np.linalg.LinAlgError:
            continue
        if l + m <= 1 and l >= 0 and m >= 0:                                        # :130
            face_idx.append(fidx)
            lam_mu.append((l, m))
    face_idx = np.asarray(face_idx, dtype=np.int64)
    lam_mu = np.asarray(lam_mu, dtype=np.float64).reshape(-1, 2)
    if len(face_idx):
        l, m = lam_mu[:, 0:1], lam_mu[:, 1:2]
        P = l * A[face_idx] + m * B[face_idx] + (1 - l - m) * C[face_idx]           # :181-182
    else:
        P = np.zeros((0, 3))
    return vertex_idx, face_idx, lam_mu, P, vmax


def face_poincare_index(coordinates, triangles, V_now, face_idx):
    """Per-face Poincare index: winding number of the plane-projected vertex
    velocities VA_p -> VB_p -> VC_p -> VA_p measured in the face plane oriented by
    the face winding (B-A)x(C-A).  Angle formula of S7_winding_line.py:59-87
    (ordered angle v1->v2, counter-clockwise positive; sum / 2pi).  -> int (+1/-1/0)."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    V_now = np.asarray(V_now, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64)[np.asarray(face_idx, dtype=np.int64)]
    A, B, C, n, VAp, VBp, VCp = _face_plane_terms(coordinates, t, V_now)
    u = B - A
    u = u / np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(n, u)
    out = np.zeros(len(t), dtype=np.int64)
    for q in range(len(t)):
        vs = [np.array([np.dot(v[q], u[q]), np.dot(v[q], w[q])]) for v in (VAp, VBp, VCp)]
        tot = 0.0
        for a in range(3):
            v1 = vs[a] / np.linalg.norm(vs[a])
            v2 = vs[(a + 1) % 3] / np.linalg.norm(vs[(a + 1) % 3])
            ang = np.arccos(np.clip(np.dot(v1, v2), -1.0, 1.0))
            if v1[0] * v2[1] - v1[1] * v2[0] < 0:
                ang = -ang
            tot += ang
        out[q] = int(np.rint(tot / (2 * np.pi)))
    return out


def find_singularity_points_for_all_Vk(V_k_coord, coordinates, triangles, eps):
    """utils/find_singularity_point.py:530-558 -> per frame list of point
    coordinates, singular vertices first, then interior points."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    out = []
    for V_now in V_k_coord:
        vi, fi, lm, P, _ = find_singularity_points(coordinates, triangles, V_now, eps)
        out.append([coordinates[i] for i in vi] + [p for p in P])
    return out


# ----------------------------------------------------------------------------
# S5_compute_wave_v.py ("next" row 1)
# ----------------------------------------------------------------------------
def angle_subtract(f1, f2):
    """S5_compute_wave_v.py:224-233."""
    return np.mod(f1 - f2 + np.pi, 2 * np.pi) - np.pi


def temporal_gradient_phase(data, dt):
    """S5_compute_wave_v.py:60-77."""
    data = np.asarray(data, dtype=np.float64)
    g = np.zeros_like(data)
    g[0] = angle_subtract(data[1], data[0]) / dt
    g[1:-1] = angle_subtract(data[2:], data[:-2]) / (2 * dt)
    g[-1] = angle_subtract(data[-1], data[-2]) / dt
    return g


def grad_M_I(coordinates, triangles, potentials, areas):
    """S5_compute_wave_v.py:136-171 -> grad_point (T,N,3): area-weighted mean over the faces of a
    vertex (ascending face index) of the per-face gradient sum_m I[T_m] grad_w[T][m]."""
    t = np.asarray(triangles, dtype=np.int64)
    pot = np.asarray(potentials, dtype=np.float64)
    areas = np.asarray(areas, dtype=np.float64)
    gw = gradient_w(coordinates, t)
    T, N = pot.shape
    gM = (pot[:, t[:, 0], None] * gw[None, :, 0] + pot[:, t[:, 1], None] * gw[None, :, 1]
          + pot[:, t[:, 2], None] * gw[None, :, 2])                       # (T,F,3)  :154-158
    gp = np.zeros((T, N, 3))
    asum = np.zeros(N)
    for f in range(len(t)):                                                # ascending faces = :161-166 order
        for m in range(3):
            gp[:, t[f, m]] += gM[:, f] * areas[f]
            asum[t[f, m]] += areas[f]
    return gp / asum[None, :, None]


def wave_velocity(coordinates, triangles, areas, data, dt, e, phase=True):
    """S5_compute_wave_v.py:79-123 (phase) / :14-58 (amplitude)."""
    data = np.asarray(data, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    td = temporal_gradient_phase(data, dt) if phase else np.gradient(data, axis=0, edge_order=2) / dt
    gp = grad_M_I(coordinates, triangles, data, areas)
    e1, e2 = e[:, 0], e[:, 1]
    n = np.cross(e1, e2)                                                   # :177
    Vn = np.einsum("tnx,nx->tn", gp, n)[:, :, None] * n[None] / np.einsum("nx,nx->n", n, n)[None, :, None]
    Vt = gp - Vn                                                           # :179
    al = np.einsum("tnx,nx->tn", Vt, e1) / np.einsum("nx,nx->n", e1, e1)   # :189
    be = np.einsum("tnx,nx->tn", Vt, e2) / np.einsum("nx,nx->n", e2, e2)   # :190
    return td / np.sqrt(al ** 2 + be ** 2)                                 # :117,:121


# ----------------------------------------------------------------------------
# Jacobian classification of critical points (utils/find_singularity_point.py:355-498,561-605;
# "next" row 2 of SURVEY 8f).  pyvista is replaced by explicit adjacency: point_neighbors = the
# 1-ring (ascending), and "the other triangle on the nearest edge" for find_cells_intersecting_line
# followed by set.pop() (fsp:434-437), which the reference does not define more precisely.
# ----------------------------------------------------------------------------
CLASS_NAMES = ("Node", "Focus", "Saddle", "Indeterminate")


def one_ring(triangles, N):
    """ascending 1-ring neighbour lists as CSR (ptr, idx)"""
    t = np.asarray(triangles, dtype=np.int64)
    pairs = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]], t[:, [1, 0]], t[:, [2, 1]], t[:, [0, 2]]])
    pairs = np.unique(pairs, axis=0)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(pairs[:, 0], minlength=N))])
    return ptr, pairs[:, 1].copy()


def face_neighbors(triangles):
    """(F,3): face across edge k of each face (edge 0 = AB, 1 = BC, 2 = CA), -1 on the boundary;
    if more than two faces share an edge the smallest other face index is taken."""
    t = np.asarray(triangles, dtype=np.int64)
    F = len(t)
    edges = {}
    for f in range(F):
        for k in range(3):
            key = tuple(sorted((int(t[f, k]), int(t[f, (k + 1) % 3]))))
            edges.setdefault(key, []).append(f)
    out = -np.ones((F, 3), dtype=np.int64)
    for f in range(F):
        for k in range(3):
            key = tuple(sorted((int(t[f, k]), int(t[f, (k + 1) % 3]))))
            others = [g for g in edges[key] if g != f]
            if others:
                out[f, k] = min(others)
    return out


def _jacobian_sum(origin, near, coordinates, V_now, vmax, e1, e2):
    """fsp:383-399 / :442-458: J += [[u/d1, u/d2], [v/d1, v/d2]] over the near points."""
    J = np.zeros((2, 2))
    n = np.cross(e1, e2)
    for nb in near:
        Vn = V_now[nb] / vmax
        Vp = Vn - np.dot(Vn, n) * n / np.dot(n, n)                       # project_vector_to_plane, fsp:206-210
        u = np.dot(Vp, e1) / np.dot(e1, e1)                              # express_vector_on_basis, fsp:266-267
        v = np.dot(Vp, e2) / np.dot(e2, e2)
        Br = coordinates[nb] - origin                                    # position_diff_on_basis_with_origin, fsp:231-239
        pr = Br - np.dot(Br, n) * n / np.dot(n, n)
        d1, d2 = np.dot(pr, e1), np.dot(pr, e2)
        with np.errstate(divide="ignore", invalid="ignore"):
            J += np.array([[u / d1, u / d2], [v / d1, v / d2]])
    return J


def jacobian_for_vertex(index, coordinates, V_now, vmax, e, ring_ptr, ring_idx):
    """compute_jacobian_matrix_for_vertex, fsp:355-402."""
    near = ring_idx[ring_ptr[index]:ring_ptr[index + 1]]
    return _jacobian_sum(coordinates[index], near, coordinates, V_now, vmax, e[index][0], e[index][1])


def nearest_edge_index(A, B, C, P):
    """find_nearest_edge_and_vertices, fsp:318-351, including its flat argmin over the three
    3-vectors |cross(P - X, v)| / |v| (index 0 -> AB, 1 -> BC, anything else -> CA)."""
    v1, v2, v3 = B - A, C - B, A - C
    d = np.array([np.abs(np.cross(P - A, v1) / np.linalg.norm(v1)), np.abs(np.cross(P - B, v2) / np.linalg.norm(v2)),
                  np.abs(np.cross(P - C, v3) / np.linalg.norm(v3))])
    k = int(np.argmin(d))
    return 0 if k == 0 else (1 if k == 1 else 2)


def jacobian_for_interior(face, P, coordinates, triangles, V_now, vmax, fnbr):
    """compute_jacobian_matrix_for_interior, fsp:405-460."""
    tri = np.asarray(triangles[face], dtype=np.int64)
    A, B, C = coordinates[tri[0]], coordinates[tri[1]], coordinates[tri[2]]
    normal = np.cross(B - A, C - A)
    normal = normal / np.linalg.norm(normal)                              # calculate_normal, fsp:285-289
    e1, e2 = orthonormal_basis(normal[None])[0]                           # compute_orthonormal_basis, fsp:304-314
    k = nearest_edge_index(A, B, C, P)
    other = fnbr[face, k]
    near = set(int(v) for v in tri)
    if other >= 0:
        near |= set(int(v) for v in triangles[other])
    return _jacobian_sum(P, sorted(near), coordinates, V_now, vmax, e1, e2)


def classify(J):
    """classify_critical_point, fsp:463-498 -> index into CLASS_NAMES."""
    trace = J[0, 0] + J[1, 1]
    det = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
    if det > 0:
        return 0 if trace ** 2 > 4 * det else 1
    if det < 0:
        return 2
    return 3


def classify_singularities(coordinates, triangles, V_now, eps, e):
    """One frame of find_singularity_points_and_classify_for_all_Vk, fsp:582-603
    -> (points (n,3), class codes (n,), jacobians (n,2,2)); vertices first, then interiors."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    V_now = np.asarray(V_now, dtype=np.float64)
    vi, fi, lm, P, vmax = find_singularity_points(coordinates, triangles, V_now, eps)
    ptr, idx = one_ring(triangles, len(coordinates))
    fnbr = face_neighbors(triangles)
    Js = [jacobian_for_vertex(int(i), coordinates, V_now, vmax, e, ptr, idx) for i in vi]
    Js += [jacobian_for_interior(int(f), P[q], coordinates, triangles, V_now, vmax, fnbr) for q, f in enumerate(fi)]
    pts = [coordinates[i] for i in vi] + [p for p in P]
    return (np.asarray(pts).reshape(-1, 3), np.asarray([classify(J) for J in Js], dtype=np.int64),
            np.asarray(Js).reshape(-1, 2, 2))


# ----------------------------------------------------------------------------
# Multi-ring winding numbers of a singular point (S7_winding_line.py:59-165; "next" row 3 of
# SURVEY 8f).  pyvista is replaced by explicit adjacency: find_closest_point = the nearest vertex
# (lowest index on ties), point_neighbors_levels = breadth-first topological rings.
# ----------------------------------------------------------------------------
def closest_vertex(coordinates, point):
    """surf.find_closest_point, S7:130."""
    d = np.linalg.norm(np.asarray(coordinates, dtype=np.float64) - np.asarray(point, dtype=np.float64), axis=1)
    return int(np.argmin(d))


def ring_levels(ring_ptr, ring_idx, index, max_level):
    """surf.point_neighbors_levels(index, max_level), S7:131: ring k+1 = unvisited neighbours of ring k
    (ascending); stops at the first empty ring."""
    seen = {int(index)}
    cur = [int(index)]
    out = []
    for _ in range(max_level):
        nxt = set()
        for v in cur:
            nxt.update(int(w) for w in ring_idx[ring_ptr[v]:ring_ptr[v + 1]])
        nxt -= seen
        if not nxt:
            break
        cur = sorted(nxt)
        seen |= nxt
        out.append(cur)
    return out


def signed_angle(v1, v2):
    """angle_between_vectors, S7:59-74 (2-vectors; counter-clockwise positive)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.asarray(v1, dtype=np.float64) / np.sqrt(v1[0] * v1[0] + v1[1] * v1[1])
        b = np.asarray(v2, dtype=np.float64) / np.sqrt(v2[0] * v2[0] + v2[1] * v2[1])
        d = a[0] * b[0] + a[1] * b[1]
        if d > 1:
            d = 1.0
        elif d < -1:
            d = -1.0
        ang = np.arccos(d)
    if a[0] * b[1] - a[1] * b[0] < 0:
        ang = -ang
    return ang


def ring_winding_number(index, ring, coordinates, V_now, e):
    """One level of S7:135-149: ring vertices and their velocities expressed in the tangent basis
    of vertex `index`, ordered by polar angle (stable, S7:93-100), summed turning / 2 pi (S7:76-87)."""
    e1, e2 = e[index][0], e[index][1]
    n = np.cross(e1, e2)
    keys, vx, vy = [], [], []
    for w in ring:
        Br = coordinates[w] - coordinates[index]                           # S7:36-45
        pr = Br - np.dot(Br, n) * n / np.dot(n, n)
        keys.append(math.atan2(np.dot(pr, e2), np.dot(pr, e1)))
        Vt = V_now[w] - np.dot(V_now[w], n) * n / np.dot(n, n)             # S7:26-33
        vx.append(np.dot(Vt, e1) / np.dot(e1, e1))                         # S7:48-57
        vy.append(np.dot(Vt, e2) / np.dot(e2, e2))
    order = np.argsort(np.asarray(keys), kind="stable")
    total = 0.0
    m = len(ring)
    for i in range(m):
        a, b = order[i], order[(i + 1) % m]
        total += signed_angle((vx[a], vy[a]), (vx[b], vy[b]))
    return total / (2 * np.pi)


def winding_numbers(coordinates, triangles, singularity_points, V_now, e, max_level=25, adjacency=None):
    """calculate_winding_numbers, S7:120-165 -> (counts (n,), types (n,) in {-1,0,1}, winding (n,max_level)
    with NaN where a level was not evaluated).  `types` has one entry per point (0 where the first
    ring is neither +1 nor -1); the reference's list keeps only the non-zero ones, in order.
    Where the mesh runs out of rings before max_level the reference raises IndexError; here the
    count simply stops."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    V_now = np.asarray(V_now, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64).reshape(len(coordinates), 2, 3)
    ptr, idx = adjacency if adjacency is not None else one_ring(triangles, len(coordinates))
    pts = np.asarray(singularity_points, dtype=np.float64).reshape(-1, 3)
    counts = np.zeros(len(pts), dtype=np.int64)
    types = np.zeros(len(pts), dtype=np.int64)
    wind = np.full((len(pts), max_level), np.nan)
    for q, P in enumerate(pts):
        index = closest_vertex(coordinates, P)
        rings = ring_levels(ptr, idx, index, max_level)
        flag = 0
        for level, ring in enumerate(rings):
            w = ring_winding_number(index, ring, coordinates, V_now, e)
            wind[q, level] = w
            if level == 0:
                if -1.01 <= w <= -0.99:
                    flag = -1
                elif 0.99 <= w <= 1.01:
                    flag = 1
                else:
                    break                                                   # S7:160 with flag 0: check_property -> None
                counts[q] += 1
            elif (flag == 1 and 0.999 <= w <= 1.001) or (flag == -1 and -1.001 <= w <= -0.999):
                counts[q] += 1
            else:
                break
        types[q] = flag
    return counts, types, wind


# ----------------------------------------------------------------------------
# RBF interpolation of electrode signals onto the surface (S2_interpolate.py:22-53,
# S2_interpolate_phases.py:22-68; "next" row 4 of SURVEY 8f, the producer of the hot path's
# input).  The arithmetic lives in scipy.interpolate.Rbf (legacy class, scipy 1.18.1 in this
# image, not vendored by the reference) with its defaults: multiquadric phi(r) =
# sqrt((r/eps)^2 + 1), eps = (prod(bounding-box edges) / m)^(1/n_edges), smooth = 0,
# weights = solve(phi(pairwise distances), data), value = phi(distances to the centres) . weights.
# ----------------------------------------------------------------------------
def rbf_epsilon(centres):
    """scipy/interpolate/_rbf.py Rbf.__init__: default epsilon."""
    c = np.asarray(centres, dtype=np.float64)
    edges = c.max(axis=0) - c.min(axis=0)
    edges = edges[np.nonzero(edges)]
    return float(np.power(np.prod(edges) / len(c), 1.0 / edges.size))


def rbf_matrix(points, centres, eps):
    d = points[:, None, :] - centres[None, :, :]
    r = np.sqrt((d * d).sum(axis=2))
    return np.sqrt((1.0 / eps * r) ** 2 + 1)


def rbf_interpolate(centres, data, vertices, phase=False):
    """S2:37-47 for all frames at once: data (T, m) real or complex -> (T, N); phase=True
    returns np.angle of the (complex) interpolant like S2_interpolate_phases.py:51-52."""
    c = np.asarray(centres, dtype=np.float64)
    v = np.asarray(vertices, dtype=np.float64)
    eps = rbf_epsilon(c)
    A = rbf_matrix(c, c, eps)
    W = np.linalg.solve(A, np.asarray(data).T)                              # (m, T)
    out = (rbf_matrix(v, c, eps) @ W).T
    return np.angle(out) if phase else out


def electrode_phases(potentials):
    """compute_phase_from_potentials, S2_interpolate_phases.py:58-68: scipy.signal.hilbert along the
    LAST axis (the electrode axis of a (T, m) array, as the reference calls it), then the angle."""
    x = np.asarray(potentials, dtype=np.float64)
    n = x.shape[-1]
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    return np.angle(np.fft.ifft(np.fft.fft(x, axis=-1) * h, axis=-1))

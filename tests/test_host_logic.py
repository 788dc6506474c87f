"""CPU-only tests (no GPU): the C-ABI library loads and exports every declared symbol, the
host-side pattern builder is correct, and the kernel bodies of csrc/mof_bodies.h -- run
through the test-only loop harness tests/hostcheck -- reproduce the reference's golden
outputs with the same frame-minor layout the CUDA kernels use."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, load_golden, rel_l2
from manifold_based_optical_flow_method_b200 import _lib, synthetic
from manifold_based_optical_flow_method_b200.mesh import Pattern
from manifold_based_optical_flow_method_b200.solver import frame_dt

import hostcheck

W = _lib.GROUP


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "mof_b200.h")).read()
    declared = set(re.findall(r"\b(mof_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mof_version() >= 100
    assert lib.mof_num_tiles(129) == 3
    assert lib.mof_state_ints(2) == 2 * 4 * 32 + 2 * 2 + 2


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.MeshDev) == 4 * 8 + 12 * 8 + 4 + 17 * 4 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(_lib.BatchDev) == 8 + 15 * 8
    assert ctypes.sizeof(_lib.PcgProfile) == 8 * 8 + 8 + 8 + 4 * 8


@pytest.mark.parametrize("reorder", [False, True])
@pytest.mark.parametrize("mesh", ["ico2", "patch", "two"])
def test_pattern(mesh, reorder):
    if mesh == "ico2":
        coords, tris, _, _ = synthetic.icosphere(2)
    elif mesh == "patch":
        coords, tris, _, _ = synthetic.open_patch(9, seed=1)
    else:
        coords, tris, _, _ = synthetic.two_hemispheres(2)
    N, F = len(coords), len(tris)
    P = Pattern(N, tris, reorder=reorder)
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    e.sort(axis=1)
    E = len(np.unique(e, axis=0))
    assert P.n_blocks == N + 2 * E and P.n_contrib == 9 * F
    assert np.array_equal(np.sort(P.perm), np.arange(N))
    if not reorder:
        assert np.array_equal(P.perm, np.arange(N))
    assert np.array_equal(P.perm[P.tri], tris)
    rows = P.block_rows()
    for v in range(N):
        c = P.col[P.rowptr[v]:P.rowptr[v + 1]]
        assert np.all(np.diff(c) > 0)
        assert P.col[P.diag[v]] == v
    # adjacency equals mesh edges
    pairs = set(map(tuple, np.stack([P.perm[rows], P.perm[P.col]], axis=1)))
    want = {(i, i) for i in range(N)} | {(a, b) for a, b in e} | {(b, a) for a, b in e}
    assert pairs == want
    # contributor lists: ascending faces, (m, n) are the positions of (row, col) in the face
    for b in range(P.n_blocks):
        ent = P.centry[P.cptr[b]:P.cptr[b + 1]]
        f, m, n = ent >> 4, (ent >> 2) & 3, ent & 3
        assert np.all(np.diff(f) > 0) and len(f) >= 1
        assert np.all(P.tri[f, m] == rows[b]) and np.all(P.tri[f, n] == P.col[b])
    assert np.sum(np.diff(P.cptr)) == 9 * F
    assert P.max_row_blocks == np.max(np.diff(P.rowptr))


def test_pattern_reorder_reduces_bandwidth():
    coords, tris, _, _ = synthetic.icosphere(4)
    p0 = Pattern(len(coords), tris, reorder=False)
    p1 = Pattern(len(coords), tris, reorder=True)
    assert p1.bandwidth < p0.bandwidth / 4


def test_pattern_rejects_bad_faces():
    with pytest.raises(_lib.MofError):
        Pattern(4, np.array([[0, 1, 4]]))
    with pytest.raises(_lib.MofError):
        Pattern(4, np.array([[0, 1, 1]]))
    with pytest.raises(ValueError):
        Pattern(4, np.array([[0, 1, 2, 3]]))


def test_block_csr_roundtrip():
    coords, tris, _, _ = synthetic.icosphere(1)
    P = Pattern(len(coords), tris)
    vals = np.random.default_rng(0).standard_normal((P.n_blocks, 4))
    m = P.block_values_to_csr(vals)
    assert m.shape == (2 * len(coords),) * 2 and m.nnz == 4 * P.n_blocks
    assert np.array_equal(P.csr_to_block_values(m), vals)


# ---------------------------------------------------------------------------------
# kernel bodies through the loop harness
# ---------------------------------------------------------------------------------
class HostMesh:
    """numpy-backed mof_mesh_dev for the harness"""

    def __init__(self, g, reorder=True):
        self.hc = hostcheck.load()
        self.coords, self.tris = g["coordinates"], g["triangles"]
        N = len(self.coords)
        self.P = P = Pattern(N, self.tris, reorder=reorder, coordinates=self.coords)
        self.areas = np.ascontiguousarray(g["areas"], dtype=np.float64)
        self.e = np.zeros((N, 2, 3))
        self.grad_w = np.zeros((P.n_faces, 3, 3))
        self.integral = np.zeros((P.n_faces, 2))
        self.a2v = np.zeros((P.n_blocks, 4))
        self.cint = np.ascontiguousarray(self.coords[P.perm])
        self.nint = np.ascontiguousarray(g["normals"][P.perm])
        ms = self.struct()
        self.hc.hc_geom(ctypes.byref(ms), self.cint.ctypes.data, self.nint.ctypes.data, self.e.ctypes.data,
                        self.grad_w.ctypes.data, self.integral.ctypes.data)
        self.hc.hc_a2(ctypes.byref(ms), self.a2v.ctypes.data)

    def struct(self):
        P = self.P
        p = lambda a: a.ctypes.data
        return _lib.MeshDev(P.n_vertices, P.n_faces, P.n_blocks, P.n_contrib, p(P.perm), p(P.rowptr), p(P.col), p(P.diag),
                            p(P.cptr), p(P.centry), p(P.tri), p(self.e), p(self.grad_w), p(self.integral), p(self.areas),
                            p(self.a2v), 0, (ctypes.c_int32 * (_lib.MAX_COLORS + 1))(), 0, 0, None)

    def assemble(self, I, t_k, lambda_, omega=0.0):
        P, hc = self.P, self.hc
        n = len(I) - 1
        G = -(-n // W)
        N, nb = P.n_vertices, P.n_blocks
        dt = frame_dt(list(t_k), 0, n)
        I = np.ascontiguousarray(I, dtype=np.float64)
        It, dIt = np.zeros((G, N, W)), np.zeros((G, N, W))
        vals, rhs, minv = np.zeros((G, nb, 4, W)), np.zeros((G, N, 2, W)), np.zeros((G, N, 3, W))
        ms = self.struct()
        hc.hc_pack(ctypes.byref(ms), G, n, I.ctypes.data, I[1:].ctypes.data, ctypes.c_int64(I.shape[1]), dt.ctypes.data,
                   It.ctypes.data, dIt.ctypes.data)
        hc.hc_assemble(ctypes.byref(ms), G, It.ctypes.data, dIt.ctypes.data, ctypes.c_double(lambda_), ctypes.c_double(omega),
                       vals.ctypes.data, rhs.ctypes.data, minv.ctypes.data)
        return vals, rhs, minv


def _frame_matrix(P, vals, k):
    return P.block_values_to_csr(vals[k // W, :, :, k % W])


def _frame_vector(P, vec, k):
    """[G][N][2][W] internal -> (2N,) reference order"""
    N = P.n_vertices
    out = np.empty(2 * N)
    out[P.perm] = vec[k // W, :, 0, k % W]
    out[P.perm + N] = vec[k // W, :, 1, k % W]
    return out


@pytest.mark.parametrize("reorder", [0, 1, 2])
def test_bodies_geometry_and_assembly_match_reference(golden, reorder):
    g = golden
    hm = HostMesh(g, reorder)
    P = hm.P
    N = P.n_vertices
    e_ref = np.empty_like(hm.e)
    e_ref[:] = g["e"][P.perm]
    assert np.max(np.abs(hm.e - e_ref)) <= 1e-15
    assert rel_l2(hm.grad_w, g["grad_w"]) <= 1e-15
    assert np.array_equal(hm.integral, g["integral_wi_wj"])
    if "a2_data" in g:
        a2_ref = sp.csr_matrix((g["a2_data"], g["a2_indices"], g["a2_indptr"]), shape=(2 * N, 2 * N))
        a2 = P.block_values_to_csr(hm.a2v)
        assert abs(a2 - a2_ref).max() <= 1e-14 * abs(a2_ref).max()
    vals, rhs, minv = hm.assemble(g["I"], g["t_k"], float(g["lambda_"]))
    if "a0_data" in g:
        a_ref = sp.csr_matrix((g["a0_data"], g["a0_indices"], g["a0_indptr"]), shape=(2 * N, 2 * N))
        a = _frame_matrix(P, vals, 0)
        assert abs(a - a_ref).max() <= 1e-14 * abs(a_ref).max()
        assert abs(a - a.T).max() == 0.0            # exactly symmetric, like the reference's mirror assignment
        assert rel_l2(_frame_vector(P, rhs, 0), g["f0"]) <= 1e-14
    # solving the harness-assembled system reproduces the reference's velocity field
    from scipy.sparse.linalg import spsolve
    for k in range(len(g["V_k"])):
        V = spsolve(sp.csc_matrix(_frame_matrix(P, vals, k)), _frame_vector(P, rhs, k))
        assert rel_l2(V, g["V_k"][k]) <= (1e-9 if "phase" in g["name"] else 1e-11)
    # block-Jacobi inverse
    k = 0
    a = _frame_matrix(P, vals, k).toarray()
    for v in (0, N // 2, N - 1):
        o = P.perm[v]
        D = a[np.ix_([o, o + N], [o, o + N])]
        Mi = np.array([[minv[0, v, 0, 0], minv[0, v, 1, 0]], [minv[0, v, 1, 0], minv[0, v, 2, 0]]])
        assert np.allclose(Mi @ D, np.eye(2), atol=1e-12)
    # padding lanes are zero
    n = len(g["V_k"])
    assert np.all(rhs[-1, :, :, n % W:] == 0) if n % W else True


def test_bodies_spmv_matches_scipy():
    g = load_golden("ico2_wave")
    hm = HostMesh(g)
    P = hm.P
    N, nb = P.n_vertices, P.n_blocks
    vals, rhs, _ = hm.assemble(g["I"], g["t_k"], float(g["lambda_"]))
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, N, 2, W))
    y = np.zeros_like(x)
    ms = hm.struct()
    hm.hc.hc_spmv(ctypes.byref(ms), 1, vals.ctypes.data, x.ctypes.data, y.ctypes.data)
    for k in range(len(g["V_k"])):
        a = _frame_matrix(P, vals, k)
        assert rel_l2(_frame_vector(P, y, k), a @ _frame_vector(P, x, k)) <= 1e-14


def _fan_case():
    """no golden for this one: the oracle (pinned by the goldens) supplies V_k"""
    from oracle import mof_oracle
    coords, tris, normals, areas = synthetic.fan_mesh(40, 3)
    t_k = synthetic.time_axis(4, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=3)
    a2, gw, e, integ = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    V, _ = mof_oracle.compute_velocity_field(1, 4, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    return dict(coordinates=coords, triangles=tris, normals=normals, areas=areas, t_k=np.asarray(t_k), I=I,
                lambda_=0.01, V_k=np.asarray(V))


@pytest.mark.parametrize("case", ["ico2_wave", "patch8_wave", "ico3_phase", "ico4_wave", "fan40"])
def test_bodies_ssor_eisenstat_pcg_matches_reference(case):
    """The SSOR path end to end on the CPU harness: block-multicolour ordering, D^-1/2 scaling from
    the assembly body, colour-by-colour backward / forward sweeps (the bodies the CUDA sweep
    kernels call) inside the same Eisenstat-form PCG the library's host loop runs."""
    g = _fan_case() if case == "fan40" else load_golden(case)
    hm = HostMesh(g, reorder=2)
    P = hm.P
    N = P.n_vertices
    assert P.n_colors >= 1 and P.color_tile_ptr[0] == 0 and P.color_tile_ptr[-1] == -(-N // _lib.TILE_ROWS)
    assert P.n_levels == 0
    # patches of one colour are mutually independent (no block couples two of them)
    tile_of = np.arange(N) // _lib.TILE_ROWS
    color_of_tile = np.repeat(np.arange(P.n_colors), np.diff(P.color_tile_ptr))
    rows, cols = P.block_rows(), P.col.astype(np.int64)
    cross = tile_of[rows] != tile_of[cols]
    assert np.all(color_of_tile[tile_of[rows[cross]]] != color_of_tile[tile_of[cols[cross]]])
    ptr = [int(x) for x in P.color_tile_ptr]
    ranges = [(ptr[c], ptr[c + 1]) for c in range(P.n_colors)]
    _eisenstat_pcg_on_host(hm, g, 1.4, fwd_ranges=ranges, back_ranges=ranges[::-1])


@pytest.mark.parametrize("case", ["ico2_wave", "patch8_wave", "ico3_phase", "ico4_wave", "fan40"])
def test_level_scheduled_ordering_and_its_ssor(case):
    """reorder = 3: Cuthill-McKee regrouped by dependency level.  Rows of a level are mutually
    independent, every neighbour with a lower number sits in a strictly lower level (so processing
    level by level IS the sequential Gauss-Seidel sweep, which is what the harness runs here), and
    the Eisenstat-form PCG on that ordering reaches the reference's fields."""
    g = _fan_case() if case == "fan40" else load_golden(case)
    hm = HostMesh(g, reorder=3)
    P = hm.P
    N = P.n_vertices
    assert P.n_colors == 0 and P.n_levels >= 1
    lp = P.level_ptr.astype(np.int64)
    assert lp[0] == 0 and lp[-1] == N and np.all(np.diff(lp) > 0)
    level_of = np.repeat(np.arange(P.n_levels), np.diff(lp))
    rows, cols = P.block_rows(), P.col.astype(np.int64)
    off = rows != cols
    assert np.all(level_of[rows[off]] != level_of[cols[off]])                        # independent rows inside a level
    assert np.all((cols[off] < rows[off]) == (level_of[cols[off]] < level_of[rows[off]]))   # orientation = level order
    # every row of level l > 0 depends on level l-1 (levels are as low as the order allows)
    lower = off & (cols < rows)
    deepest = np.full(N, -1)
    np.maximum.at(deepest, rows[lower], level_of[cols[lower]])
    assert np.array_equal(deepest + 1, level_of)
    assert sorted(P.perm.tolist()) == list(range(N))
    ntiles = -(-N // _lib.TILE_ROWS)
    iters = _eisenstat_pcg_on_host(hm, g, 1.6, fwd_ranges=[(0, ntiles)], back_ranges=[(t, t + 1) for t in range(ntiles - 1, -1, -1)])
    assert iters > 0


def _eisenstat_pcg_on_host(hm, g, omega, fwd_ranges, back_ranges):
    tol = 1e-12
    P, hc = hm.P, hm.hc
    N = P.n_vertices
    vals0, rhs0, _ = hm.assemble(g["I"], g["t_k"], float(g["lambda_"]), 0.0)       # unscaled system (block-Jacobi layout)
    vals, rhs, S = hm.assemble(g["I"], g["t_k"], float(g["lambda_"]), omega)       # S A S, S b, S = D^-1/2
    n = len(g["V_k"])
    G = vals.shape[0]
    # the scaling really is D^-1/2: scaled diagonal blocks are the identity
    dblk = vals[:, P.diag]                                                       # (G, N, 4, W)
    assert np.allclose(dblk[:, :, 0, :n % W or W], 1.0, atol=1e-13) and np.allclose(dblk[:, :, 1, :n % W or W], 0.0, atol=1e-13)
    ms = hm.struct()
    ref = ctypes.byref(ms)

    def back(mode, r, p, t, beta=None, zs=None):
        for t0, t1 in back_ranges:
            hc.hc_sweep_back(ref, G, vals.ctypes.data, r.ctypes.data, p.ctypes.data, t.ctypes.data, t0, t1,
                             beta.ctypes.data if beta is not None else None, zs.ctypes.data if zs is not None else None, omega, mode)

    def fwd(mode, pin, t, w, dot=None):
        for t0, t1 in fwd_ranges:
            hc.hc_sweep_fwd(ref, G, vals.ctypes.data, pin.ctypes.data, t.ctypes.data, w.ctypes.data, t0, t1,
                            omega, mode, dot.ctypes.data if dot is not None else None)

    def apply_S(v):
        o = np.empty_like(v)
        o[:, :, 0] = S[:, :, 0] * v[:, :, 0] + S[:, :, 1] * v[:, :, 1]
        o[:, :, 1] = S[:, :, 1] * v[:, :, 0] + S[:, :, 2] * v[:, :, 1]
        return o

    lanes = lambda a, b: np.einsum("gvcl,gvcl->gl", a, b)
    t = np.zeros_like(rhs)
    w = np.zeros_like(rhs)
    r = np.zeros_like(rhs)
    fwd(1, rhs, t, r)                                   # r = (Dt+L)^-1 S b
    x = np.zeros_like(rhs)
    p = np.zeros_like(rhs)
    rr = lanes(r, r)
    rz = rr / omega                                     # z = Dt r = r / omega
    bb = rr.copy()
    beta = np.zeros((G, W))
    zs = np.ones((G, W))
    active = bb > 0
    iters = 0
    while active.any() and iters < 2000:
        back(0, r, p, t, beta, zs)                      # p <- zs r/omega + beta p ; t = (Dt+U)^-1 p
        dot = np.zeros((G, W))
        fwd(0, p, t, w, dot)
        alpha = np.where(active, rz / np.where(dot != 0, dot, 1), 0.0)
        x += alpha[:, None, None, :] * p
        r -= alpha[:, None, None, :] * (t + w)
        rr = lanes(r, r)
        rz_new = rr / omega
        conv = active & (rr <= (0.3 * tol) ** 2 * bb)
        beta = np.where(active & ~conv, rz_new / np.where(rz != 0, rz, 1), 1.0)
        zs = np.where(active & ~conv, 1.0, 0.0)
        rz = np.where(active, rz_new, rz)
        active = active & ~conv
        iters += 1
    assert not active.any()
    xs = np.zeros_like(rhs)
    back(1, r, x, xs)                                   # xs = (Dt+U)^-1 xhat : solution of the scaled system
    xphys = apply_S(xs)                                 # x = S xs
    for k in range(n):
        a = _frame_matrix(P, vals0, k)
        b = _frame_vector(P, rhs0, k)
        V = _frame_vector(P, xphys, k)
        assert np.linalg.norm(a @ V - b) / np.linalg.norm(b) <= 1e-12
        assert rel_l2(V, g["V_k"][k]) <= 1e-8
    # padding lanes stay exactly zero
    assert np.all(xphys[-1, :, :, n % W:] == 0.0) if n % W else True
    return iters


def test_bodies_tangent_and_detection_match_reference(golden):
    g = golden
    hc = hostcheck.load()
    N, F = len(g["coordinates"]), len(g["triangles"])
    V = np.ascontiguousarray(g["V_k"])
    n = len(V)
    e = np.ascontiguousarray(g["e"])
    Vxyz, speed, vmax = np.zeros((n, N, 3)), np.zeros((n, N)), np.zeros(n)
    hc.hc_tangent(ctypes.c_int64(N), ctypes.c_int64(n), V.ctypes.data, ctypes.c_int64(2 * N), e.ctypes.data,
                  Vxyz.ctypes.data, speed.ctypes.data, vmax.ctypes.data)
    assert np.array_equal(Vxyz, g["V_xyz"])                      # bit-exact (separate mul/add like numpy)
    assert np.array_equal(vmax, g["v_length_max"])
    assert np.array_equal(speed, np.sqrt(np.sum(g["V_xyz"] ** 2, axis=2)))
    coords = np.ascontiguousarray(g["coordinates"])
    tri = np.ascontiguousarray(g["triangles"], dtype=np.int32)
    off_v = off_f = 0
    for k in range(n):
        nv, nf = ctypes.c_int32(), ctypes.c_int32()
        vidx, fidx = np.zeros(N, np.int32), np.zeros(F, np.int32)
        lam_mu, sign = np.zeros((F, 2)), np.zeros(F, np.int8)
        hc.hc_detect(ctypes.c_int64(N), ctypes.c_int64(F), coords.ctypes.data, tri.ctypes.data, Vxyz[k].ctypes.data,
                     ctypes.c_double(vmax[k]), ctypes.c_double(float(g["eps"])), ctypes.byref(nv), vidx.ctypes.data,
                     ctypes.byref(nf), fidx.ctypes.data, lam_mu.ctypes.data, sign.ctypes.data)
        cv, cf = g["sing_counts"][k]
        assert nv.value == cv and nf.value == cf
        assert np.array_equal(vidx[:cv], g["sing_vertex_idx"][off_v:off_v + cv])
        assert np.array_equal(fidx[:cf], g["sing_face_idx"][off_f:off_f + cf])
        assert np.allclose(lam_mu[:cf], g["sing_face_lam_mu"][off_f:off_f + cf], rtol=0, atol=1e-10)
        assert np.all(np.abs(sign[:cf]) == 1)
        off_v += cv
        off_f += cf


def test_bodies_detection_with_singular_vertices():
    g = load_golden("ico2_vertex_singular")
    hc = hostcheck.load()
    coords = np.ascontiguousarray(g["coordinates"])
    tri = np.ascontiguousarray(g["triangles"], dtype=np.int32)
    N, F = len(coords), len(tri)
    V = np.ascontiguousarray(g["V_now"])
    nv, nf = ctypes.c_int32(), ctypes.c_int32()
    vidx, fidx = np.zeros(N, np.int32), np.zeros(F, np.int32)
    lam_mu, sign = np.zeros((F, 2)), np.zeros(F, np.int8)
    hc.hc_detect(ctypes.c_int64(N), ctypes.c_int64(F), coords.ctypes.data, tri.ctypes.data, V.ctypes.data,
                 ctypes.c_double(float(g["v_length_max"])), ctypes.c_double(float(g["eps"])), ctypes.byref(nv),
                 vidx.ctypes.data, ctypes.byref(nf), fidx.ctypes.data, lam_mu.ctypes.data, sign.ctypes.data)
    assert np.array_equal(vidx[:nv.value], g["sing_vertex_idx"])
    assert np.array_equal(fidx[:nf.value], g["sing_face_idx"])


def test_frame_dt_matches_python_float_arithmetic():
    t_k = synthetic.time_axis(50, 512.0)
    dt = frame_dt(t_k, 3, 40)
    assert all(dt[i] == t_k[3 + i + 1] - t_k[3 + i] for i in range(37))


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from manifold_based_optical_flow_method_b200 import compute_optical_flow, find_singularity_point
    coords, tris, normals, areas = synthetic.icosphere(1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compute_optical_flow.compute_geometrical_quantities(coords, normals, tris, areas)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        find_singularity_point.process_V_k(np.zeros((1, 84)), np.zeros((42, 2, 3)))


def test_classification_host_helpers(capsys):
    """classify_critical_point / analyze_classification (reference fsp:463-527) against the oracle's rule."""
    from manifold_based_optical_flow_method_b200 import find_singularity_point as fsp
    from oracle import mof_oracle
    rng = np.random.default_rng(0)
    for J in list(rng.normal(size=(200, 2, 2))) + [np.zeros((2, 2)), np.eye(2), np.array([[0.0, 1.0], [-1.0, 0.0]]),
                                                   np.array([[np.nan, 0.0], [0.0, 1.0]])]:
        assert fsp.classify_critical_point(J) == mof_oracle.CLASS_NAMES[mof_oracle.classify(J)]
    got = fsp.analyze_classification([["Focus", "Saddle"], [], ["Node", "Indeterminate", "Saddle"]])
    assert got == {"Focus": 1, "Saddle": 2, "Node": 2}
    assert capsys.readouterr().out == "Focus: 1\nSaddle: 2\nNode: 2\n"


def test_level_sweeps_do_not_depend_on_the_order_inside_a_level():
    """What the level kernels rely on: inside one dependency level the rows can be processed in any order
    (the GPU runs them concurrently) -- the backward and forward sweeps give bit-identical vectors whether a
    level's rows are taken ascending, descending or shuffled, and the same as the plain sequential sweep."""
    g = load_golden("ico3_phase")
    omega = 1.9
    hm = HostMesh(g, reorder=3)
    P, hc = hm.P, hm.hc
    N = P.n_vertices
    vals, rhs, S = hm.assemble(g["I"], g["t_k"], float(g["lambda_"]), omega)
    G = vals.shape[0]
    ms = hm.struct()
    ref = ctypes.byref(ms)
    lp = P.level_ptr.astype(np.int64)
    rng = np.random.default_rng(0)
    pin = rng.normal(size=rhs.shape)
    rvec = rng.normal(size=rhs.shape)
    beta = rng.normal(size=(G, W))
    zs = np.ones((G, W))

    def run(order_of_level):
        p = pin.copy()
        t = np.zeros_like(rhs)
        w = np.zeros_like(rhs)
        dot = np.zeros((G, W))
        for l in range(P.n_levels - 1, -1, -1):                       # backward sweep: levels descending
            for row in order_of_level(np.arange(lp[l], lp[l + 1])):
                hc.hc_sweep_rows_back(ref, G, vals.ctypes.data, rvec.ctypes.data, p.ctypes.data, t.ctypes.data,
                                      int(row), int(row) + 1, beta.ctypes.data, zs.ctypes.data, omega, 0)
        for l in range(P.n_levels):                                    # forward sweep: levels ascending
            for row in order_of_level(np.arange(lp[l], lp[l + 1])):
                hc.hc_sweep_rows_fwd(ref, G, vals.ctypes.data, p.ctypes.data, t.ctypes.data, w.ctypes.data,
                                     int(row), int(row) + 1, omega, 0, dot.ctypes.data)
        return p, t, w

    base = run(lambda rows: rows[::-1])
    for order in (lambda rows: rows, lambda rows: rng.permutation(rows)):
        got = run(order)
        assert all(np.array_equal(a, b) for a, b in zip(got, base))
    # and equal to the plain sequential sweeps over all rows
    p = pin.copy()
    t = np.zeros_like(rhs)
    w = np.zeros_like(rhs)
    hc.hc_sweep_rows_back(ref, G, vals.ctypes.data, rvec.ctypes.data, p.ctypes.data, t.ctypes.data, 0, N,
                          beta.ctypes.data, zs.ctypes.data, omega, 0)
    hc.hc_sweep_rows_fwd(ref, G, vals.ctypes.data, p.ctypes.data, t.ctypes.data, w.ctypes.data, 0, N, omega, 0, None)
    assert all(np.array_equal(a, b) for a, b in zip((p, t, w), base))


def test_upload_chunks_cover_every_frame():
    """Streamed upload (solver.SignalUpload): copy c must bring every row that the frames of chunk c read (rows k and
    k + 1 of frame k), rows are copied once and in order."""
    from manifold_based_optical_flow_method_b200.solver import upload_chunks
    for step in (32, 128):
        for n_rows in (1, 2, step, step + 1, step + 2, 3 * step, 3 * step + 1, 1000):
            ch = upload_chunks(n_rows, step)
            assert ch[0][0] == 0 and ch[-1][1] == n_rows
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(r1 > r0 for r0, r1 in ch)
            n_frames = n_rows - 1
            for c, f0 in enumerate(range(0, n_frames, step)):
                f1 = min(n_frames, f0 + step)
                assert ch[min(c, len(ch) - 1)][1] >= f1 + 1, (step, n_rows, c)

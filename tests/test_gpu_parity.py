"""GPU parity tests (run on the B200 box): the CUDA path, called through the reference-shaped
Python API and the C ABI, against (a) the golden outputs of the unmodified reference
(tests/golden/*.npz), (b) the numpy/scipy oracle on seeded inputs the oracle finishes in
seconds, and (c) size-independent properties at the full BASELINE.json sizes.

Tolerances (BASELINE.json north_star): velocity fields rel-L2 <= 1e-8 (fp64, PCG true
residual <= 1e-12); assembled values/rhs <= 1e-13 relative; singular vertex / face indices
exactly equal; (lam, mu) within 1e-10.
"""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden, rel_l2
from manifold_based_optical_flow_method_b200 import _lib, synthetic
from oracle import mof_oracle

from manifold_based_optical_flow_method_b200.solver import DEFAULT_PRECOND  # noqa: E402

pytestmark = pytest.mark.gpu

W = _lib.GROUP
V_TOL = 1e-8
RES_TOL = 1e-12


@pytest.fixture(scope="module", params=["ssor_level", "ssor", "jacobi"])
def mods(request):
    """Every test runs with all three preconditioners (and therefore all three vertex numberings)."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from manifold_based_optical_flow_method_b200 import compute_optical_flow, find_singularity_point
    old = compute_optical_flow.settings["precond"]
    compute_optical_flow.settings["precond"] = request.param
    yield compute_optical_flow, find_singularity_point
    compute_optical_flow.settings["precond"] = old


def _csr(g, prefix, n):
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))


def _frame_matrix(P, vals, k):
    return P.block_values_to_csr(vals[k // W, :, :, k % W])


def _frame_vector(P, vec, k):
    N = P.n_vertices
    out = np.empty(2 * N)
    out[P.perm] = vec[k // W, :, 0, k % W]
    out[P.perm + N] = vec[k // W, :, 1, k % W]
    return out


def _assemble_on_gpu(cof, op, I, t_k, lambda_):
    """pack + K1 through the C ABI; returns host copies of vals / rhs / minv and the batch"""
    import torch
    from manifold_based_optical_flow_method_b200.solver import VelocitySolver, frame_dt
    n = len(I) - 1
    s = VelocitySolver(op, batch_groups=-(-n // W), precond="jacobi")
    batch = s.batch(-(-n // W))
    I_dev = torch.from_numpy(np.ascontiguousarray(I, dtype=np.float64)).to(op.device)
    dt = torch.from_numpy(frame_dt(list(t_k), 0, n)).to(op.device)
    s.assemble(batch, I_dev[:n], I_dev[1:n + 1], dt, lambda_, n)
    torch.cuda.synchronize()
    return s, batch, batch.vals.cpu().numpy(), batch.rhs.cpu().numpy(), batch.minv.cpu().numpy()


# ---------------------------------------------------------------------------------
# golden vectors of the unmodified reference
# ---------------------------------------------------------------------------------
def test_geometry_matches_reference(golden, mods):
    cof, _ = mods
    g = golden
    a2, grad_w, e, integral, secs = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    assert a2.shape == (2 * len(g["coordinates"]),) * 2 and secs >= 0
    assert np.max(np.abs(e - g["e"])) <= 1e-15
    assert rel_l2(grad_w, g["grad_w"]) <= 1e-14
    assert np.allclose(integral, g["integral_wi_wj"], rtol=1e-15, atol=0)
    if "a2_data" in g:
        ref = _csr(g, "a2", a2.shape[0])
        assert abs(a2.tocsr() - ref).max() <= 1e-13 * abs(ref).max()


def test_assembled_system_matches_reference(golden, mods):
    cof, _ = mods
    g = golden
    N = len(g["coordinates"])
    op, gw, e, integ, _ = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    _, _, vals, rhs, minv = _assemble_on_gpu(cof, op, g["I"], g["t_k"], float(g["lambda_"]))
    P = op.pattern
    if "a0_data" in g:
        ref = _csr(g, "a0", 2 * N)
        a = _frame_matrix(P, vals, 0)
        assert abs(a - ref).max() <= 1e-13 * abs(ref).max()
        assert abs(a - a.T).max() == 0.0
        assert rel_l2(_frame_vector(P, rhs, 0), g["f0"]) <= 1e-13
    # every frame against the oracle's assembly
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    for k in range(len(g["I"]) - 1):
        dt = g["t_k"][k + 1] - g["t_k"][k]
        a1, f = mof_oracle.assemble_frame(gwo, eo, into, g["triangles"], g["areas"], dt, g["I"][k], g["I"][k + 1])
        ao = mof_oracle.system_matrix(a1, a2o, float(g["lambda_"]))
        a = _frame_matrix(P, vals, k)
        assert abs(a - ao).max() <= 1e-13 * abs(ao).max()
        assert rel_l2(_frame_vector(P, rhs, k), f) <= 1e-13
    # padded lanes carry a zero rhs
    n = len(g["I"]) - 1
    assert np.all(rhs[-1, :, :, n % W:] == 0.0)
    assert np.all(np.isfinite(minv))


def test_spmv_matches_scipy(mods):
    import torch
    cof, _ = mods
    g = load_golden("ico3_phase")
    op, *_ = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    s, batch, vals, _, _ = _assemble_on_gpu(cof, op, g["I"], g["t_k"], float(g["lambda_"]))
    P = op.pattern
    x = torch.randn(batch.x.shape, dtype=torch.float64, device=op.device, generator=torch.Generator(op.device).manual_seed(3))
    y = torch.full_like(x, float("nan"))
    ms, bs = op.struct(), batch.struct()
    _lib.check(_lib.load().mof_spmv_batch(ctypes.byref(ms), ctypes.byref(bs), x.data_ptr(), y.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    for k in range(len(g["I"]) - 1):
        a = _frame_matrix(P, vals, k)
        assert rel_l2(_frame_vector(P, yh, k), a @ _frame_vector(P, xh, k)) <= 1e-14


def test_velocity_field_matches_reference(golden, mods):
    cof, _ = mods
    g = golden
    T = len(g["t_k"])
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    V_k, secs = cof.compute_velocity_field(4, T, a2, gw, e, integ, g["triangles"], list(g["t_k"]), g["areas"],
                                           float(g["lambda_"]), g["I"], g["I"])
    assert isinstance(V_k, list) and len(V_k) == T - 1 and V_k[0].shape == (2 * len(g["coordinates"]),)
    info = cof.last_solve_info
    assert info.converged and np.all(info.relres <= RES_TOL)
    for k in range(T - 1):
        assert rel_l2(V_k[k], g["V_k"][k]) <= V_TOL, (k, rel_l2(V_k[k], g["V_k"][k]))
    # the reference's own consumers of V_k (compute_optical_flow.py:314-319)
    assert np.array(V_k).reshape(len(V_k), -1).shape == (T - 1, 2 * len(g["coordinates"]))


def test_worker_single_frame_matches_reference_and_batch(mods):
    cof, _ = mods
    g = load_golden("ico2_wave")
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    t_k = list(g["t_k"])
    V1 = cof.worker(1, a2, gw, e, integ, g["triangles"], t_k, g["areas"], float(g["lambda_"]), g["I"][1], g["I"][2])
    assert rel_l2(V1, g["V_k"][1]) <= V_TOL
    V_k, _ = cof.compute_velocity_field(1, len(t_k), a2, gw, e, integ, g["triangles"], t_k, g["areas"],
                                        float(g["lambda_"]), g["I"], g["I"])
    # reductions are deterministic and lane-private: a frame does not depend on its batch
    assert np.array_equal(V1, V_k[1])


def test_second_signal_argument(mods):
    """I_k_2 != I_k: frame k must use I_k[k] and I_k_2[k+1] (compute_optical_flow.py:174-175)."""
    cof, _ = mods
    g = load_golden("ico1_wave")
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    I2 = g["I"] + 0.01 * np.random.default_rng(5).standard_normal(g["I"].shape)
    T = len(g["t_k"])
    Vo, _ = mof_oracle.compute_velocity_field(1, T, a2o, gwo, eo, into, g["triangles"], list(g["t_k"]), g["areas"], 0.02, g["I"], I2)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, g["triangles"], list(g["t_k"]), g["areas"], 0.02, g["I"], I2)
    for k in range(T - 1):
        assert rel_l2(V_k[k], Vo[k]) <= V_TOL


def test_process_V_k_and_singularities_match_reference(golden, mods):
    _, fsp = mods
    g = golden
    V_xyz = fsp.process_V_k(list(g["V_k"]), g["e"])
    assert V_xyz.shape == g["V_xyz"].shape
    assert np.array_equal(V_xyz, g["V_xyz"])                       # bit-exact
    assert np.array_equal(fsp.speed_magnitude(g["V_k"], g["e"]), np.sqrt(np.sum(g["V_xyz"][:, :, :3] ** 2, axis=2)))
    s = fsp.detect_singularities(g["V_xyz"], g["coordinates"], g["triangles"], float(g["eps"]))
    assert np.array_equal(s.v_length_max, g["v_length_max"])
    assert np.array_equal(np.diff(s.vertex_offsets), g["sing_counts"][:, 0])
    assert np.array_equal(np.diff(s.face_offsets), g["sing_counts"][:, 1])
    assert np.array_equal(s.vertex_idx, g["sing_vertex_idx"])
    assert np.array_equal(s.face_idx, g["sing_face_idx"])          # bit-exact index lists
    assert np.allclose(s.lam_mu, g["sing_face_lam_mu"], rtol=0, atol=1e-10)
    R = np.max(np.linalg.norm(g["coordinates"], axis=1))
    assert np.allclose(s.P, g["sing_face_P"], rtol=0, atol=1e-10 * R)
    assert np.all(np.abs(s.index) == 1)
    # per-face Poincare index agrees with the winding-angle definition of S7_winding_line.py
    for k in range(len(g["V_xyz"])):
        _, fi, _, _, idx = s.frame(k)
        assert np.array_equal(idx, mof_oracle.face_poincare_index(g["coordinates"], g["triangles"], g["V_xyz"][k], fi))
    # reference-shaped API, one frame
    sv, si, vmax = fsp.find_singularity_points(g["coordinates"], g["triangles"], g["V_xyz"][0], float(g["eps"]))
    nv, nf = g["sing_counts"][0]
    assert [r[0] for r in sv] == list(g["sing_vertex_idx"][:nv]) and [r[0] for r in si] == list(g["sing_face_idx"][:nf])
    assert vmax == g["v_length_max"][0]
    for r in si:
        assert len(r) == 5 and np.array_equal(r[2], g["triangles"][r[0]]) and abs(sum(r[3]) - 1) < 1e-15
    pts = fsp.find_singularity_points_for_all_Vk(g["V_xyz"], g["coordinates"], g["triangles"], float(g["eps"]))
    flat = np.asarray([p for fr in pts for p in fr]).reshape(-1, 3)
    assert np.allclose(flat, g["all_points_flat"], rtol=0, atol=1e-10 * R)


def test_singular_vertices_and_face_skip(mods):
    _, fsp = mods
    g = load_golden("ico2_vertex_singular")
    sv, si, vmax = fsp.find_singularity_points(g["coordinates"], g["triangles"], g["V_now"], float(g["eps"]))
    assert [r[0] for r in sv] == list(g["sing_vertex_idx"]) and len(sv) == 2
    assert np.array_equal(sv[0][1], g["coordinates"][sv[0][0]])
    assert [r[0] for r in si] == list(g["sing_face_idx"])
    assert vmax == float(g["v_length_max"])


def test_interior_zero_next_to_an_edge_matches_reference(mods):
    """fsp:126-131 at the edge of its domain: fields whose zero lies at barycentric (0.3, delta, 0.7 - delta) of one
    face, goldens from the unmodified reference (tests/golden/make_golden_edge.py).  The reference decides with an
    SVD least squares and hard thresholds, this library with the closed-form solution of the same system: for
    |delta| >= 1e-12 the reported faces must be identical; below that (rounding decides in both) the zero must be
    reported in at least one of the two faces sharing the edge and nothing else may change."""
    _, fsp = mods
    g = load_golden("edge_zero_ico2")
    coords, tris, eps = g["coordinates"], g["triangles"], float(g["eps"])
    pair = {int(g["face"]), 32}                              # the face and its neighbour across the edge (see the generator's log)
    off_v = off_f = 0
    for k, delta in enumerate(g["deltas"]):
        nv, nf = (int(x) for x in g["counts"][k])
        ref_faces = [int(x) for x in g["sing_face_idx"][off_f:off_f + nf]]
        ref_lm = g["sing_face_lam_mu"][off_f:off_f + nf]
        off_v, off_f = off_v + nv, off_f + nf
        sv, si, vmax = fsp.find_singularity_points(coords, tris, g["V_now"][k], eps)
        faces = [int(r[0]) for r in si]
        assert len(sv) == nv and vmax == float(g["v_length_max"][k])
        if abs(delta) >= 1e-12:
            assert faces == ref_faces, (delta, faces, ref_faces)
            assert np.allclose([r[3][:2] for r in si], ref_lm, rtol=0, atol=1e-10)
        else:
            assert [f for f in faces if f not in pair] == [f for f in ref_faces if f not in pair]
            assert pair & set(faces), (delta, faces)
        for r in si:                                         # whatever is reported is a valid barycentric triple
            lam, mu = r[3][0], r[3][1]
            assert lam >= 0 and mu >= 0 and lam + mu <= 1


# ---------------------------------------------------------------------------------
# oracle on seeded inputs (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["C1_wave", "patch_wave", "two_phase", "fan_wave"])
def test_velocity_field_matches_oracle(case, mods):
    cof, fsp = mods
    if case == "C1_wave":          # BASELINE.json configs[0] mesh (ico5, 10,242 vertices)
        coords, tris, normals, areas = synthetic.icosphere(5)
        T, SF, kind = 7, 512.0, "wave"
    elif case == "patch_wave":     # open surface with boundary, irregular valence
        coords, tris, normals, areas = synthetic.open_patch(40, seed=2)
        T, SF, kind = 36, 256.0, "wave"      # 35 frames: one full group + 3 ragged lanes
    elif case == "fan_wave":       # a vertex of valence 40: block rows longer than a warp / than a load batch
        coords, tris, normals, areas = synthetic.fan_mesh(40, 3)
        T, SF, kind = 5, 512.0, "wave"
    else:                          # two components + wrapped-phase input (config 4 style)
        coords, tris, normals, areas = synthetic.two_hemispheres(4, radius=80.0)
        T, SF, kind = 4, 512.0, "phase"
    t_k = synthetic.time_axis(T, SF)
    I = synthetic.travelling_wave(coords, t_k, seed=1) if kind == "wave" else synthetic.wrapped_phase(coords, t_k, seed=1)
    lam = 0.01
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    Vo, _ = mof_oracle.compute_velocity_field(4, T, a2o, gwo, eo, into, tris, t_k, areas, lam, I, I)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    assert abs(a2.tocsr() - a2o).max() <= 1e-13 * abs(a2o).max()
    V_k, _ = cof.compute_velocity_field(8, T, a2, gw, e, integ, tris, t_k, areas, lam, I, I)
    info = cof.last_solve_info
    assert info.converged and np.all(info.relres <= RES_TOL)
    worst = max(rel_l2(V_k[k], Vo[k]) for k in range(T - 1))
    assert worst <= V_TOL, worst
    # detection on the GPU field vs the oracle's detection on the oracle field
    Vx = fsp.process_V_k(V_k, e)
    Vxo = mof_oracle.process_V_k(Vo, eo)
    s = fsp.detect_singularities(Vx, coords, tris, 1e-4)
    for k in range(min(T - 1, 3)):
        vi, fi, lm, P, vmax = mof_oracle.find_singularity_points(coords, tris, Vxo[k], 1e-4)
        gvi, gfi, glm, gP, _ = s.frame(k)
        assert np.array_equal(gvi, vi) and np.array_equal(gfi, fi)
        assert np.allclose(glm, lm, rtol=0, atol=1e-6)      # fields differ by <=1e-8 rel-L2


# ---------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n_frames", [1, 31, 32, 33, 70])
def test_ragged_frame_counts(n_frames, mods):
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(2)
    T = n_frames + 1
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=4)
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    cof.settings["batch_groups"] = 2          # forces several batches for 70 frames
    try:
        V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    finally:
        cof.settings["batch_groups"] = None
    assert len(V_k) == n_frames
    for k in sorted({0, n_frames // 2, n_frames - 1}):
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        assert rel_l2(V_k[k], Vo) <= V_TOL


def test_pinned_and_pageable_host_delivery_agree(mods):
    """Results delivered by direct DMA into pinned memory (default) and through the staged drain into
    pageable memory are the same bits; a held result is not overwritten by the next call."""
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(2)
    T = 71
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=5)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    cof.settings["batch_groups"] = 1
    try:
        assert cof.settings["pinned_results"] is True
        V_pin, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
        keep = np.array(V_pin)
        cof.settings["pinned_results"] = False
        V_page, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
        cof.settings["pinned_results"] = True
        V_other, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, 2.0 * I, 2.0 * I)
    finally:
        cof.settings["batch_groups"] = None
        cof.settings["pinned_results"] = True
    assert np.array_equal(np.array(V_page), keep)
    assert np.array_equal(np.array(V_pin), keep)            # still intact while V_other exists
    assert not np.array_equal(np.array(V_other), keep)
    assert V_pin[0].flags.writeable and V_pin[0].dtype == np.float64


def test_torch_tensor_inputs(mods):
    """I_k may be a torch tensor (CPU or CUDA, float32 is promoted) as well as numpy or a list of rows."""
    import torch
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(2)
    T = 5
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=6)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    ref, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    for form in (torch.from_numpy(I), torch.from_numpy(I).cuda(), [row for row in I]):
        got, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, form, form)
        assert np.array_equal(np.array(got), np.array(ref))
    with pytest.raises(ValueError):
        cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, torch.zeros(T, 7), torch.zeros(T, 7))


def test_concurrent_streams_match_single_stream(mods):
    """Batches dealt to two concurrent solve streams give bit-identical fields to the single-stream
    path (every frame's arithmetic is private to its lane; reductions are deterministic)."""
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(3)
    T = 200                           # 199 frames = 7 groups -> 2 streams x batches of 2 groups
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=6)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    out = {}
    try:
        for streams in (1, 2):
            cof.settings["streams"] = streams
            cof.settings["batch_groups"] = 4
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            assert cof.last_solve_info.converged and len(cof.last_solve_info.iterations) == T - 1
            out[streams] = np.asarray(V_k)
    finally:
        cof.settings["streams"] = None
        cof.settings["batch_groups"] = None
    assert np.array_equal(out[1], out[2])
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    for k in (0, 100, 198):
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        assert rel_l2(out[2][k], Vo) <= V_TOL


@pytest.mark.parametrize("mesh", ["ico3", "patch"])
def test_no_out_of_bounds_writes_canary(mesh, mods):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are checked with guard
    bands: every batch buffer is re-seated inside a larger allocation whose margins hold a sentinel;
    after pack + assemble + solve + unpack the margins must be untouched (and the result right)."""
    import torch
    from manifold_based_optical_flow_method_b200.solver import VelocitySolver, FrameBatch, frame_dt
    cof, _ = mods
    if mesh == "ico3":
        coords, tris, normals, areas = synthetic.icosphere(3)        # 642 vertices: 10 full tiles + 2 rows
    else:
        coords, tris, normals, areas = synthetic.open_patch(19, seed=4)   # 361 vertices, irregular valence
    T = 41                                                            # 40 frames: one full group + 8 lanes
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=8)
    op, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    precond = cof.settings["precond"]
    s = VelocitySolver(op, batch_groups=2, precond=precond)
    batch = FrameBatch(op, 2, with_t=precond != "jacobi")
    SENT, PAD = 1.2345e300, 4096
    guards = {}
    for name in ("It", "dIt", "vals", "rhs", "minv", "x", "r", "z", "p", "ap", "t", "partial", "scal"):
        buf = getattr(batch, name)
        if buf is None:
            continue
        big = torch.full((buf.numel() + 2 * PAD,), SENT, dtype=torch.float64, device=op.device)
        inner = big[PAD:PAD + buf.numel()].view(buf.shape)
        inner.copy_(buf)
        setattr(batch, name, inner)
        guards[name] = big
    ibig = torch.full((batch.state.numel() + 2 * PAD,), 0x5A5A5A5A, dtype=torch.int32, device=op.device)
    inner = ibig[PAD:PAD + batch.state.numel()]
    inner.copy_(batch.state)
    batch.state = inner
    N = len(coords)
    vbig = torch.full(((T - 1) * 2 * N + 2 * PAD,), SENT, dtype=torch.float64, device=op.device)
    V = vbig[PAD:PAD + (T - 1) * 2 * N].view(T - 1, 2 * N)
    I_dev = torch.from_numpy(I).to(op.device)
    dt = torch.from_numpy(frame_dt(t_k, 0, T - 1)).to(op.device)
    info = s.solve_batch(I_dev[:T - 1], I_dev[1:T], dt, 0.01, V, batch=batch)
    torch.cuda.synchronize()
    assert info.converged
    for name, big in list(guards.items()) + [("V", vbig)]:
        assert bool((big[:PAD] == SENT).all()) and bool((big[-PAD:] == SENT).all()), f"guard band of {name} was written"
    assert bool((ibig[:PAD] == 0x5A5A5A5A).all()) and bool((ibig[-PAD:] == 0x5A5A5A5A).all()), "guard band of state was written"
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    Vh = V.cpu().numpy()
    for k in (0, 31, 39):
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        assert rel_l2(Vh[k], Vo) <= V_TOL


def test_empty_time_axis(mods):
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(1)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    V_k, secs = cof.compute_velocity_field(1, 1, a2, gw, e, integ, tris, [0.0], areas, 0.01, np.zeros((1, 42)), np.zeros((1, 42)))
    assert V_k == [] and secs == 0.0


def test_constant_and_nan_frames_do_not_poison_others(mods):
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(2)
    T = 6
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=9)
    I[2] = 0.0                      # frame 2: grad I = 0 exactly -> a1 = 0, f = 0 -> V = 0 (spsolve(a, 0) = 0)
    I[5, 7] = np.nan                # frame 4 reads I[5] as "next" -> NaN rhs
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    cof.settings["allow_unconverged"] = True
    try:
        V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    finally:
        cof.settings["allow_unconverged"] = False
    info = cof.last_solve_info
    assert info.status[2] == _lib.STATUS_ZERO_RHS and np.all(V_k[2] == 0.0)
    assert info.status[4] == _lib.STATUS_BREAKDOWN
    for k in (0, 1, 3):
        assert info.status[k] == _lib.STATUS_CONVERGED
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        assert rel_l2(V_k[k], Vo) <= V_TOL
    with pytest.raises(cof.UnconvergedError):
        cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)


def test_max_iter_reports_unconverged(mods):
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(3)
    t_k = synthetic.time_axis(3, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=2)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    cof.settings["max_iter"] = 5
    try:
        with pytest.raises(cof.UnconvergedError):
            cof.compute_velocity_field(1, 3, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
        assert np.all(cof.last_solve_info.status == _lib.STATUS_MAXITER)
        assert np.all(cof.last_solve_info.iterations == 5)
        assert np.all(cof.last_solve_info.relres > 1e-12)
    finally:
        cof.settings["max_iter"] = 20000


def test_caller_supplied_geometry_and_reference_a2(mods):
    """worker uses the grad_w / e / integral it is handed (compute_optical_flow.py:100-101)."""
    cof, _ = mods
    from manifold_based_optical_flow_method_b200.mesh import MeshOperator
    g = load_golden("ico2_wave")
    t_k = list(g["t_k"])
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(g["coordinates"], g["normals"], g["triangles"], g["areas"])
    # rotate every tangent basis by 30 degrees: a different but valid e
    c, s_ = np.cos(0.5), np.sin(0.5)
    e2 = np.stack([c * eo[:, 0] + s_ * eo[:, 1], -s_ * eo[:, 0] + c * eo[:, 1]], axis=1)
    a2_rot = mof_oracle.a2_matrix(g["triangles"], g["areas"], e2, gwo)
    Vo = mof_oracle.worker(0, a2_rot, gwo, e2, into, g["triangles"], t_k, g["areas"], 0.01, g["I"][0], g["I"][1])
    op = MeshOperator.from_reference_a2(a2_rot, g["coordinates"], g["normals"], g["triangles"], g["areas"])
    V = cof.worker(0, op, gwo, e2, into, g["triangles"], t_k, g["areas"], 0.01, g["I"][0], g["I"][1])
    assert rel_l2(V, Vo) <= V_TOL
    with pytest.raises(TypeError):
        cof.worker(0, a2o, gwo, eo, into, g["triangles"], t_k, g["areas"], 0.01, g["I"][0], g["I"][1])


def test_c_abi_argument_errors(mods):
    lib = _lib.load()
    assert lib.mof_geom_basis(0, None, None, None) < 0
    assert "bad arguments" in _lib.last_error()
    with pytest.raises(_lib.MofError):
        _lib.check(lib.mof_tangent_to_xyz(10, 1, None, 20, None, None, None, None, None))


# ---------------------------------------------------------------------------------
# full size (BASELINE.json configs[1]: ~160k-vertex pial-like mesh): size-independent properties
# ---------------------------------------------------------------------------------
def test_full_size_properties(mods):
    import torch
    cof, fsp = mods
    if cof.settings["precond"] != DEFAULT_PRECOND:
        pytest.skip("full-size properties run once, with the default preconditioner")
    coords, tris, normals, areas = synthetic.pial_like(7)
    N = len(coords)
    assert N == 163842
    T = 34                          # 33 frames: one full group + one ragged lane
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=0)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    info = cof.last_solve_info
    assert info.converged and np.all(info.relres <= RES_TOL)
    # (1) residual against the ORACLE's independently assembled system, frames 0 and 32
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    assert abs(a2.tocsr() - a2o).max() <= 1e-13 * abs(a2o).max()
    for k in (0, 32):
        a1, f = mof_oracle.assemble_frame(gwo, eo, into, tris, areas, t_k[k + 1] - t_k[k], I[k], I[k + 1])
        A = mof_oracle.system_matrix(a1, a2o, 0.01)
        assert np.linalg.norm(A @ V_k[k] - f) / np.linalg.norm(f) <= 1e-11
    # (2) linearity in the rhs: doubling every time step halves the field (f ~ 1/dt)
    t2 = [2.0 * t for t in t_k]
    Vh = cof.worker(5, a2, gw, e, integ, tris, t2, areas, 0.01, I[5], I[6])
    assert rel_l2(2.0 * Vh, V_k[5]) <= 1e-9
    # (3) batch independence (deterministic, lane-private reductions): frame alone == frame in batch
    V5 = cof.worker(5, a2, gw, e, integ, tris, t_k, areas, 0.01, I[5], I[6])
    assert np.array_equal(V5, V_k[5])
    # (4) one full-size frame against the reference algorithm's direct solve (SuperLU, ~40 s)
    Vo = mof_oracle.worker(0, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[0], I[1])
    assert rel_l2(V_k[0], Vo) <= V_TOL
    # (5) detection at full size against the oracle (reference lstsq decisions), exact index lists
    Vx = fsp.process_V_k(V_k[:4], e)
    assert np.array_equal(Vx, mof_oracle.process_V_k(V_k[:4], e))
    s = fsp.detect_singularities(Vx, coords, tris, 1e-4)
    for k in range(4):
        vi, fi, lm, P, idx = s.frame(k)
        assert np.all((lm >= 0) & (lm.sum(axis=1, keepdims=True) <= 1)) and len(fi) >= 2
    for k in (0, 3):
        vio, fio, lmo, Po, vmaxo = mof_oracle.find_singularity_points(coords, tris, Vx[k], 1e-4)
        vi, fi, lm, P, idx = s.frame(k)
        assert np.array_equal(vi, vio) and np.array_equal(fi, fio) and s.v_length_max[k] == vmaxo
        assert np.allclose(lm, lmo, rtol=0, atol=1e-10)
        assert np.array_equal(idx, mof_oracle.face_poincare_index(coords, tris, Vx[k], fio))
    torch.cuda.empty_cache()


def test_full_size_phase_config4(mods):
    """BASELINE.json configs[3]: ~320k-vertex two-component mesh with wrapped-phase input (values in
    (-pi, pi], ill conditioned: cond ~1e6) -> velocity solve + singularity detection."""
    cof, fsp = mods
    if cof.settings["precond"] != DEFAULT_PRECOND:
        pytest.skip("run once, with the default preconditioner")
    coords, tris, normals, areas = synthetic.two_hemispheres(7)
    N = len(coords)
    assert N == 327684
    T = 4
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.wrapped_phase(coords, t_k, seed=0)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    info = cof.last_solve_info
    assert info.converged and np.all(info.relres <= RES_TOL)
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    assert abs(a2.tocsr() - a2o).max() <= 1e-13 * abs(a2o).max()
    # residual of every frame against the oracle-assembled system
    for k in range(T - 1):
        a1, f = mof_oracle.assemble_frame(gwo, eo, into, tris, areas, t_k[k + 1] - t_k[k], I[k], I[k + 1])
        A = mof_oracle.system_matrix(a1, a2o, 0.01)
        assert np.linalg.norm(A @ V_k[k] - f) / np.linalg.norm(f) <= 1e-11
    # the reference algorithm's direct solve (SuperLU) for one frame: the parity gate itself
    from scipy.sparse.linalg import spsolve
    Vo = spsolve(A, f)
    worst = rel_l2(V_k[T - 2], Vo)
    assert worst <= V_TOL, worst
    # detection: exact index lists against the oracle on the same field
    Vx = fsp.process_V_k(V_k[:1], e)
    s = fsp.detect_singularities(Vx, coords, tris, 1e-4)
    vio, fio, lmo, Po, vmaxo = mof_oracle.find_singularity_points(coords, tris, Vx[0], 1e-4)
    vi, fi, lm, P, idx = s.frame(0)
    assert np.array_equal(vi, vio) and np.array_equal(fi, fio) and s.v_length_max[0] == vmaxo
    assert len(fi) > 100      # wrapped phases create many critical points


def test_full_size_level_scheduled_path():
    """ico7 (163,842 vertices, ~1000 dependency levels per sweep, replayed as a CUDA graph): the
    level-scheduled SSOR path converges to the true-residual tolerance in far fewer iterations than the
    block-multicolour path and both give the same fields."""
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.pial_like(7)
    T = 34
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=0)
    old = cof.settings["precond"]
    out = {}
    try:
        for precond in ("ssor", "ssor_level"):
            cof.settings["precond"] = precond
            a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            info = cof.last_solve_info
            assert info.converged and info.relres.max() <= RES_TOL
            out[precond] = (np.array(V_k), int(info.iterations.max()))
            del a2
    finally:
        cof.settings["precond"] = old
    (Va, ita), (Vb, itb) = out["ssor"], out["ssor_level"]
    assert max(rel_l2(Vb[k], Va[k]) for k in range(T - 1)) <= 1e-9
    assert itb < 0.6 * ita, (ita, itb)


@pytest.mark.parametrize("stage_blocks", ["3", "4"])
@pytest.mark.parametrize("level,frames", [(3, 5), (5, 70), (6, 33)])
def test_persistent_level_kernel_bit_identical_to_per_level_launches(level, frames, stage_blocks):
    """The persistent cooperative kernel (row-level dataflow, bulk-async prefetch) and the one-launch-per-level
    path run the same per-row arithmetic and the same reduction order: fields, iteration counts and residuals
    must be identical bit for bit (cof:147 replacement, default path vs its fallback)."""
    import os
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.pial_like(level)
    T = frames + 1
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=3)
    old = cof.settings["precond"]
    cof.settings["precond"] = "ssor_level"
    out = {}
    os.environ["MOF_LEVEL_STAGE_BLOCKS"] = stage_blocks        # both stage configurations of the sweeps (<3,3> and <4,2>)
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        assert a2.d_level_desc is not None and a2.level_stage_blocks == int(stage_blocks)
        for persist in ("1", "0"):
            os.environ["MOF_LEVEL_PERSIST"] = persist
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            info = cof.last_solve_info
            assert info.converged and info.relres.max() <= RES_TOL
            out[persist] = (np.array(V_k), info.iterations.copy(), info.relres.copy())
            want = _lib.PATH_LEVEL_PERSISTENT if persist == "1" else _lib.PATH_LEVEL_LAUNCHES
            assert info.path[0] == want, info.path          # no silent fallback
    finally:
        os.environ.pop("MOF_LEVEL_PERSIST", None)
        os.environ.pop("MOF_LEVEL_STAGE_BLOCKS", None)
        cof.settings["precond"] = old
    assert np.array_equal(out["1"][1], out["0"][1])
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.array_equal(out["1"][2], out["0"][2])


def test_config1_exactly_as_configured():
    """BASELINE.json configs[0]: icosphere level 5 (10,242 vertices), 64-frame travelling wave -> 63 solves, every
    one against the reference algorithm's direct solve (oracle: P1 assembly + SuperLU) and detection on EVERY
    frame against the oracle's lstsq decisions (exact index lists)."""
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof, find_singularity_point as fsp
    coords, tris, normals, areas = synthetic.icosphere(5)
    T = 64
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=0)
    old = cof.settings["precond"]
    cof.settings["precond"] = DEFAULT_PRECOND                # whatever the module fixture left active
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    finally:
        cof.settings["precond"] = old
    info = cof.last_solve_info
    assert len(V_k) == T - 1 and info.converged and info.relres.max() <= RES_TOL
    assert info.path[0] == _lib.PATH_LEVEL_PERSISTENT
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    worst = 0.0
    for k in range(T - 1):
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        worst = max(worst, rel_l2(V_k[k], Vo))
    assert worst <= V_TOL, worst
    Vx = fsp.process_V_k(V_k, e)
    assert np.array_equal(Vx, mof_oracle.process_V_k(V_k, e))
    s = fsp.detect_singularities(Vx, coords, tris, 1e-4)
    for k in range(T - 1):
        vio, fio, lmo, Po, vmaxo = mof_oracle.find_singularity_points(coords, tris, Vx[k], 1e-4)
        vi, fi, lm, P, idx = s.frame(k)
        assert np.array_equal(vi, vio) and np.array_equal(fi, fio) and s.v_length_max[k] == vmaxo, k
        assert np.allclose(lm, lmo, rtol=0, atol=1e-10)


def test_full_batch_every_group_against_oracle_system():
    """A full 1,024-frame batch at config 2 size (32 groups of 32 lanes in one launch of the persistent kernel): one
    lane of EVERY group, a different lane in each, is checked against the ORACLE's independently assembled system
    (true residual), so lanes beyond the first group are covered at size."""
    import torch
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.pial_like(7)
    T = 1025
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=0)
    old = cof.settings["precond"]
    cof.settings["precond"] = DEFAULT_PRECOND
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        I_dev = torch.from_numpy(I).to(a2.device)
        V_dev, info = cof.solve_on_device(a2, I_dev, I_dev, t_k, 0.01, 0, T - 1)
    finally:
        cof.settings["precond"] = old
    assert info.converged and info.relres.max() <= RES_TOL and info.path[0] == _lib.PATH_LEVEL_PERSISTENT
    assert info.iterations.max() <= 160                      # 143 mean / 149 max at omega = 1.9
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    for g in range(32):
        k = 32 * g + (5 * g + 3) % 32
        a1, f = mof_oracle.assemble_frame(gwo, eo, into, tris, areas, t_k[k + 1] - t_k[k], I[k], I[k + 1])
        A = mof_oracle.system_matrix(a1, a2o, 0.01)
        Vk = V_dev[k].cpu().numpy()
        assert np.linalg.norm(A @ Vk - f) / np.linalg.norm(f) <= 1e-11, (g, k)
    del V_dev, I_dev
    torch.cuda.empty_cache()


def test_level_probe_mode_changes_nothing(tmp_path):
    """MOF_LEVEL_PROBE=1 selects the instrumented build of the persistent kernel (cycle counters per item segment on
    stderr, per-row finish times to MOF_LEVEL_PROBE_FILE): a development aid reachable from the environment, so it is
    run here -- same bits as the plain kernel, and the timeline file has one stamp per row and sweep direction."""
    import os
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.pial_like(4)
    T = 40
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=7)
    old = cof.settings["precond"]
    cof.settings["precond"] = DEFAULT_PRECOND
    path = str(tmp_path / "rows.bin")
    out = {}
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        for probe in ("0", "1"):
            os.environ["MOF_LEVEL_PROBE"] = probe
            os.environ["MOF_LEVEL_PROBE_FILE"] = path
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            assert cof.last_solve_info.path[0] == _lib.PATH_LEVEL_PERSISTENT
            out[probe] = (np.array(V_k), cof.last_solve_info.iterations.copy())
    finally:
        os.environ.pop("MOF_LEVEL_PROBE", None)
        os.environ.pop("MOF_LEVEL_PROBE_FILE", None)
        cof.settings["precond"] = old
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    stamps = np.fromfile(path, dtype=np.uint64)
    assert stamps.shape == (2 * len(coords),) and np.all(stamps > 0)


@pytest.mark.parametrize("signal", ["wave", "phase"])
def test_threshold_calibration_saves_a_verification_round(signal):
    """After the first check interval the SSOR paths measure every frame's true / recurrence residual ratio once and
    set its threshold from it (init_kernel MODE_CALIBRATE); MOF_CALIBRATE=0 is round 1's behaviour (freeze on the
    recurrence residual, verify, rescale, resume).  Both must meet the true-residual tolerance and agree to the parity
    tolerance; with the calibration no frame may need more kernel launches than without."""
    import os
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.icosphere(5)
    T = 65
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=1) if signal == "wave" else synthetic.wrapped_phase(coords, t_k, seed=1)
    old = cof.settings["precond"]
    cof.settings["precond"] = DEFAULT_PRECOND
    out = {}
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        s = cof._solver(a2)
        for cal in ("1", "0"):
            os.environ["MOF_CALIBRATE"] = cal
            s.profile = _lib.PcgProfile()
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            info = cof.last_solve_info
            assert info.converged and info.relres.max() <= RES_TOL
            out[cal] = (np.array(V_k), int(s.profile.launches_total), float(info.iterations.mean()))
            s.profile = None
    finally:
        os.environ.pop("MOF_CALIBRATE", None)
        cof.settings["precond"] = old
    assert max(rel_l2(out["1"][0][k], out["0"][0][k]) for k in range(T - 1)) <= 1e-9
    assert out["1"][1] <= out["0"][1] + 3, (out["1"][1:], out["0"][1:])      # + the three kernels of the calibration itself
    assert out["1"][2] <= out["0"][2] * 1.05


@pytest.mark.parametrize("valence,want", [(20, "persistent"), (40, "launches")])
def test_high_valence_rows_in_the_level_path(valence, want):
    """A vertex with many neighbours: up to 32 blocks on one side of the diagonal the persistent kernel handles the
    row (four blocks staged by the bulk copy, the next two from the row descriptor, the rest through the column
    array); beyond that mof_level_desc_build declines and the solver falls back to one launch per level.  Either way
    the fields match the oracle and the path taken is the one expected (no silent fallback, no silent misuse)."""
    from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof
    coords, tris, normals, areas = synthetic.fan_mesh(valence, 3)
    T = 5
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=1)
    old = cof.settings["precond"]
    cof.settings["precond"] = DEFAULT_PRECOND
    try:
        a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
    finally:
        cof.settings["precond"] = old
    info = cof.last_solve_info
    assert (a2.d_level_desc is not None) == (want == "persistent")
    assert info.path[0] == (_lib.PATH_LEVEL_PERSISTENT if want == "persistent" else _lib.PATH_LEVEL_LAUNCHES), info.path
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    for k in range(T - 1):
        Vo = mof_oracle.worker(k, a2o, gwo, eo, into, tris, t_k, areas, 0.01, I[k], I[k + 1])
        assert rel_l2(V_k[k], Vo) <= V_TOL


@pytest.mark.parametrize("n_frames", [5, 129, 300])
def test_streamed_upload_is_bit_identical(n_frames, mods):
    """settings['streamed_upload']: the host signal arrives in chunks on a side stream and pack + K1 run chunk by chunk as
    the rows land (the copy hides behind the assembly).  Same kernels on the same data: fields identical bit for bit to the
    blocking upload, for a single ragged group, a chunk boundary and several chunks."""
    cof, _ = mods
    coords, tris, normals, areas = synthetic.icosphere(3)
    T = n_frames + 1
    t_k = synthetic.time_axis(T, 512.0)
    I = synthetic.travelling_wave(coords, t_k, seed=11)
    a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
    out = {}
    old = cof.settings["streamed_upload"]
    try:
        for mode in (True, False):
            cof.settings["streamed_upload"] = mode
            V_k, _ = cof.compute_velocity_field(1, T, a2, gw, e, integ, tris, t_k, areas, 0.01, I, I)
            out[mode] = np.array(V_k)
    finally:
        cof.settings["streamed_upload"] = old
    assert np.array_equal(out[True], out[False])

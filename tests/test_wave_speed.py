"""S5 wave-speed row ("next" row 1 of SURVEY 8f): oracle vs outputs of the unmodified
S5_compute_wave_v.py (tests/golden/s5_*.npz, made by tests/golden/make_golden_s5.py) on CPU,
and the CUDA kernel vs the same goldens / the oracle on the GPU box."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from manifold_based_optical_flow_method_b200 import synthetic
from oracle import mof_oracle

CASES = ["s5_ico2", "s5_patch7"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_s5(case):
    g = load_golden(case)
    dt = float(g["dt"])
    assert np.allclose(mof_oracle.temporal_gradient_phase(g["phases"], dt), g["temporal_gradient_phase"], rtol=1e-13, atol=1e-9)
    gp = mof_oracle.grad_M_I(g["coordinates"], g["triangles"], g["phases"], g["areas"])
    assert rel_l2(gp, g["grad_point"]) <= 1e-14
    wp = mof_oracle.wave_velocity(g["coordinates"], g["triangles"], g["areas"], g["phases"], dt, g["e"], phase=True)
    assert rel_l2(wp, g["wave_velocity_phase"]) <= 1e-12
    wa = mof_oracle.wave_velocity(g["coordinates"], g["triangles"], g["areas"], g["potentials"], dt, g["e"], phase=False)
    assert rel_l2(wa, g["wave_velocity_amplitude"]) <= 1e-12
    # the wrapped derivative really wraps in this fixture
    raw = np.abs(np.diff(g["phases"], axis=0)).max()
    assert raw > np.pi


def test_surface_mesh_stub_point_cells():
    coords, tris, normals, areas = synthetic.icosphere(1)
    s = synthetic.SurfaceMesh(coords, tris, normals, areas)
    assert np.array_equal(s.faces.reshape(-1, 4)[:, 1:], tris)
    for i in (0, 5, 41):
        ids = s.point_cell_ids(i)
        assert ids == sorted(ids) and all(i in tris[c] for c in ids)
        assert len(ids) == int(np.sum(tris == i))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_wave_speed_matches_reference_s5(case):
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    g = load_golden(case)
    dt = float(g["dt"])
    surf = synthetic.SurfaceMesh(g["coordinates"], g["triangles"], g["normals"], g["areas"])
    T = len(g["phases"])
    wp = s5.wave_velocity_phase(surf, g["phases"], dt, T, g["e"])
    assert wp.shape == g["wave_velocity_phase"].shape
    assert rel_l2(wp, g["wave_velocity_phase"]) <= 1e-12
    wa = s5.wave_velocity_amplitude(surf, g["potentials"], dt, T, g["e"])
    assert rel_l2(wa, g["wave_velocity_amplitude"]) <= 1e-12
    gp = s5.compute_grad_M_I(g["coordinates"], g["triangles"], g["phases"], surf, g["areas"])
    assert rel_l2(gp, g["grad_point"]) <= 1e-13
    assert np.allclose(s5.compute_temporal_gradient_phase(g["phases"], dt), g["temporal_gradient_phase"], rtol=1e-13, atol=1e-9)


@pytest.mark.gpu
def test_wave_speed_matches_oracle_medium():
    """BASELINE.json configs[4]-style input at a size the oracle handles in seconds."""
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    coords, tris, normals, areas = synthetic.pial_like(4)
    T, SF = 40, 512.0
    t_k = synthetic.time_axis(T, SF)
    phases = synthetic.wrapped_phase(coords, t_k, seed=4, omega=500.0)
    e = mof_oracle.orthonormal_basis(normals)
    surf = synthetic.SurfaceMesh(coords, tris, normals, areas)
    w = s5.wave_velocity_phase(surf, phases, 1 / SF, T, e)
    wo = mof_oracle.wave_velocity(coords, tris, areas, phases, 1 / SF, e, phase=True)
    assert rel_l2(w, wo) <= 1e-12


def test_halo_rows_cover_the_time_derivative():
    """Host logic of the frame sharding (CPU): the rows a shard is given contain every frame its time
    derivative touches (central difference; one-sided second-order ends in amplitude mode, S5:24)."""
    from manifold_based_optical_flow_method_b200.S5_compute_wave_v import halo_rows
    from manifold_based_optical_flow_method_b200.distributed import shard_range
    for T in (1, 2, 3, 5, 33, 100):
        for world in (1, 2, 4, 8):
            for phase in (True, False):
                if not phase and T < 3:
                    continue
                for r in range(world):
                    k0, k1 = shard_range(T, world, r)
                    a, b = halo_rows(k0, k1, T, phase)
                    assert 0 <= a <= k0 and k1 <= b <= T
                    for t in range(k0, k1):
                        need = {t}
                        if phase:
                            need |= {t + 1} if t == 0 and T > 1 else ({t - 1} if t == T - 1 and T > 1 else {t - 1, t + 1} if T > 1 else set())
                        else:
                            need |= {1, 2} if t == 0 else ({T - 2, T - 3} if t == T - 1 else {t - 1, t + 1})
                        assert all(a <= q < b for q in need), (T, world, r, phase, t)


@pytest.mark.gpu
@pytest.mark.parametrize("phase", [True, False])
def test_wave_speed_shards_equal_whole_trial(phase):
    """Config 5 shards frames over GPUs: a shard computed from its rows plus the halo must be bit-identical to
    the same frames of the whole-trial call (ragged shard sizes, group boundaries at 32 frames)."""
    import torch
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    from manifold_based_optical_flow_method_b200.distributed import shard_range
    coords, tris, normals, areas = synthetic.pial_like(3)
    T, SF = 70, 512.0
    t_k = synthetic.time_axis(T, SF)
    data = synthetic.wrapped_phase(coords, t_k, seed=2, omega=300.0) if phase else synthetic.travelling_wave(coords, t_k, seed=2)
    e = mof_oracle.orthonormal_basis(normals)
    op = s5._operator(coords, tris, areas, e)
    d = torch.from_numpy(np.ascontiguousarray(data)).to(op.device)
    _, whole = s5.wave_speed_device(op, d, 0, T, 0, T, 1 / SF, phase)
    wo = mof_oracle.wave_velocity(coords, tris, areas, data, 1 / SF, e, phase=phase)
    assert rel_l2(whole.cpu().numpy(), wo) <= 1e-12
    for world in (2, 3, 8):
        for r in range(world):
            k0, k1 = shard_range(T, world, r)
            a, b = s5.halo_rows(k0, k1, T, phase)
            _, part = s5.wave_speed_device(op, d[a:b], a, T, k0 - a, k1 - k0, 1 / SF, phase)
            assert torch.equal(part, whole[k0:k1]), (world, r)


@pytest.mark.gpu
def test_wave_speed_full_size_config5():
    """BASELINE.json configs[4] size: 163,842 vertices.  The oracle (numpy, vectorised over frames) handles a
    few frames at this size in seconds; gradient and wave speed must match it."""
    import torch
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    coords, tris, normals, areas = synthetic.pial_like(7)
    T, SF = 6, 512.0
    t_k = synthetic.time_axis(T, SF)
    phases = synthetic.wrapped_phase(coords, t_k, seed=1, omega=500.0)
    e = mof_oracle.orthonormal_basis(normals)
    op = s5._operator(coords, tris, areas, e)
    d = torch.from_numpy(np.ascontiguousarray(phases)).to(op.device)
    grad, wave = s5.wave_speed_device(op, d, 0, T, 0, T, 1 / SF, True, want_grad=True)
    wo = mof_oracle.wave_velocity(coords, tris, areas, phases, 1 / SF, e, phase=True)
    go = mof_oracle.grad_M_I(coords, tris, phases, areas)
    assert rel_l2(grad.cpu().numpy(), go) <= 1e-13
    assert rel_l2(wave.cpu().numpy(), wo) <= 1e-12


def _host_wave(case, reorder):
    """The K6 kernel bodies (csrc/mof_bodies.h) in the plain loops of tests/hostcheck, CPU only."""
    import ctypes
    from test_host_logic import HostMesh
    g = load_golden(case)
    hm = HostMesh(g, reorder)
    P = hm.P
    hm.e[:] = np.asarray(g["e"], dtype=np.float64).reshape(-1, 2, 3)[P.perm]      # S5 is handed e by its caller
    cw, cg = np.zeros((P.n_blocks, 2)), np.zeros((P.n_blocks, 3))
    ms = hm.struct()
    hm.hc.hc_wave_coef(ctypes.byref(ms), cw.ctypes.data, cg.ctypes.data)

    def rows(data, phase, want_grad=False, a=0, b=None, k0=0, k1=None):
        """frames [k0, k1) of the trial `data` from its rows [a, b)"""
        T, N = data.shape
        b, k1 = T if b is None else b, T if k1 is None else k1
        part = np.ascontiguousarray(data[a:b], dtype=np.float64)
        out = np.full((k1 - k0, N, 3) if want_grad else (k1 - k0, N), np.nan)
        coef = cg if want_grad else cw
        hm.hc.hc_wave_rows(ctypes.byref(ms), 3 if want_grad else 2, b - a, k0 - a, k1 - k0, a, T, part.ctypes.data, N,
                           ctypes.c_double(float(g["dt"])), 1 if phase else 0, coef.ctypes.data, out.ctypes.data)
        return out
    return g, rows


@pytest.mark.parametrize("reorder", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_wave_bodies_match_reference_s5(case, reorder):
    """Coefficient rows + row products (the formulation of csrc/wave.cu) against the outputs of the unmodified S5."""
    g, rows = _host_wave(case, reorder)
    assert rel_l2(rows(g["phases"], True), g["wave_velocity_phase"]) <= 1e-12
    assert rel_l2(rows(g["potentials"], False), g["wave_velocity_amplitude"]) <= 1e-12
    assert rel_l2(rows(g["phases"], True, want_grad=True), g["grad_point"]) <= 1e-13


@pytest.mark.parametrize("phase", [True, False])
def test_wave_bodies_shards_equal_whole_trial(phase):
    """A shard computed from its rows plus the halo is bit-identical to the same frames of the whole trial."""
    from manifold_based_optical_flow_method_b200.S5_compute_wave_v import halo_rows
    from manifold_based_optical_flow_method_b200.distributed import shard_range
    g, rows = _host_wave("s5_patch7", 0)
    data = g["phases"] if phase else g["potentials"]
    T = len(data)
    whole = rows(data, phase)
    for world in (2, 3, T):
        for r in range(world):
            k0, k1 = shard_range(T, world, r)
            if k1 == k0:
                continue
            a, b = halo_rows(k0, k1, T, phase)
            assert np.array_equal(rows(data, phase, a=a, b=b, k0=k0, k1=k1), whole[k0:k1], equal_nan=True), (world, r)


def test_angle_subtract_body_is_numpy_mod():
    """The range branches of mof_angle_subtract_body are exactly np.mod(f1 - f2 + pi, 2 pi) - pi (S5:230), also for
    arguments outside [-pi, pi] (a two-frame phase trial through the harness: wave = td / 1 on a flat unit mesh is
    overkill, so the formula is restated here and compared on the values the kernel branches on)."""
    from oracle import mof_oracle
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-np.pi, np.pi, 4000), rng.uniform(-30, 30, 4000), [np.pi, -np.pi, 0.0, 2 * np.pi, 3 * np.pi]])
    b = np.concatenate([rng.uniform(-np.pi, np.pi, 4000), rng.uniform(-30, 30, 4000), [-np.pi, np.pi, 0.0, -2 * np.pi, -np.pi]])
    d = a - b + np.pi
    two_pi = 2 * np.pi
    m = np.where((d >= 0) & (d < two_pi), d, np.where((d < 0) & (d > -two_pi), d + two_pi,
                 np.where((d >= two_pi) & (d < 2 * two_pi), d - two_pi, np.mod(d, two_pi))))
    assert np.array_equal(m - np.pi, mof_oracle.angle_subtract(a, b))


def test_wave_c_abi_argument_errors():
    """mof_wave_speed / mof_wave_stencil / mof_wave_set_variant validate their arguments before anything is launched
    (CPU: the calls return an error code and a message, no device is touched)."""
    import ctypes
    from manifold_based_optical_flow_method_b200 import _lib
    lib = _lib.load()
    ms = _lib.MeshDev()
    ms.n_vertices, ms.n_faces, ms.n_blocks, ms.n_contrib = 10, 16, 58, 144
    P = ctypes.c_void_p
    buf = (ctypes.c_double * 16)()
    ptr = ctypes.cast(buf, P)
    M = ctypes.byref(ms)
    # scratch size: the packed signal (G N 32), its time halo (G N 2), the padded columns (N 8 int32) and, per output,
    # CSR-aligned + padded coefficient rows
    G, N, nb = 2, 10, 58
    base = G * N * 32 + G * N * 2 + N * 8 // 2
    assert lib.mof_wave_work_doubles(M, 40, 0, 1) == base + 2 * nb + N * 8 * 2
    assert lib.mof_wave_work_doubles(M, 40, 1, 0) == base + 3 * nb + N * 8 * 4
    assert lib.mof_wave_work_doubles(None, 40, 0, 1) == -1

    def speed(n_rows=8, out0=0, n_out=8, t_first=0, T=8, I=ptr, ld=10, dt=0.5, phase=1, grad=None, wave=ptr, work=ptr):
        return lib.mof_wave_speed(M, n_rows, out0, n_out, t_first, T, I, ld, dt, phase, grad, wave, work, None)
    for kw, msg in ((dict(I=None), "bad arguments"), (dict(ld=9), "bad arguments"), (dict(dt=0.0), "bad arguments"),
                    (dict(wave=None), "nothing to compute"), (dict(out0=4, n_out=5), "output rows outside"),
                    (dict(t_first=1), "rows outside the trial"), (dict(phase=0, T=2, n_rows=2, n_out=2), "at least 3 frames"),
                    (dict(n_rows=4, n_out=4, t_first=2, T=8), "lack the time-derivative halo"),
                    (dict(phase=0, n_rows=5, out0=0, n_out=5, t_first=3, T=8), "lack the time-derivative halo")):
        assert speed(**kw) == -1, kw
        assert msg in _lib.last_error(), (kw, _lib.last_error())
    assert speed(work=ctypes.c_void_p(ptr.value + 8)) == -1 and "16-byte aligned" in _lib.last_error()
    assert speed(n_out=0) == 0                       # nothing asked for: nothing launched
    assert lib.mof_wave_stencil(M, 0, 0, 0, 0, 8, 0.5, 1, None, ptr, ptr, None) == -1
    assert lib.mof_wave_stencil(M, 8, 0, 8, 0, 8, 0.5, 1, None, None, ptr, None) == -1
    keep = lib.mof_wave_get_variant()
    assert 0 <= keep <= 6
    assert lib.mof_wave_set_variant(7) == -1 and lib.mof_wave_set_variant(-1) == -1
    assert lib.mof_wave_set_variant(0) == 0 and lib.mof_wave_get_variant() == 0
    assert lib.mof_wave_set_variant(keep) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("reorder", [1, 3])
def test_wave_speed_on_a_renumbered_mesh(reorder):
    """The C ABI takes any mesh handle: with a renumbered mesh (Cuthill-McKee, level-scheduled) the transposes gather /
    scatter through mesh->perm and a row's products are summed in another column order -- same result to rounding."""
    import torch
    from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5
    from manifold_based_optical_flow_method_b200.mesh import MeshOperator
    coords, tris, normals, areas = synthetic.pial_like(3)
    T, SF = 40, 512.0
    t_k = synthetic.time_axis(T, SF)
    e = mof_oracle.orthonormal_basis(normals)
    op = MeshOperator(coords, normals, tris, areas, reorder=reorder)
    op.use_geometry(None, e, None, areas)
    for phase in (True, False):
        data = synthetic.wrapped_phase(coords, t_k, seed=5, omega=300.0) if phase else synthetic.travelling_wave(coords, t_k, seed=5)
        d = torch.from_numpy(np.ascontiguousarray(data)).to(op.device)
        grad, wave = s5.wave_speed_device(op, d, 0, T, 0, T, 1 / SF, phase, want_grad=True)
        assert rel_l2(wave.cpu().numpy(), mof_oracle.wave_velocity(coords, tris, areas, data, 1 / SF, e, phase=phase)) <= 1e-12
        assert rel_l2(grad.cpu().numpy(), mof_oracle.grad_M_I(coords, tris, data, areas)) <= 1e-13

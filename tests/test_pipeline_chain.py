"""The reference's scripts chained on one synthetic recording, every stage through the drop-in modules and
checked against the oracle: S2 (electrodes -> surface, potentials and phases) -> CSV -> S3 (velocity fields,
critical points + classification) -> S5 (phase wave speed) -> S7 (multi-ring winding numbers) -> pkl.bz2."""
import numpy as np
import pytest

from conftest import rel_l2
from manifold_based_optical_flow_method_b200 import synthetic
from oracle import mof_oracle

pytestmark = pytest.mark.gpu


def test_s2_s3_s5_s7_chain(tmp_path):
    from manifold_based_optical_flow_method_b200 import (S2_interpolate, S2_interpolate_phases, S5_compute_wave_v,
                                                         S7_winding_line, compute_optical_flow, find_singularity_point,
                                                         pickle_io)
    rng = np.random.default_rng(11)
    coords, tris, normals, areas = synthetic.pial_like(4, radius=80.0, seed=1)
    surface = synthetic.SurfaceMesh(coords, tris, normals, areas)
    N, SF, T, lambda_, eps = len(coords), 512.0, 10, 0.01, 1e-4
    sel = rng.choice(N, 72, replace=False)
    electrodes = coords[sel] + rng.normal(0, 0.2, (72, 3))
    t_k = [i / SF for i in range(T)]
    recording = synthetic.travelling_wave(coords, np.asarray(t_k), seed=4)[:, sel]          # (T, electrodes)

    # ---- S2: potentials and phases on the surface, saved the way S3 / S5 read them
    pot_path, ph_path = str(tmp_path / "interpolation_data.csv"), str(tmp_path / "interpolation_phases_data.csv")
    pot = S2_interpolate.interpolation(surface, recording, electrodes, 0, T, pot_path, True)
    z = np.exp(1j * S2_interpolate_phases.compute_phase_from_potentials(recording))
    ph = S2_interpolate_phases.interpolation(surface, z, electrodes, 0, T, ph_path, True)
    pot_o = mof_oracle.rbf_interpolate(electrodes, recording, coords)
    ph_o = mof_oracle.rbf_interpolate(electrodes, np.exp(1j * mof_oracle.electrode_phases(recording)), coords, phase=True)
    assert np.abs(pot - pot_o).max() <= 1e-10 * np.abs(pot_o).max()
    assert np.abs(np.angle(np.exp(1j * (ph - ph_o)))).max() <= 1e-8

    # ---- S3: velocity fields of the potentials read back from disk, critical points, classes
    I_k = compute_optical_flow.load_potentials(pot_path)
    assert np.array_equal(I_k, pot)
    a2, grad_w, e, integral, _ = compute_optical_flow.compute_geometrical_quantities(coords, normals, tris, areas)
    V_k, _ = compute_optical_flow.compute_velocity_field(8, T, a2, grad_w, e, integral, tris, t_k, areas, lambda_, I_k, I_k)
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    Vo, _ = mof_oracle.compute_velocity_field(1, T, a2o, gwo, eo, into, tris, t_k, areas, lambda_, I_k, I_k)
    assert max(rel_l2(V_k[k], Vo[k]) for k in range(T - 1)) <= 1e-8
    V_k_coord = np.array(find_singularity_point.process_V_k(V_k, e))
    points, kinds = find_singularity_point.find_singularity_points_and_classify_for_all_Vk(V_k_coord, coords, tris, eps, surface, e)
    Vxo = mof_oracle.process_V_k(Vo, eo)
    for k in range(T - 1):
        pts_o, codes_o, jac_o = mof_oracle.classify_singularities(coords, tris, Vxo[k], eps, eo)
        assert len(points[k]) == len(pts_o)
        if len(pts_o):
            assert np.allclose(np.asarray(points[k]), pts_o, atol=1e-5)
            ok = np.isfinite(jac_o).all(axis=(1, 2))
            assert [kinds[k][q] for q in np.nonzero(ok)[0]] == [mof_oracle.CLASS_NAMES[c] for c in codes_o[ok]]

    # ---- S7: winding numbers around the detected points, frame by frame like the reference's loop
    for k in (0, T - 2):
        counts, types = S7_winding_line.calculate_winding_numbers(surface, points[k], V_k_coord[k], e, coords)
        oc, ot, _ = mof_oracle.winding_numbers(coords, tris, points[k], V_k_coord[k], e)
        assert counts == [int(c) for c in oc] and types == [int(t) for t in ot if t != 0]

    # ---- S5: phase wave speed from the phases read back from disk
    phases = compute_optical_flow.load_potentials(ph_path)
    wave = S5_compute_wave_v.wave_velocity_phase(surface, phases, 1.0 / SF, T, e)
    wave_o = mof_oracle.wave_velocity(coords, tris, areas, phases, 1.0 / SF, eo, phase=True)
    finite = np.isfinite(wave_o)
    assert np.array_equal(np.isfinite(wave), finite)
    assert np.abs(wave[finite] - wave_o[finite]).max() <= 1e-9 * np.abs(wave_o[finite]).max()

    # ---- outputs the later scripts pick up
    V_c = np.sqrt(np.sum(V_k_coord[:, :, :3] ** 2, axis=2))
    pickle_io.dump(V_c, str(tmp_path / "wave_velocity_opticalflow.pkl.bz2"))
    assert np.array_equal(pickle_io.load(str(tmp_path / "wave_velocity_opticalflow.pkl.bz2")), V_c)

"""The body of S3_compute_v_and_detection_singularity.py (:75-137, plus its commented-out detection
call :139-144) run against the drop-in modules on a synthetic mesh: only the import line and the
pyvista surface differ from the reference script.  Files on disk are compared with what the
reference's own pandas calls produce for the oracle's fields."""
import bz2
import pickle

import numpy as np
import pandas as pd
import pytest

from conftest import rel_l2
from manifold_based_optical_flow_method_b200 import synthetic
from oracle import mof_oracle

pytestmark = pytest.mark.gpu


def test_s3_script_body(tmp_path):
    # --- what S2 would have left on disk: the interpolated potentials CSV (index column + header)
    coords, tris, normals, areas = synthetic.pial_like(3, radius=80.0, seed=2)
    surface = synthetic.SurfaceMesh(coords, tris, normals, areas)
    SF, T = 512.0, 9
    pot = synthetic.travelling_wave(coords, synthetic.time_axis(T, SF), seed=7)
    potentials_path = tmp_path / "ave-interpolation_data.csv"
    pd.DataFrame(pot).to_csv(potentials_path)

    # --- S3 body, reference variable names; `from utils import ...` replaced by the drop-in import
    from manifold_based_optical_flow_method_b200 import compute_optical_flow, find_singularity_point
    lambda_, eps, processes_num = 0.01, 1e-4, 32                      # config.yaml:3-6
    potentials = compute_optical_flow.load_potentials(potentials_path)             # S3:76
    coordinates = surface.points                                                    # S3:79
    triangles = surface.faces.reshape(-1, 4)[:, 1:]                                 # S3:80
    normals_ = surface.point_normals                                                # S3:83
    areas_ = surface.compute_cell_sizes(length=False, volume=False)['Area']         # S3:84
    time_steps = len(potentials)                                                    # S3:86
    t_k = [i / SF for i in range(time_steps)]                                       # S3:87
    t_k_ = [i for i in range(time_steps)]
    I_k = potentials[t_k_]                                                          # S3:89
    a2, grad_w, e, integral_wi_wj, execution_time = compute_optical_flow.compute_geometrical_quantities(
        coordinates, normals_, triangles, areas_)                                   # S3:97
    V_k, execution_time = compute_optical_flow.compute_velocity_field(
        processes_num, time_steps, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas_, lambda_, I_k, I_k)   # S3:104
    e_path, V_k_path = tmp_path / "e.csv", tmp_path / "V_k.csv"
    compute_optical_flow.reshape_and_save_data(e, e_path)                           # S3:122
    compute_optical_flow.reshape_and_save_data(V_k, V_k_path)                       # S3:123
    V_k_coord = find_singularity_point.process_V_k(V_k, e)                          # S3:128
    V_k_coord = np.array(V_k_coord)                                                 # S3:130
    V_c = np.sqrt(np.sum(V_k_coord[:, :, :3] ** 2, axis=2))                         # S3:132
    sl_fname = tmp_path / "wave_velocity_opticalflow.pkl.bz2"
    with bz2.BZ2File(sl_fname, 'wb') as file:                                       # S3:136-137
        pickle.dump(V_c, file)
    singularity_points = find_singularity_point.find_singularity_points_for_all_Vk(V_k_coord, coordinates, triangles, eps)   # S3:139

    # --- the same through the oracle
    a2o, gwo, eo, into = mof_oracle.geometrical_quantities(coords, normals, tris, areas)
    Vo, _ = mof_oracle.compute_velocity_field(1, T, a2o, gwo, eo, into, tris, t_k, areas, lambda_, pot, pot)
    assert np.array_equal(potentials, pot)                                          # CSV round trip is exact
    assert len(V_k) == T - 1 and max(rel_l2(V_k[k], Vo[k]) for k in range(T - 1)) <= 1e-8
    # files: e.csv is byte-identical to what pandas writes; V_k.csv parses back to the returned fields
    pd.DataFrame(eo.reshape(len(eo), -1)).to_csv(tmp_path / "e_ref.csv")
    assert e_path.read_bytes() == (tmp_path / "e_ref.csv").read_bytes()
    back = pd.read_csv(V_k_path, index_col=0, float_precision="round_trip").values
    assert back.shape == (T - 1, 2 * len(coords)) and np.array_equal(back, np.array(V_k))
    with bz2.BZ2File(sl_fname, 'rb') as file:
        assert np.allclose(pickle.load(file), mof_oracle.speed_magnitude(mof_oracle.process_V_k(Vo, eo)), rtol=1e-7)
    ref_points = mof_oracle.find_singularity_points_for_all_Vk(mof_oracle.process_V_k(Vo, eo), coords, tris, eps)
    assert [len(f) for f in singularity_points] == [len(f) for f in ref_points]
    for a, b in zip(singularity_points, ref_points):
        assert np.allclose(np.asarray(a).reshape(-1, 3), np.asarray(b).reshape(-1, 3), atol=1e-5)

"""Jacobian classification of critical points ("next" row 2 of SURVEY 8f): oracle vs outputs of the
unmodified reference functions (tests/golden/classify_*.npz, tests/golden/make_golden_classify.py)
on CPU; CUDA kernel vs the same goldens on the GPU box."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import mof_oracle

CASES = ["classify_ico2_wave", "classify_ico3_phase", "classify_ico2_vertex_singular"]


def _close(a, b):
    """Jacobian entries may be inf / nan when a neighbour's offset along e1 or e2 is exactly 0
    (the reference divides by it, fsp:396-399); compare those positions exactly, the rest to 1e-9."""
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b)
    return (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
            and np.allclose(a[fin], b[fin], rtol=1e-9, atol=1e-9 * np.max(np.abs(b[fin]), initial=1.0)))


@pytest.mark.parametrize("case", CASES)
def test_oracle_classification_matches_reference(case):
    g = load_golden(case)
    off = 0
    for k, V_now in enumerate(g["V"]):
        pts, codes, J = mof_oracle.classify_singularities(g["coordinates"], g["triangles"], V_now, float(g["eps"]), g["e"])
        n = int(g["counts"][k])
        assert len(pts) == n
        assert np.allclose(pts, g["points"][off:off + n], atol=1e-12)
        assert _close(J, g["jacobians"][off:off + n])
        assert np.array_equal(codes, g["codes"][off:off + n])
        off += n


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_classification_matches_reference(case):
    from manifold_based_optical_flow_method_b200 import find_singularity_point as fsp
    from manifold_based_optical_flow_method_b200 import synthetic
    g = load_golden(case)
    surf = synthetic.SurfaceMesh(g["coordinates"], g["triangles"])
    pts, cls = fsp.find_singularity_points_and_classify_for_all_Vk(g["V"], g["coordinates"], g["triangles"], float(g["eps"]), surf, g["e"])
    assert [len(p) for p in pts] == list(g["counts"]) and [len(c) for c in cls] == list(g["counts"])
    flat_pts = np.asarray([p for fr in pts for p in fr]).reshape(-1, 3)
    assert np.allclose(flat_pts, g["points"], atol=1e-9)
    res = fsp.classify_singularities(g["V"], g["coordinates"], g["triangles"], float(g["eps"]), g["e"])
    # A neighbour whose offset along e1 / e2 is exactly 0 makes the reference divide by +-0 (fsp:396-399); the
    # sign of that zero comes out of BLAS' ddot and decides between +inf and -inf, i.e. the class of such a
    # point is an accident of the reference's arithmetic.  Those points must be non-finite in the same
    # entries; everything else must agree in value and class.
    regular = np.isfinite(g["jacobians"]).all(axis=(1, 2))
    assert np.array_equal(np.isfinite(res.jacobians), np.isfinite(g["jacobians"]))
    assert _close(res.jacobians[regular], g["jacobians"][regular])
    assert np.array_equal(res.codes[regular], g["codes"][regular])
    names = np.asarray([c for fr in cls for c in fr])
    assert list(names[regular]) == [mof_oracle.CLASS_NAMES[c] for c in g["codes"][regular]]
    assert regular.sum() >= len(regular) - 1

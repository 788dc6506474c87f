"""On-disk formats ("next" row 4 of SURVEY 8f): the C++ CSV writer/reader against pandas, which is
what the reference uses (compute_optical_flow.py:203-207, :314-320)."""
import numpy as np
import pandas as pd
import pytest

from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof


def _special_values(rng, shape):
    a = rng.standard_normal(shape) * 10.0 ** rng.integers(-12, 12, shape)
    flat = a.ravel()
    flat[:24] = [0.0, -0.0, 1.0, -1.0, 5.0, 100.0, 1e15, 1e16, 9.999999999999999e15, 1e-4, 9.99e-5, 1e-5, 0.1 + 0.2,
                 1 / 3, 2.0 ** 53, 123456789012345680.0, 1e22, 1e-300, 5e-324, 1.7976931348623157e308, np.inf, -np.inf,
                 np.nan, 0.30000000000000004]
    return a


@pytest.mark.parametrize("shape", [(1, 1), (7, 30), (50, 3), (3, 4000)])
def test_writer_is_byte_identical_to_pandas(tmp_path, shape):
    rng = np.random.default_rng(0)
    a = _special_values(rng, shape) if shape[0] * shape[1] >= 24 else rng.standard_normal(shape)
    cof.reshape_and_save_data(a, tmp_path / "mine.csv")
    pd.DataFrame(a.reshape(a.shape[0], -1)).to_csv(tmp_path / "ref.csv")
    assert (tmp_path / "mine.csv").read_bytes() == (tmp_path / "ref.csv").read_bytes()


def test_writer_accepts_lists_and_higher_rank(tmp_path):
    e = np.random.default_rng(1).standard_normal((5, 2, 3))          # like the e array (N,2,3) -> (N,6)
    cof.reshape_and_save_data(e, tmp_path / "e.csv")
    pd.DataFrame(e.reshape(5, -1)).to_csv(tmp_path / "e_ref.csv")
    assert (tmp_path / "e.csv").read_bytes() == (tmp_path / "e_ref.csv").read_bytes()
    V = [np.arange(4.0), np.arange(4.0) * 0.5]                        # like V_k, a list of arrays
    cof.reshape_and_save_data(V, tmp_path / "v.csv")
    assert np.array_equal(cof.load_potentials(tmp_path / "v.csv"), np.array(V))


@pytest.mark.parametrize("shape", [(1, 1), (9, 17), (64, 300)])
def test_reader_matches_pandas(tmp_path, shape):
    rng = np.random.default_rng(2)
    a = _special_values(rng, shape) if shape[0] * shape[1] >= 24 else rng.standard_normal(shape)
    pd.DataFrame(a).to_csv(tmp_path / "x.csv")
    mine = cof.load_potentials(tmp_path / "x.csv")
    exact = pd.read_csv(tmp_path / "x.csv", sep=',', header='infer', index_col=0, float_precision="round_trip").values
    assert mine.shape == a.shape
    assert np.array_equal(mine, exact, equal_nan=True) and np.array_equal(mine, a, equal_nan=True)
    default = pd.read_csv(tmp_path / "x.csv", sep=',', header='infer', index_col=0).values   # the reference's call
    fin = np.isfinite(a) & (np.abs(a) > 1e-290) & (np.abs(a) < 1e290)
    # pandas' default converter keeps ~15 significant characters counting leading zeros (0.008414588934539998 is
    # read back as 0.0084145889345399, rel 1e-14..1e-12); the C++ reader is correctly rounded like 'round_trip'
    assert np.allclose(mine[fin], default[fin], rtol=2e-12, atol=0)


def test_reader_rejects_ragged_file(tmp_path):
    (tmp_path / "bad.csv").write_text(",0,1\n0,1.0,2.0\n1,3.0\n")
    from manifold_based_optical_flow_method_b200 import _lib
    with pytest.raises(_lib.MofError):
        cof.load_potentials(tmp_path / "bad.csv")

"""pickle + bz2 files of the reference (S3:136-137, S5:317-318): the multi-stream file
written by pickle_io.dump must read back with the reference's own idiom."""
import bz2
import pickle
import time

import numpy as np

from manifold_based_optical_flow_method_b200 import pickle_io


def test_round_trip_with_the_reference_idiom(tmp_path):
    rng = np.random.default_rng(0)
    V_c = np.abs(rng.normal(size=(37, 2562)))
    for obj, chunk in ((V_c, 1 << 16), (V_c, 8 << 20), ([[np.arange(3.0)], ["Node", "Saddle"]], 7), ({}, 8 << 20)):
        path = str(tmp_path / "x.pkl.bz2")
        n = pickle_io.dump(obj, path, chunk_bytes=chunk)
        assert n > 0
        with bz2.BZ2File(path, "rb") as f:                            # S7_winding_line.py:218-219
            back = pickle.load(f)
        if isinstance(obj, np.ndarray):
            assert np.array_equal(back, obj) and back.dtype == obj.dtype
        else:
            assert repr(back) == repr(obj)
        assert repr(pickle_io.load(path)) == repr(back)
    # the reference's own writer remains readable by load()
    with bz2.BZ2File(path, "wb") as f:                                # S3:136-137
        pickle.dump(V_c, f)
    assert np.array_equal(pickle_io.load(path), V_c)


def test_parallel_writer_is_not_slower(tmp_path):
    rng = np.random.default_rng(1)
    V_c = np.abs(rng.normal(size=(24, 163842)))                       # 31 MB
    t0 = time.time()
    pickle_io.dump(V_c, str(tmp_path / "a.pkl.bz2"), chunk_bytes=2 << 20)
    t_par = time.time() - t0
    t0 = time.time()
    with bz2.BZ2File(str(tmp_path / "b.pkl.bz2"), "wb") as f:
        pickle.dump(V_c, f)
    t_ref = time.time() - t0
    assert np.array_equal(pickle_io.load(str(tmp_path / "a.pkl.bz2")), V_c)
    print(f"pickle+bz2 of {V_c.nbytes / 1e6:.0f} MB: parallel {t_par:.2f} s, reference idiom {t_ref:.2f} s")
    assert t_par < 1.5 * t_ref

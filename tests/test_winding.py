"""Multi-ring winding numbers (S7_winding_line.py:59-165; SURVEY 8f "next" row 3).
Goldens tests/golden/s7_*.npz come from the unmodified reference (make_golden_s7.py)."""
import os

import numpy as np
import pytest

from oracle import mof_oracle as oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("ico4_wave", "ico3_phase", "pial3_wave")


def _frames(name):
    g = np.load(os.path.join(GOLD, "s7_" + name + ".npz"))
    off = 0
    for k, n in enumerate(g["npts"]):
        yield g, k, slice(off, off + int(n))
        off += int(n)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    for g, k, sl in _frames(name):
        c, t, w = oracle.winding_numbers(g["coordinates"], g["triangles"], g["points"][sl], g["V"][k], g["e"])
        assert np.array_equal(c, g["counts"][sl])
        assert np.array_equal(t, g["types"][sl])
        # every evaluated winding number is an integer up to rounding
        ev = w[np.isfinite(w)]
        assert np.max(np.abs(ev - np.round(ev))) < 1e-9


def test_oracle_rings_exhausted_and_zero_field():
    from manifold_based_optical_flow_method_b200 import synthetic
    coords, tris, normals, _ = synthetic.icosphere(1)                              # 42 vertices: 4-5 rings only
    e = oracle.orthonormal_basis(normals)
    V = np.cross([0.0, 0.0, 1.0], coords)                                 # rotation about z: +1 at both poles
    pole = coords[np.argmax(coords[:, 2])]
    c, t, w = oracle.winding_numbers(coords, tris, [pole], V, e)
    assert t[0] == 1 and 1 <= c[0] <= 6 and np.isnan(w[0, c[0]:]).all()
    c, t, w = oracle.winding_numbers(coords, tris, [pole], np.zeros_like(V), e)
    assert c[0] == 0 and t[0] == 0 and np.isnan(w[0, 0])                  # 0/0 turning angles


@pytest.mark.parametrize("name", CASES)
def test_kernel_steps_on_host(name):
    """The element / angle / acceptance bodies and the kernel's ring + rank-sort steps, run
    sequentially by the g++ harness, against the reference goldens and the oracle."""
    import ctypes
    import hostcheck
    lib = hostcheck.load()
    for g, k, sl in _frames(name):
        coords = np.ascontiguousarray(g["coordinates"], dtype=np.float64)
        V = np.ascontiguousarray(g["V"][k][:, :3], dtype=np.float64)
        e = np.ascontiguousarray(g["e"], dtype=np.float64)
        ptr, idx = oracle.one_ring(g["triangles"], len(coords))
        ptr, idx = ptr.astype(np.int32), idx.astype(np.int32)
        oc, ot, ow = oracle.winding_numbers(coords, g["triangles"], g["points"][sl], V, e)
        for q, P in enumerate(np.ascontiguousarray(g["points"][sl], dtype=np.float64)):
            closest, count, typ = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
            w = np.empty(25)
            P = np.ascontiguousarray(P)
            lib.hc_winding(len(coords), coords.ctypes.data, V.ctypes.data, e.ctypes.data, ptr.ctypes.data, idx.ctypes.data,
                           P.ctypes.data, 25, ctypes.byref(closest), ctypes.byref(count), ctypes.byref(typ), w.ctypes.data)
            assert closest.value == oracle.closest_vertex(coords, P)
            assert count.value == g["counts"][sl][q] and typ.value == g["types"][sl][q]
            assert np.array_equal(np.isnan(w), np.isnan(ow[q]))
            assert np.allclose(w[np.isfinite(w)], ow[q][np.isfinite(w)], rtol=0, atol=1e-12)


# ------------------------------------------------------------------ GPU (through the C ABI)
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference(name):
    from manifold_based_optical_flow_method_b200 import S7_winding_line as s7, synthetic
    for g, k, sl in _frames(name):
        surf = synthetic.SurfaceMesh(g["coordinates"], g["triangles"])
        counts, types = s7.calculate_winding_numbers(surf, list(g["points"][sl]), g["V"][k], g["e"], g["coordinates"])
        assert counts == [int(c) for c in g["counts"][sl]]
        assert types == [int(t) for t in g["types"][sl] if t != 0]


@pytest.mark.gpu
def test_gpu_all_frames_one_launch_matches_oracle():
    from manifold_based_optical_flow_method_b200 import S7_winding_line as s7
    g = np.load(os.path.join(GOLD, "s7_ico4_wave.npz"))
    npts = g["npts"]
    fop = np.repeat(np.arange(len(npts)), npts)
    r = s7.winding_numbers(g["triangles"], g["coordinates"], g["points"], fop, g["V"], g["e"])
    assert np.array_equal(r.counts, g["counts"]) and np.array_equal(r.types, g["types"])
    off = 0
    for k, n in enumerate(npts):
        c, t, w = oracle.winding_numbers(g["coordinates"], g["triangles"], g["points"][off:off + n], g["V"][k], g["e"])
        got = r.winding[off:off + n]
        assert np.array_equal(np.isnan(got), np.isnan(w))
        assert np.allclose(got[np.isfinite(w)], w[np.isfinite(w)], rtol=0, atol=1e-12)
        assert np.array_equal(r.closest[off:off + n], [oracle.closest_vertex(g["coordinates"], P) for P in g["points"][off:off + n]])
        off += n


@pytest.mark.gpu
def test_gpu_edge_cases():
    from manifold_based_optical_flow_method_b200 import S7_winding_line as s7, synthetic
    coords, tris, normals, _ = synthetic.icosphere(1)
    e = oracle.orthonormal_basis(normals)
    V = np.cross([0.0, 0.0, 1.0], coords)
    pole = coords[np.argmax(coords[:, 2])]
    surf = synthetic.SurfaceMesh(coords, tris)
    for field in (V, np.zeros_like(V)):
        oc, ot, ow = oracle.winding_numbers(coords, tris, [pole, coords[3] * 1.01], field, e)
        r = s7.winding_numbers(tris, coords, [pole, coords[3] * 1.01], 0, field, e)
        assert np.array_equal(r.counts, oc) and np.array_equal(r.types, ot)
        assert np.array_equal(np.isnan(r.winding), np.isnan(ow))
    assert s7.calculate_winding_numbers(surf, [], V, e, coords) == ([], [])
    # many points (every vertex of a larger mesh), two frames, max_level 3
    coords, tris, normals, _ = synthetic.icosphere(3)
    e = oracle.orthonormal_basis(normals)
    V = np.stack([np.cross([0.0, 0.0, 1.0], coords), np.cross([1.0, 0.3, 0.0], coords) + 0.2 * np.cross(coords, np.cross([0, 1.0, 0], coords))])
    pts = np.concatenate([coords, coords])
    fop = np.repeat([0, 1], len(coords))
    r = s7.winding_numbers(tris, coords, pts, fop, V, e, max_level=3)
    for k in (0, 1):
        oc, ot, ow = oracle.winding_numbers(coords, tris, coords, V[k], e, max_level=3)
        sl = slice(k * len(coords), (k + 1) * len(coords))
        assert np.array_equal(r.counts[sl], oc) and np.array_equal(r.types[sl], ot)
        assert np.array_equal(r.closest[sl], np.arange(len(coords)))

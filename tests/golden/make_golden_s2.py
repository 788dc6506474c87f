"""Golden vectors for the RBF interpolation of S2_interpolate.py / S2_interpolate_phases.py
(`interpolation`, :22-53 / :22-56): the UNMODIFIED reference functions, with pyvista.read stubbed to
return the synthetic surface and the CSV they write read back.
Build container only:  python tests/golden/make_golden_s2.py"""
import importlib.util
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, surface):
    for mod in ("pyvista", "matplotlib", "matplotlib.pyplot", "mne"):
        if mod not in sys.modules:
            try:
                __import__(mod)
            except Exception:
                sys.modules[mod] = types.ModuleType(mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyvista"].read = lambda path: surface
    spec = importlib.util.spec_from_file_location("_mof_reference_" + name, os.path.join(reference_shim.REFERENCE_ROOT, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def electrodes(coords, m, seed):
    """m electrode positions: farthest-point sample of the upper cap of the surface, jittered off the vertices"""
    rng = np.random.default_rng(seed)
    cap = np.nonzero(coords[:, 2] > 0.3 * np.abs(coords).max())[0]
    sel = [int(cap[0])]
    d = np.linalg.norm(coords[cap] - coords[sel[0]], axis=1)
    for _ in range(m - 1):
        k = int(np.argmax(d))
        sel.append(int(cap[k]))
        d = np.minimum(d, np.linalg.norm(coords[cap] - coords[cap[k]], axis=1))
    return coords[sel] + rng.normal(0, 0.004 * np.abs(coords).max(), (m, 3)), np.asarray(sel)


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    for tag, level, m, T in (("ico3_m20", 3, 20, 9), ("ico4_m61", 4, 61, 7)):
        coords, tris, normals, areas = synthetic.pial_like(level)
        surface = synthetic.SurfaceMesh(coords, tris)
        C, sel = electrodes(coords, m, seed=level)
        t_k = synthetic.time_axis(T + 4, 512.0)
        data = synthetic.travelling_wave(coords, t_k, seed=1)[:, sel]            # (T+4, m) electrode potentials
        start, end = 2, 2 + T
        out = {}
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "x.csv")
            s2 = load("S2_interpolate", surface)
            with reference_shim.quiet():
                s2.interpolation("unused.ply", data, C, start, end, path, True)
            out["potentials"] = pd.read_csv(path, index_col=0, float_precision="round_trip").values
            s2p = load("S2_interpolate_phases", surface)
            with reference_shim.quiet():
                ph = np.exp(1j * s2p.compute_phase_from_potentials(data))
                s2p.interpolation("unused.ply", ph, C, start, end, path, True)
            out["phases"] = pd.read_csv(path, index_col=0, float_precision="round_trip").values
            out["electrode_phases"] = s2p.compute_phase_from_potentials(data)
        np.savez_compressed(os.path.join(OUT, "s2_" + tag + ".npz"), coordinates=coords, electrodes=C, data=data,
                            start=start, end=end, **out)
        print(tag, out["potentials"].shape, out["phases"].shape, float(np.abs(out["potentials"]).max()))

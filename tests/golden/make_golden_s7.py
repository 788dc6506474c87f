"""Golden vectors for the multi-ring winding numbers of S7_winding_line.py (:59-165): the UNMODIFIED
calculate_winding_numbers run with synthetic.SurfaceMesh in place of the pyvista surface
(find_closest_point = nearest vertex, point_neighbors_levels = pyvista's topological rings).
Build container only:  python tests/golden/make_golden_s7.py"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_s7():
    for name in ("pyvista", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mne"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, reference_shim.REFERENCE_ROOT)          # S7 does `from utils import draw_optical_flow_field`
    spec = importlib.util.spec_from_file_location("_mof_reference_S7", os.path.join(reference_shim.REFERENCE_ROOT, "S7_winding_line.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    s7 = load_s7()
    for src, frames in (("ico4_wave", (0, 1)), ("ico3_phase", (0,)), ("pial3_wave", (0, 2))):
        g = np.load(os.path.join(OUT, src + ".npz"))
        coords, tris, e = g["coordinates"], g["triangles"], g["e"]
        surf = synthetic.SurfaceMesh(coords, tris)
        pts_all, counts_all, types_all, npts, V = [], [], [], [], []
        for k in frames:
            nf = g["sing_counts"][:k, 1].sum()
            P = g["sing_face_P"][nf:nf + g["sing_counts"][k, 1]]
            try:
                with reference_shim.quiet():
                    counts, types_ = s7.calculate_winding_numbers(surf, list(P), g["V_xyz"][k], e, coords)
            except Exception as exc:
                print(src, k, "reference raised", type(exc).__name__, exc)
                continue
            # types lists only the points whose first ring has winding +-1; expand to one entry per point
            it = iter(types_)
            full_types = [next(it) if c > 0 else 0 for c in counts]
            pts_all.append(P); counts_all += list(counts); types_all += full_types; npts.append(len(P)); V.append(g["V_xyz"][k])
            print(src, k, "points", len(P), "counts", counts[:12], "types", types_[:12])
        if npts:
            np.savez_compressed(os.path.join(OUT, "s7_" + src + ".npz"), coordinates=coords, triangles=tris, e=e, V=np.asarray(V),
                                npts=np.asarray(npts), points=np.concatenate(pts_all), counts=np.asarray(counts_all),
                                types=np.asarray(types_all))

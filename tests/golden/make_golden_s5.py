"""Golden vectors for the S5 wave-speed row: runs the UNMODIFIED reference functions of
/root/reference/S5_compute_wave_v.py (pyvista replaced by synthetic.SurfaceMesh, which offers
the four PolyData members S5 touches).  Build container only:  python tests/golden/make_golden_s5.py"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_s5():
    for name in ("pyvista", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("_mof_reference_S5", os.path.join(reference_shim.REFERENCE_ROOT, "S5_compute_wave_v.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    s5 = load_s5()
    cof, _ = reference_shim.load()
    for name, mesh in (("s5_ico2", synthetic.icosphere(2, radius=3.0)), ("s5_patch7", synthetic.open_patch(7, seed=5))):
        coords, tris, normals, areas = mesh
        surf = synthetic.SurfaceMesh(coords, tris, normals, areas)
        T, SF = 6, 250.0
        t_k = synthetic.time_axis(T, SF)
        phases = synthetic.wrapped_phase(coords, t_k, seed=2, omega=300.0)      # wraps between frames
        pots = synthetic.travelling_wave(coords, t_k, seed=2)
        e = np.zeros((len(coords), 2, 3))
        for i in range(len(coords)):
            e[i][0], e[i][1] = s5.compute_orthonormal_basis(normals[i])
        with reference_shim.quiet():
            wv_phase = s5.wave_velocity_phase(surf, phases, 1 / SF, T, e)
            wv_amp = s5.wave_velocity_amplitude(surf, pots, 1 / SF, T, e)
            gp = s5.compute_grad_M_I(coords, tris, phases, surf, areas)
            tg = s5.compute_temporal_gradient_phase(phases, 1 / SF)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), coordinates=coords, triangles=tris, normals=normals, areas=areas,
                            e=e, phases=phases, potentials=pots, dt=1 / SF, wave_velocity_phase=wv_phase,
                            wave_velocity_amplitude=wv_amp, grad_point=gp, temporal_gradient_phase=tg)
        print(name, wv_phase.shape, float(np.abs(wv_phase).max()), float(np.abs(wv_amp).max()))

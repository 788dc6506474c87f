"""Golden vectors for the Jacobian classification row: the UNMODIFIED reference functions
compute_jacobian_matrix_for_vertex / _for_interior, classify_critical_point and
find_singularity_points_and_classify_for_all_Vk (utils/find_singularity_point.py:355-498,561-605)
run with synthetic.SurfaceMesh in place of the pyvista surface (its point_neighbors and
find_cells_intersecting_line document the adjacency semantics assumed where VTK's are not defined
by the reference).  Build container only:  python tests/golden/make_golden_classify.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NAMES = ("Node", "Focus", "Saddle", "Indeterminate")

if __name__ == "__main__":
    cof, fsp = reference_shim.load()
    for src in ("ico2_wave", "ico3_phase", "ico2_vertex_singular"):
        g = np.load(os.path.join(OUT, src + ".npz"))
        coords, tris = g["coordinates"], g["triangles"]
        surf = synthetic.SurfaceMesh(coords, tris)
        eps = float(g["eps"])
        if "V_xyz" in g:
            V = g["V_xyz"][:2]
            e = g["e"]
        else:
            V = g["V_now"][None]
            e = np.load(os.path.join(OUT, "ico2_wave.npz"))["e"]
        pts, codes, jac, counts = [], [], [], []
        with reference_shim.quiet():
            sp_all, cl_all = fsp.find_singularity_points_and_classify_for_all_Vk(V, coords, tris, eps, surf, e)
            for k, V_now in enumerate(V):
                sv, si, vmax = fsp.find_singularity_points(coords, tris, V_now, eps)
                for s in sv:
                    jac.append(fsp.compute_jacobian_matrix_for_vertex(s, V_now, surf, vmax, e))
                for s in si:
                    jac.append(fsp.compute_jacobian_matrix_for_interior(s, V_now, surf, vmax))
                counts.append(len(sv) + len(si))
                pts += list(sp_all[k])
                codes += [NAMES.index(c) for c in cl_all[k]]
        np.savez_compressed(os.path.join(OUT, "classify_" + src + ".npz"), coordinates=coords, triangles=tris, V=V, e=e, eps=eps,
                            counts=np.asarray(counts), points=np.asarray(pts).reshape(-1, 3), codes=np.asarray(codes),
                            jacobians=np.asarray(jac).reshape(-1, 2, 2))
        print(src, counts, np.bincount(np.asarray(codes), minlength=4))

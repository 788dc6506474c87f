"""Golden for the interior-zero test next to a triangle edge (find_singularity_point.py:93-137): run the UNMODIFIED
reference's find_singularity_points on fields whose zero sits at barycentric (0.3, delta, 0.7 - delta) of one face,
delta = 0, +-1e-14, +-1e-12, +-1e-9.  The reference decides with np.linalg.lstsq (SVD, :128) and the hard
thresholds lam >= 0, mu >= 0, lam + mu <= 1 (:130); this library solves the same 3x2 system in closed form.
The fixture pins which of the two faces sharing the edge the reference reports at each offset.

    python tests/golden/make_golden_edge.py        (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EPS = 1e-4
DELTAS = [0.0, 1e-14, -1e-14, 1e-12, -1e-12, 1e-9, -1e-9]
FACE = 37


def edge_field(coords, tris, normals, delta):
    """Tangent field, linear inside face FACE with its only zero at barycentric (0.3, delta, 0.7 - delta) of
    (A, B, C); everywhere else a rigid rotation about a tilted axis plus a constant (no other small vectors)."""
    N = len(coords)
    axis = np.array([0.3, -0.5, 0.8])
    V = np.cross(axis, coords) + np.array([0.4, 0.1, -0.2])
    V -= np.einsum("ij,ij->i", V, normals)[:, None] * normals
    a, b, c = tris[FACE]
    A, B, C = coords[a], coords[b], coords[c]
    P0 = 0.3 * A + delta * B + (0.7 - delta) * C
    n = np.cross(B - A, C - A)
    n /= np.linalg.norm(n)
    u = (B - A) / np.linalg.norm(B - A)
    w = np.cross(n, u)
    J = np.array([[0.9, -0.4], [0.3, 0.7]])                    # a focus-like linear field in the face plane
    for idx in (a, b, c):
        d = coords[idx] - P0
        q = J @ np.array([d @ u, d @ w])
        V[idx] = q[0] * u + q[1] * w
    return V


def main():
    _, fsp = reference_shim.load()
    coords, tris, normals, areas = synthetic.icosphere(2)
    out = {"coordinates": coords, "triangles": tris, "eps": EPS, "deltas": np.asarray(DELTAS), "face": FACE}
    fields, vidx, fidx, lm, counts, vmaxs = [], [], [], [], [], []
    for d in DELTAS:
        V = edge_field(coords, tris, normals, d)
        with reference_shim.quiet():
            sv, si, vmax = fsp.find_singularity_points(coords, tris, V, EPS)
        fields.append(V)
        counts.append((len(sv), len(si)))
        vidx += [r[0] for r in sv]
        fidx += [r[0] for r in si]
        lm += [r[3][:2] for r in si]
        vmaxs.append(vmax)
        print(f"delta {d:+.0e}: vertices {[r[0] for r in sv]} faces {[r[0] for r in si]} "
              f"lam_mu {[tuple(float(x) for x in r[3][:2]) for r in si if r[0] == FACE or True][:4]}")
    np.savez_compressed(os.path.join(OUT, "edge_zero_ico2.npz"), V_now=np.asarray(fields), counts=np.asarray(counts),
                        sing_vertex_idx=np.asarray(vidx, dtype=np.int64), sing_face_idx=np.asarray(fidx, dtype=np.int64),
                        sing_face_lam_mu=np.asarray(lm, dtype=np.float64).reshape(-1, 2), v_length_max=np.asarray(vmaxs), **out)


if __name__ == "__main__":
    main()

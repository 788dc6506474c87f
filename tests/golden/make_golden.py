"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/utils/{compute_optical_flow,find_singularity_point}.py) on seeded
synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no fixtures of its own; these files are
the pins for oracle/mof_oracle.py and for the CUDA path.  Inputs are stored next
to the outputs so the fixtures do not depend on the generators staying unchanged.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LAMBDA = 0.01   # config.yaml:3
EPS = 1e-4      # config.yaml:4

CASES = {
    # name: (mesh factory, signal kind, frames, SF, store_a2, pool)
    "ico1_wave": (lambda: synthetic.icosphere(1), "wave", 4, 512.0, True, 1),
    "ico2_wave": (lambda: synthetic.icosphere(2), "wave", 5, 512.0, True, 2),
    "patch8_wave": (lambda: synthetic.open_patch(8, seed=3), "wave", 4, 256.0, True, 1),
    "ico3_phase": (lambda: synthetic.icosphere(3, radius=2.0), "phase", 3, 512.0, False, 1),
    "pial3_wave": (lambda: synthetic.pial_like(3, radius=80.0, seed=0), "wave", 4, 2048.0, False, 1),
    "ico4_wave": (lambda: synthetic.icosphere(4), "wave", 3, 512.0, False, 2),
}


def run_case(name, cof, fsp):
    factory, kind, T, SF, store_a2, pool = CASES[name]
    coords, tris, normals, areas = factory()
    t_k = synthetic.time_axis(T, SF)
    if kind == "wave":
        I = synthetic.travelling_wave(coords, t_k, seed=0)
    else:
        I = synthetic.wrapped_phase(coords, t_k, seed=0)
    with reference_shim.quiet():
        a2, grad_w, e, integral, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        if pool > 1:   # the reference's own Pool path (compute_optical_flow.py:152-194)
            V_k, _ = cof.compute_velocity_field(pool, T, a2, grad_w, e, integral, tris, t_k, areas,
                                                LAMBDA, I, I)
        else:
            V_k = [cof.worker(k, a2, grad_w, e, integral, tris, t_k, areas, LAMBDA, I[k], I[k + 1])
                   for k in range(T - 1)]
        V_k = np.asarray(V_k)
        V_xyz = np.asarray(fsp.process_V_k(V_k, e))
        sv_idx, sf_idx, sf_lm, sf_P, vmaxs, counts = [], [], [], [], [], []
        for k in range(T - 1):
            sv, si, vmax = fsp.find_singularity_points(coords, tris, V_xyz[k], EPS)
            counts.append((len(sv), len(si)))
            sv_idx += [row[0] for row in sv]
            sf_idx += [row[0] for row in si]
            sf_lm += [row[3][:2] for row in si]
            sf_P += [row[1] for row in si]
            vmaxs.append(vmax)
        all_pts = fsp.find_singularity_points_for_all_Vk(V_xyz, coords, tris, EPS)
    out = dict(
        coordinates=coords, triangles=tris, normals=normals, areas=areas,
        t_k=np.asarray(t_k), I=I, lambda_=LAMBDA, eps=EPS,
        grad_w=grad_w, e=e, integral_wi_wj=integral,
        V_k=V_k, V_xyz=V_xyz,
        sing_counts=np.asarray(counts, dtype=np.int64).reshape(-1, 2),
        sing_vertex_idx=np.asarray(sv_idx, dtype=np.int64),
        sing_face_idx=np.asarray(sf_idx, dtype=np.int64),
        sing_face_lam_mu=np.asarray(sf_lm, dtype=np.float64).reshape(-1, 2),
        sing_face_P=np.asarray(sf_P, dtype=np.float64).reshape(-1, 3),
        v_length_max=np.asarray(vmaxs),
        all_points_flat=np.asarray([p for fr in all_pts for p in fr], dtype=np.float64).reshape(-1, 3),
    )
    if store_a2:
        a2c = sp.csr_matrix(a2)
        a2c.sort_indices()
        out.update(a2_data=a2c.data, a2_indices=a2c.indices, a2_indptr=a2c.indptr)
        # first frame's system matrix and rhs, literally as worker builds them (:106-146)
        a1 = sp.lil_matrix((2 * len(coords), 2 * len(coords)))
        a_ref, f_ref = _literal_a_f(cof, 0, a2, grad_w, e, integral, tris, t_k, areas, LAMBDA, I[0], I[1])
        out.update(a0_data=a_ref.data, a0_indices=a_ref.indices, a0_indptr=a_ref.indptr, f0=f_ref)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", len(coords), "F", len(tris), "frames", T - 1, "singular (v,f) per frame", counts)


def _literal_a_f(cof, k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k_k, I_k_kplus1):
    """The reference's worker returns only V; to pin the assembled system itself we
    run its spsolve with a capturing stub (the module-level name ``spsolve`` is the
    only thing replaced; the assembly code that runs is the reference's own)."""
    captured = {}
    orig = cof.spsolve

    def capture(a, f):
        captured["a"], captured["f"] = a.copy(), f.copy()
        return orig(a, f)

    cof.spsolve = capture
    try:
        cof.worker(k, a2, grad_w, e, integral_wi_wj, triangles, t_k, areas, lambda_, I_k_k, I_k_kplus1)
    finally:
        cof.spsolve = orig
    a = sp.csr_matrix(captured["a"])
    a.sort_indices()
    return a, captured["f"]


def special_vertex_case(cof, fsp):
    """find_singularity_points with singular *vertices* (faces touching them are
    skipped, fsp:171-172): reuse ico2 field and zero two vertices."""
    g = np.load(os.path.join(OUT, "ico2_wave.npz"))
    V = g["V_xyz"][0].copy()
    V[7] = 0.0
    V[100] *= 1e-7
    with reference_shim.quiet():
        sv, si, vmax = fsp.find_singularity_points(g["coordinates"], g["triangles"], V, EPS)
    np.savez_compressed(
        os.path.join(OUT, "ico2_vertex_singular.npz"),
        coordinates=g["coordinates"], triangles=g["triangles"], V_now=V, eps=EPS,
        sing_vertex_idx=np.asarray([r[0] for r in sv], dtype=np.int64),
        sing_face_idx=np.asarray([r[0] for r in si], dtype=np.int64),
        sing_face_lam_mu=np.asarray([r[3][:2] for r in si], dtype=np.float64).reshape(-1, 2),
        sing_face_P=np.asarray([r[1] for r in si], dtype=np.float64).reshape(-1, 3),
        v_length_max=vmax)
    print("ico2_vertex_singular", [r[0] for r in sv], [r[0] for r in si])


if __name__ == "__main__":
    cof, fsp = reference_shim.load()
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        run_case(nm, cof, fsp)
    special_vertex_case(cof, fsp)

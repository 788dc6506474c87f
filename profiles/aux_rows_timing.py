"""Timing of the auxiliary 'next' rows at full size on one GPU (ico7 pial-like mesh): RBF interpolation (K8)
and multi-ring winding numbers (K7).  python profiles/aux_rows_timing.py > gpurun_out/aux_rows.json"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from manifold_based_optical_flow_method_b200 import S2_interpolate as s2, S7_winding_line as s7, synthetic  # noqa: E402
from manifold_based_optical_flow_method_b200 import find_singularity_point as fsp  # noqa: E402

dev = torch.device("cuda:0")
coords, tris, normals, areas = synthetic.pial_like(7)
N = len(coords)
T, m = 1000, 128
rng = np.random.default_rng(0)
cap = np.nonzero(coords[:, 2] > 0.3 * np.abs(coords).max())[0]
sel = cap[rng.choice(len(cap), m, replace=False)]
c = coords[sel] + rng.normal(0, 0.3, (m, 3))
t_k = synthetic.time_axis(T, 512.0)
d = np.stack([np.sin(0.04 * coords[sel] @ np.array([1.0, 0.5, 0.2]) - 40.0 * t) for t in t_k])
v_dev = torch.from_numpy(coords).to(dev)
out = torch.empty((T, N), dtype=torch.float64, device=dev)
res = {}
for phase in (False, True):
    data = np.exp(1j * d) if phase else d
    best = 1e30
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s2.rbf_interpolate_device(data, c, v_dev, phase=phase, out=out)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    res["rbf_phase" if phase else "rbf"] = {"ms": best, "frames_per_s": T / best * 1e3,
                                            "tflops_fp64": (4.0 if phase else 2.0) * T * N * m / best / 1e9}
# winding numbers: rotation field about a tilted axis + its two poles and 62 more points, 64 frames
from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof  # noqa: E402
_, _, e, _, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
axis = np.array([0.3, 0.2, 1.0]); axis /= np.linalg.norm(axis)
V = np.cross(axis, coords)
Vk = np.stack([V] * 8)
off_axis = np.linalg.norm(coords - np.outer(coords @ axis, axis), axis=1)          # the field vanishes where the axis pierces the surface
poles = coords[[np.argmin(np.where(coords @ axis > 0, off_axis, np.inf)), np.argmin(np.where(coords @ axis < 0, off_axis, np.inf))]]
pts = np.concatenate([poles, coords[rng.choice(N, 62, replace=False)]])
pts_all = np.concatenate([pts] * 8)
fop = np.repeat(np.arange(8), len(pts))
best = 1e30
for rep in range(3):
    t0 = time.time()
    r = s7.winding_numbers(tris, coords, pts_all, fop, Vk, e)
    torch.cuda.synchronize()
    if rep:
        best = min(best, time.time() - t0)
res["winding"] = {"points": len(pts_all), "wall_ms_including_uploads": best * 1e3,
                  "counts_at_poles": r.counts[:2].tolist(), "types_at_poles": r.types[:2].tolist()}
print(json.dumps(res))

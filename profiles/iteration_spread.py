"""Per-frame iteration counts of one C2 step and what the 32-frame grouping costs (frozen lanes of groups that still iterate move bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, json
from manifold_based_optical_flow_method_b200 import compute_optical_flow as cof, synthetic
coords, tris, normals, areas = synthetic.pial_like(7)
T=1000; t_k = synthetic.time_axis(T, 512.0); I = synthetic.travelling_wave(coords, t_k, seed=0)
a2, gw, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
I_dev = torch.from_numpy(I).cuda()
V, info = cof.solve_on_device(a2, I_dev, I_dev, t_k, 0.01, 0, T-1)
it = info.iterations
pad = np.concatenate([it, np.zeros(1024-len(it), int)])
g = pad.reshape(32,32)
gmax = g.max(axis=1)
print('mean', it.mean(), 'min', it.min(), 'max', it.max())
print('group max', gmax.tolist())
print('lane-iterations moved', int((gmax*32).sum()), 'useful', int(it.sum()), 'ratio', it.sum()/(gmax*32).sum())
print('per-iteration active groups from 130:', [(k, int((gmax>k).sum())) for k in range(128, 151, 2)])
np.save('gpurun_out/iters_c2.npy', it)

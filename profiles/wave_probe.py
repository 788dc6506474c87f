"""Config 5 on one GPU, kernel by kernel: S5 wave speed of a T-frame wrapped-phase trial on the pial-like ico7 mesh.

    python profiles/wave_probe.py [--level 7] [--frames 1000] [--steps 10] [--orders 0,1]

For every mesh ordering asked for (0 = reference vertex order, what S5_compute_wave_v.py uses; 1 = Cuthill-McKee, the
round-2 first version's) and every variant of the row kernel (include/mof_b200.h: mof_wave_set_variant) it times, with CUDA events on the launching stream after
warm-up: the whole mof_wave_speed call (coefficient rows + transpose in + row kernel), the row kernel alone
(mof_wave_stencil), and reports them against the algorithmic 16 N bytes per frame.  It also records whether a variant
produces the same array bit for bit as the first one (the kernel variant must not matter; another mesh ordering sums a
row's products in another column order and may differ in the last bits) and compares a few frames with the numpy
oracle.  One JSON line on stdout.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from manifold_based_optical_flow_method_b200 import _lib, synthetic  # noqa: E402
from manifold_based_optical_flow_method_b200 import S5_compute_wave_v as s5  # noqa: E402
from manifold_based_optical_flow_method_b200.mesh import MeshOperator  # noqa: E402


def edge_checks(lib, variants):
    """The parity cases of tests/test_wave_speed.py under every kernel variant, in this process: the golden of the
    unmodified S5 on the open patch (irregular valence up to 8: rows longer than the staged slots), and ragged shards
    of a 70-frame trial (three 32-frame groups, the last one partial) against the whole trial, bit for bit."""
    from manifold_based_optical_flow_method_b200.distributed import shard_range
    from oracle import mof_oracle
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "s5_patch7.npz")))
    dt = float(g["dt"])
    surf = synthetic.SurfaceMesh(g["coordinates"], g["triangles"], g["normals"], g["areas"])
    coords, tris, normals, areas = synthetic.pial_like(3)
    e3 = mof_oracle.orthonormal_basis(normals)
    T = 70
    t_k = synthetic.time_axis(T, 512.0)
    data = {True: synthetic.wrapped_phase(coords, t_k, seed=2, omega=300.0), False: synthetic.travelling_wave(coords, t_k, seed=2)}

    def rel(a, b):
        m = np.isfinite(b)
        return float(np.linalg.norm(a[m] - b[m]) / np.linalg.norm(b[m]))
    res = {}
    for variant in variants:
        _lib.check(lib.mof_wave_set_variant(variant))
        r = {"golden_phase": rel(s5.wave_velocity_phase(surf, g["phases"], dt, len(g["phases"]), g["e"]), g["wave_velocity_phase"]),
             "golden_amplitude": rel(s5.wave_velocity_amplitude(surf, g["potentials"], dt, len(g["potentials"]), g["e"]),
                                     g["wave_velocity_amplitude"])}
        op = s5._operator(coords, tris, areas, e3)
        ok = True
        for phase in (True, False):
            d = torch.from_numpy(np.ascontiguousarray(data[phase])).to(op.device)
            _, whole = s5.wave_speed_device(op, d, 0, T, 0, T, 1 / 512.0, phase)
            r["oracle_" + ("phase" if phase else "amplitude")] = rel(
                whole.cpu().numpy(), mof_oracle.wave_velocity(coords, tris, areas, data[phase], 1 / 512.0, e3, phase=phase))
            for world in (2, 3, 8):
                for rank in range(world):
                    k0, k1 = shard_range(T, world, rank)
                    a, b = s5.halo_rows(k0, k1, T, phase)
                    _, part = s5.wave_speed_device(op, d[a:b], a, T, k0 - a, k1 - k0, 1 / 512.0, phase)
                    ok = ok and bool(torch.equal(part, whole[k0:k1]))
        r["shards_bit_identical"] = ok
        r["pass"] = bool(ok and max(v for k, v in r.items() if k.startswith(("golden", "oracle"))) <= 1e-12)
        res[str(variant)] = r
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=7)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--orders", default="0,1")
    ap.add_argument("--oracle-frames", type=int, default=4)
    ap.add_argument("--edge-checks", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    coords, tris, normals, areas = synthetic.pial_like(args.level)
    N, T = len(coords), args.frames
    t_k = synthetic.time_axis(T, 512.0)
    phases = synthetic.wrapped_phase(coords, t_k, seed=1, omega=500.0)
    from oracle import mof_oracle
    e = mof_oracle.orthonormal_basis(normals)
    d = torch.from_numpy(np.ascontiguousarray(phases)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        pass
    out = {"n_vertices": N, "frames": T, "steps": args.steps, "algorithmic_bytes": 16.0 * N * T, "peak_gbs": peak, "variants": []}
    first = None
    default_variant = int(lib.mof_wave_get_variant())
    out["default_variant"] = default_variant
    for order in [int(x) for x in args.orders.split(",")]:
        nrm = np.zeros_like(coords)
        nrm[:, 2] = 1.0
        op = MeshOperator(coords, nrm, tris, areas, reorder=order)
        op.use_geometry(None, e, None, areas)
        ms = op.struct()
        work = torch.empty((int(lib.mof_wave_work_doubles(ctypes.byref(ms), T, 0, 1)),), dtype=torch.float64, device=dev)
        wv = torch.empty((T, N), dtype=torch.float64, device=dev)
        for gp in range(7):
            _lib.check(lib.mof_wave_set_variant(gp))
            wv.fill_(float("nan"))
            for _ in range(3):
                s5.wave_speed_device(op, d, 0, T, 0, T, 1 / 512.0, True, work=work, wave_out=wv)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                s5.wave_speed_device(op, d, 0, T, 0, T, 1 / 512.0, True, work=work, wave_out=wv)
            e1.record()
            for _ in range(args.steps):
                _lib.check(lib.mof_wave_stencil(ctypes.byref(ms), T, 0, T, 0, T, 1 / 512.0, 1, None, wv.data_ptr(), work.data_ptr(), st))
            e2.record()
            torch.cuda.synchronize()
            call_ms, rows_ms = e0.elapsed_time(e1) / args.steps, e1.elapsed_time(e2) / args.steps
            rel_to_first = 0.0
            if first is None:
                first = wv.clone()
                same = True
            else:
                fin = torch.isfinite(first) & torch.isfinite(wv)
                rel_to_first = float(torch.linalg.vector_norm(wv[fin] - first[fin]) / torch.linalg.vector_norm(first[fin]))
                same = bool(torch.equal(torch.nan_to_num(wv, nan=0.0, posinf=1e300, neginf=-1e300),
                                        torch.nan_to_num(first, nan=0.0, posinf=1e300, neginf=-1e300)))
            rec = {"order": order, "variant": gp, "call_ms": call_ms, "rows_ms": rows_ms, "coef_plus_pack_ms": call_ms - rows_ms,
                   "call_gbs": 16.0 * N * T / call_ms / 1e6, "rows_gbs": 16.0 * N * T / rows_ms / 1e6, "bit_identical_to_first": same, "rel_l2_to_first": rel_to_first}
            if peak:
                rec["call_frac"], rec["rows_frac"] = rec["call_gbs"] / peak, rec["rows_gbs"] / peak
            out["variants"].append(rec)
        del op, work, wv
    k = args.oracle_frames
    if k:
        wo = mof_oracle.wave_velocity(coords, tris, areas, phases[:k + 1], 1 / 512.0, e, phase=True)[:k]
        got = first[:k].cpu().numpy()
        m = np.isfinite(wo)
        out["rel_l2_vs_oracle"] = float(np.linalg.norm(got[m] - wo[m]) / np.linalg.norm(wo[m]))
    if args.edge_checks:
        out["edge_checks"] = edge_checks(lib, range(7))
    _lib.check(lib.mof_wave_set_variant(default_variant))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Context number for BASELINE.md section 3 (not a target): the UNMODIFIED reference on config 1, timed in the
build container (the reference tree does not travel to the GPU box).

    python profiles/time_literal_reference_c1.py [--frames 64] [--procs N]

Config 1: icosphere level 5 (10,242 vertices), 64-frame travelling wave -> 63 solves through the reference's own
compute_velocity_field (multiprocessing.Pool, utils/compute_optical_flow.py:152-194), plus
compute_geometrical_quantities (:27-97) and find_singularity_points (utils/find_singularity_point.py:140-189)
on a few frames.  Writes profiles/r2_literal_reference_c1.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifold_based_optical_flow_method_b200 import synthetic  # noqa: E402
from oracle import reference_shim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--detect-frames", type=int, default=2)
    args = ap.parse_args()
    cof, fsp = reference_shim.load()
    coords, tris, normals, areas = synthetic.icosphere(5)
    T = args.frames
    t_k = list(synthetic.time_axis(T, 512.0))
    I = synthetic.travelling_wave(coords, np.asarray(t_k), seed=0)
    out = {"config": "C1: icosphere level 5 (10,242 vertices / 20,480 faces), %d-frame travelling wave" % T,
           "host": {"cpu_count": os.cpu_count(), "procs": args.procs}, "kind": "reference (unmodified, pure Python)"}
    with reference_shim.quiet():
        t0 = time.time()
        a2, grad_w, e, integ, _ = cof.compute_geometrical_quantities(coords, normals, tris, areas)
        out["geometry_seconds"] = time.time() - t0
        t0 = time.time()
        V_k, _ = cof.compute_velocity_field(args.procs, T, a2, grad_w, e, integ, tris, t_k, areas, 0.01, I, I)
        out["velocity_seconds"] = time.time() - t0
        out["solves"] = len(V_k)
        out["frames_per_s"] = len(V_k) / out["velocity_seconds"]
        t0 = time.time()
        Vx = fsp.process_V_k(V_k[:args.detect_frames], e)
        out["process_V_k_seconds_per_frame"] = (time.time() - t0) / args.detect_frames
        t0 = time.time()
        for k in range(args.detect_frames):
            fsp.find_singularity_points(coords, tris, Vx[k], 1e-4)
        out["detection_seconds_per_frame"] = (time.time() - t0) / args.detect_frames
    path = os.path.join(ROOT, "profiles", "r2_literal_reference_c1.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

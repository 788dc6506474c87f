"""TEST INFRASTRUCTURE: builds tests/hostcheck/libmof_hostcheck.so with g++ (no CUDA) from
hostcheck.cpp + the product's host-side sources (pattern.cpp, error.cpp) and the shared
kernel bodies (csrc/mof_bodies.h).  Used by tests/test_host_logic.py only."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "manifold_based_optical_flow_method_b200", "csrc")
LIB = os.path.join(HERE, "libmof_hostcheck.so")


def load():
    srcs = [os.path.join(HERE, "hostcheck.cpp"), os.path.join(CSRC, "pattern.cpp"), os.path.join(CSRC, "error.cpp")]
    deps = srcs + [os.path.join(CSRC, "mof_bodies.h"), os.path.join(ROOT, "include", "mof_b200.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        # -ffp-contract=off: numpy does not fuse multiply-add either
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                        "-I", CSRC, "-o", LIB] + srcs, check=True)
    lib = ctypes.CDLL(LIB)
    P, I64, I32, D = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    lib.hc_geom.argtypes = [P] * 6
    lib.hc_a2.argtypes = [P, P]
    lib.hc_pack.argtypes = [P, I32, I32, P, P, I64, P, P, P]
    lib.hc_assemble.argtypes = [P, I32, P, P, D, D, P, P, P]
    lib.hc_sweep_back.argtypes = [P, I32, P, P, P, P, I32, I32, P, P, D, I32]
    lib.hc_sweep_fwd.argtypes = [P, I32, P, P, P, P, I32, I32, D, I32, P]
    lib.hc_sweep_rows_back.argtypes = [P, I32, P, P, P, P, I64, I64, P, P, D, I32]
    lib.hc_sweep_rows_fwd.argtypes = [P, I32, P, P, P, P, I64, I64, D, I32, P]
    lib.hc_sweep_rows_back.restype = lib.hc_sweep_rows_fwd.restype = None
    lib.hc_spmv.argtypes = [P, I32, P, P, P]
    lib.hc_tangent.argtypes = [I64, I64, P, I64, P, P, P, P]
    lib.hc_detect.argtypes = [I64, I64, P, P, P, D, D, P, P, P, P, P, P]
    lib.hc_winding.argtypes = [I64, P, P, P, P, P, P, I32, P, P, P, P]
    lib.hc_winding.restype = None
    lib.hc_wave_coef.argtypes = [P, P, P]
    lib.hc_wave_rows.argtypes = [P, I32, I64, I64, I64, I64, I64, P, I64, D, I32, P, P]
    lib.hc_wave_coef.restype = lib.hc_wave_rows.restype = None
    for f in (lib.hc_geom, lib.hc_a2, lib.hc_pack, lib.hc_assemble, lib.hc_spmv, lib.hc_tangent, lib.hc_detect,
              lib.hc_sweep_back, lib.hc_sweep_fwd):
        f.restype = None
    return lib

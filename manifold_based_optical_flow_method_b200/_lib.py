"""ctypes binding of csrc/libmof_b200.so (C ABI declared in include/mof_b200.h).

There is no CPU fallback: if the library is missing the loader tries to build it with
nvcc (in-tree), and if that is impossible it raises.  Every device entry point goes
through ``check()`` which turns a non-zero return code into ``MofError``.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmof_b200.so")

GROUP = 32           # MOF_GROUP
TILE_ROWS = 64       # MOF_TILE_ROWS
DETECT_CHUNK = 1024  # MOF_DETECT_CHUNK
MAX_COLORS = 16      # MOF_MAX_COLORS
SCAL_SLOTS = 20      # MOF_SCAL_SLOTS

PATH_JACOBI, PATH_MULTICOLOUR, PATH_LEVEL_LAUNCHES, PATH_LEVEL_PERSISTENT = 0, 1, 2, 3
STATUS_CONVERGED, STATUS_MAXITER, STATUS_BREAKDOWN, STATUS_ZERO_RHS = 0, 1, 2, 3

# every symbol include/mof_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "mof_last_error_string", "mof_version",
    "mof_pattern_create", "mof_pattern_destroy", "mof_pattern_num_blocks", "mof_pattern_num_contrib",
    "mof_pattern_max_row_blocks", "mof_pattern_bandwidth", "mof_pattern_export", "mof_pattern_colors", "mof_pattern_levels",
    "mof_num_tiles", "mof_state_ints",
    "mof_geom_basis", "mof_geom_gradw", "mof_geom_a2",
    "mof_pack_frames", "mof_assemble_batch",
    "mof_level_desc_build", "mof_spmv_batch", "mof_pcg_solve_batch", "mof_pcg_last_path", "mof_unpack_solution",
    "mof_tangent_to_xyz", "mof_vmax", "mof_singularity_flags", "mof_singularity_compact",
    "mof_classify_singularities", "mof_winding_numbers", "mof_wave_speed", "mof_wave_work_doubles", "mof_wave_stencil", "mof_wave_set_variant", "mof_wave_get_variant", "mof_rbf_fit", "mof_rbf_evaluate", "mof_csv_write", "mof_csv_dims", "mof_csv_read",
]


class MofError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libmof_b200 error {code}: {text}")
        self.code = code


class MeshDev(Structure):
    """mof_mesh_dev"""
    _fields_ = [
        ("n_vertices", c_int64), ("n_faces", c_int64), ("n_blocks", c_int64), ("n_contrib", c_int64),
        ("perm", c_void_p), ("rowptr", c_void_p), ("col", c_void_p), ("diag", c_void_p),
        ("cptr", c_void_p), ("centry", c_void_p), ("tri", c_void_p),
        ("e", c_void_p), ("grad_w", c_void_p), ("integral", c_void_p), ("areas", c_void_p), ("a2v", c_void_p),
        ("n_colors", c_int32), ("color_tile_ptr", c_int32 * (MAX_COLORS + 1)),
        ("n_levels", c_int32), ("level_stage_blocks", c_int32), ("level_ptr", c_void_p),     # level_ptr: HOST int32[n_levels+1]
        ("level_desc", c_void_p),                                                   # device int32 [2][N][8] or NULL
    ]


class BatchDev(Structure):
    """mof_batch_dev"""
    _fields_ = [
        ("n_groups", c_int32), ("n_frames", c_int32),
        ("It", c_void_p), ("dIt", c_void_p), ("vals", c_void_p), ("rhs", c_void_p), ("minv", c_void_p),
        ("x", c_void_p), ("r", c_void_p), ("z", c_void_p), ("p", c_void_p), ("ap", c_void_p), ("t", c_void_p),
        ("partial", c_void_p), ("scal", c_void_p), ("state", c_void_p), ("ready", c_void_p),
    ]


class PcgProfile(Structure):
    """mof_pcg_profile"""
    _fields_ = [
        ("ms_spmv", c_double), ("ms_update", c_double), ("ms_pupdate", c_double),
        ("samples", c_int64), ("group_launches", c_int64), ("frame_launches", c_int64),
        ("iterations_total", c_int64), ("launches_total", c_int64),
        ("ms_iter", c_double), ("iter_launches", c_int64), ("phase_ns", c_double * 4),
    ]


def add_profile(dst, src):
    """dst += src, field by field (PcgProfile)."""
    for name, ctype in PcgProfile._fields_:
        a, b = getattr(dst, name), getattr(src, name)
        if hasattr(a, "__len__"):
            for i in range(len(a)):
                a[i] += b[i]
        else:
            setattr(dst, name, a + b)


_lib = None


def _declare(lib):
    P = c_void_p
    lib.mof_last_error_string.restype = c_char_p
    lib.mof_last_error_string.argtypes = []
    lib.mof_version.restype = c_int
    lib.mof_pattern_create.restype = c_int
    lib.mof_pattern_create.argtypes = [c_int64, c_int64, P, c_int, P, POINTER(c_void_p)]
    lib.mof_pattern_levels.restype = c_int32
    lib.mof_pattern_levels.argtypes = [P, P]
    lib.mof_pattern_colors.restype = c_int
    lib.mof_pattern_colors.argtypes = [P, POINTER(c_int32), P]
    lib.mof_pattern_destroy.restype = None
    lib.mof_pattern_destroy.argtypes = [P]
    for name in ("mof_pattern_num_blocks", "mof_pattern_num_contrib", "mof_pattern_max_row_blocks", "mof_pattern_bandwidth"):
        getattr(lib, name).restype = c_int64
        getattr(lib, name).argtypes = [P]
    lib.mof_pattern_export.restype = c_int
    lib.mof_pattern_export.argtypes = [P] * 8
    lib.mof_num_tiles.restype = c_int64
    lib.mof_num_tiles.argtypes = [c_int64]
    lib.mof_state_ints.restype = c_int64
    lib.mof_state_ints.argtypes = [c_int32]
    lib.mof_geom_basis.restype = c_int
    lib.mof_geom_basis.argtypes = [c_int64, P, P, P]
    lib.mof_geom_gradw.restype = c_int
    lib.mof_geom_gradw.argtypes = [c_int64, P, P, P, P, P, P]
    lib.mof_geom_a2.restype = c_int
    lib.mof_geom_a2.argtypes = [POINTER(MeshDev), P, P]
    lib.mof_pack_frames.restype = c_int
    lib.mof_pack_frames.argtypes = [POINTER(MeshDev), POINTER(BatchDev), P, P, c_int64, P, P]
    lib.mof_assemble_batch.restype = c_int
    lib.mof_assemble_batch.argtypes = [POINTER(MeshDev), POINTER(BatchDev), c_double, c_double, P]
    lib.mof_level_desc_build.restype = c_int
    lib.mof_level_desc_build.argtypes = [POINTER(MeshDev), P, P]
    lib.mof_spmv_batch.restype = c_int
    lib.mof_spmv_batch.argtypes = [POINTER(MeshDev), POINTER(BatchDev), P, P, P]
    lib.mof_pcg_solve_batch.restype = c_int
    lib.mof_pcg_solve_batch.argtypes = [POINTER(MeshDev), POINTER(BatchDev), c_double, c_double, c_int32, c_int32, c_int32, P, P, P,
                                        POINTER(PcgProfile), P]
    lib.mof_pcg_last_path.restype = c_int
    lib.mof_pcg_last_path.argtypes = [P]
    lib.mof_unpack_solution.restype = c_int
    lib.mof_unpack_solution.argtypes = [POINTER(MeshDev), POINTER(BatchDev), P, c_int64, P]
    lib.mof_tangent_to_xyz.restype = c_int
    lib.mof_tangent_to_xyz.argtypes = [c_int64, c_int64, P, c_int64, P, P, P, P, P]
    lib.mof_vmax.restype = c_int
    lib.mof_vmax.argtypes = [c_int64, c_int64, P, P, P]
    lib.mof_singularity_flags.restype = c_int
    lib.mof_singularity_flags.argtypes = [c_int64, c_int64, c_int64, P, P, P, P, c_double, P, P, P, P, P, P]
    lib.mof_classify_singularities.restype = c_int
    lib.mof_classify_singularities.argtypes = [c_int64, c_int64, c_int64] + [P] * 10 + [c_int64, c_int64] + [P] * 8
    lib.mof_winding_numbers.restype = c_int
    lib.mof_winding_numbers.argtypes = [c_int64, c_int64] + [P] * 5 + [c_int64, P, P, c_int] + [P] * 6
    lib.mof_rbf_fit.restype = c_int
    lib.mof_rbf_fit.argtypes = [c_int64, P, c_double, c_int64, P, c_int64, P, P, P, P, P]
    lib.mof_rbf_evaluate.restype = c_int
    lib.mof_rbf_evaluate.argtypes = [c_int64, c_int64, c_int64, P, P, c_double, P, c_int, P, c_int64, P]
    lib.mof_csv_write.restype = c_int
    lib.mof_csv_write.argtypes = [c_char_p, P, c_int64, c_int64, c_int]
    lib.mof_csv_dims.restype = c_int
    lib.mof_csv_dims.argtypes = [c_char_p, POINTER(c_int64), POINTER(c_int64)]
    lib.mof_csv_read.restype = c_int
    lib.mof_csv_read.argtypes = [c_char_p, P, c_int64, c_int64, c_int]
    lib.mof_wave_speed.restype = c_int
    lib.mof_wave_speed.argtypes = [POINTER(MeshDev), c_int64, c_int64, c_int64, c_int64, c_int64, P, c_int64, c_double, c_int, P, P, P, P]
    lib.mof_wave_stencil.restype = c_int
    lib.mof_wave_stencil.argtypes = [POINTER(MeshDev), c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_int, P, P, P, P]
    lib.mof_wave_work_doubles.restype = c_int64
    lib.mof_wave_set_variant.restype = c_int
    lib.mof_wave_set_variant.argtypes = [c_int]
    lib.mof_wave_get_variant.restype = c_int
    lib.mof_wave_get_variant.argtypes = []
    lib.mof_wave_work_doubles.argtypes = [POINTER(MeshDev), c_int64, c_int, c_int]
    lib.mof_singularity_compact.restype = c_int
    lib.mof_singularity_compact.argtypes = [c_int64, c_int64, c_int64] + [P] * 16


def load():
    """Load libmof_b200.so, (re)building it first when it is missing or older than its sources.
    Raises if it cannot be had.  The build is serialised across processes (torchrun starts every
    rank at once) with a file lock; nvcc writes to a temporary name that is renamed into place."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    try:
        stale = _build.needs_build()
    except OSError:
        stale = not os.path.exists(LIB_PATH)
    if stale:
        try:
            _build.build_locked()
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise
            import warnings
            warnings.warn(f"{LIB_PATH} is older than its sources and could not be rebuilt (nvcc missing?); "
                          "loading the existing binary")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise RuntimeError(
            f"cannot load {LIB_PATH}: {exc}. This package has no CPU fallback; build the CUDA "
            "library with `python -m manifold_based_optical_flow_method_b200.build`.") from exc
    _declare(lib)
    _lib = lib
    return lib


def last_error():
    return load().mof_last_error_string().decode("utf-8", "replace")


def check(rc, allow_positive=False):
    """rc < 0: CUDA/runtime error -> raise.  rc > 0: numerical status -> raise unless allowed."""
    if rc < 0 or (rc > 0 and not allow_positive):
        raise MofError(rc, last_error())
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError(
            "manifold_based_optical_flow_method_b200 needs a CUDA device (B200, sm_100a); "
            "there is no CPU fallback.")
    return torch

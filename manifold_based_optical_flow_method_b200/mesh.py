"""Mesh-level objects: the block sparsity pattern (host) and the device-resident
geometry handle that stands in for the reference's ``a2`` matrix.

Reference: compute_geometrical_quantities, utils/compute_optical_flow.py:27-97.
"""
import ctypes
import time

import numpy as np

from . import _lib


def _as_f64(a, name, shape_tail):
    a = np.asarray(a)
    if a.dtype != np.float64:
        # pyvista hands out float32 points/normals; the reference then evaluates
        # compute_gradient_w in float32.  Parity is defined on float64
        # inputs: promote once here.
        a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    if a.ndim != 1 + len(shape_tail) or tuple(a.shape[1:]) != tuple(shape_tail):
        raise ValueError(f"{name} must have shape (n, {', '.join(map(str, shape_tail))}), got {a.shape}")
    return a


class Pattern:
    """Host-side block pattern of the 2N x 2N system (works without a GPU).

    Attributes (numpy int32): perm, iperm, rowptr, col, diag, cptr, centry, tri; ints
    n_vertices, n_faces, n_blocks, n_contrib, max_row_blocks, bandwidth."""

    def __init__(self, n_vertices, triangles, reorder=True, coordinates=None):
        """reorder: False/0 identity, True/1 Cuthill-McKee, 2 block multicolour (needs coordinates),
        3 level-scheduled Cuthill-McKee (rows grouped by dependency level)."""
        lib = _lib.load()
        tri64 = np.ascontiguousarray(np.asarray(triangles), dtype=np.int64)
        if tri64.ndim != 2 or tri64.shape[1] != 3:
            raise ValueError(f"triangles must have shape (F, 3), got {tri64.shape}")
        N, F = int(n_vertices), int(tri64.shape[0])
        handle = ctypes.c_void_p()
        mode = int(reorder)
        xyz = None
        if coordinates is not None:
            xyz = np.ascontiguousarray(np.asarray(coordinates, dtype=np.float64))
            if xyz.shape != (N, 3):
                raise ValueError(f"coordinates must have shape ({N}, 3), got {xyz.shape}")
        _lib.check(lib.mof_pattern_create(N, F, tri64.ctypes.data, mode, xyz.ctypes.data if xyz is not None else None,
                                          ctypes.byref(handle)))
        self.reorder = mode
        try:
            nc = ctypes.c_int32(0)
            ctp = (ctypes.c_int32 * (_lib.MAX_COLORS + 1))()
            _lib.check(lib.mof_pattern_colors(handle, ctypes.byref(nc), ctp))
            self.n_colors = int(nc.value)
            self.color_tile_ptr = np.array(list(ctp)[:self.n_colors + 1], dtype=np.int32) if self.n_colors else np.zeros(1, np.int32)
            self.n_levels = int(lib.mof_pattern_levels(handle, None))
            self.level_ptr = np.zeros(self.n_levels + 1, np.int32)
            if self.n_levels:
                lib.mof_pattern_levels(handle, self.level_ptr.ctypes.data)
            nb = int(lib.mof_pattern_num_blocks(handle))
            nc = int(lib.mof_pattern_num_contrib(handle))
            self.n_vertices, self.n_faces, self.n_blocks, self.n_contrib = N, F, nb, nc
            self.max_row_blocks = int(lib.mof_pattern_max_row_blocks(handle))
            self.bandwidth = int(lib.mof_pattern_bandwidth(handle))
            self.perm = np.empty(N, np.int32)
            self.rowptr = np.empty(N + 1, np.int32)
            self.col = np.empty(nb, np.int32)
            self.diag = np.empty(N, np.int32)
            self.cptr = np.empty(nb + 1, np.int32)
            self.centry = np.empty(nc, np.int32)
            self.tri = np.empty((F, 3), np.int32)
            _lib.check(lib.mof_pattern_export(handle, self.perm.ctypes.data, self.rowptr.ctypes.data,
                                              self.col.ctypes.data, self.diag.ctypes.data, self.cptr.ctypes.data,
                                              self.centry.ctypes.data, self.tri.ctypes.data))
        finally:
            lib.mof_pattern_destroy(handle)
        self.iperm = np.empty(N, np.int32)
        self.iperm[self.perm] = np.arange(N, dtype=np.int32)

    def block_rows(self):
        """internal row vertex of every block, (nb,) int64"""
        return np.repeat(np.arange(self.n_vertices, dtype=np.int64), np.diff(self.rowptr))

    def block_values_to_csr(self, block_vals):
        """(nb,4) block values [2*alpha+beta] in internal numbering -> scipy CSR (2N,2N) in
        the reference's indexing  vertex + N*alpha  (compute_optical_flow.py:83-84)."""
        import scipy.sparse as sp
        N = self.n_vertices
        bv = np.asarray(block_vals, dtype=np.float64).reshape(self.n_blocks, 4)
        r = self.perm[self.block_rows()].astype(np.int64)
        c = self.perm[self.col].astype(np.int64)
        rows = np.stack([r, r, r + N, r + N], axis=1)
        cols = np.stack([c, c + N, c, c + N], axis=1)
        m = sp.coo_matrix((bv.ravel(), (rows.ravel(), cols.ravel())), shape=(2 * N, 2 * N)).tocsr()
        m.sort_indices()
        return m

    def csr_to_block_values(self, mat):
        """inverse of block_values_to_csr for a matrix given in the reference's indexing."""
        import scipy.sparse as sp
        N = self.n_vertices
        m = sp.csr_matrix(mat)
        r = self.perm[self.block_rows()].astype(np.int64)
        c = self.perm[self.col].astype(np.int64)
        out = np.empty((self.n_blocks, 4))
        for al in range(2):
            for be in range(2):
                out[:, 2 * al + be] = np.asarray(m[r + N * al, c + N * be]).ravel()
        return out


class MeshOperator:
    """Device-resident mesh geometry + a2 block values.

    This is what ``compute_geometrical_quantities`` returns in the position of the
    reference's ``a2`` (a scipy lil_matrix, compute_optical_flow.py:49): the only
    consumers of ``a2`` are ``worker`` / ``compute_velocity_field``, which accept this
    handle.  ``tocsr()`` gives the same matrix as the reference's for comparison."""

    def __init__(self, coordinates, normals, triangles, areas, device=None, reorder=3):
        torch = _lib.require_cuda()
        lib = _lib.load()
        t0 = time.time()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        coords = _as_f64(coordinates, "coordinates", (3,))
        nrm = _as_f64(normals, "normals", (3,))
        ar = np.ascontiguousarray(np.asarray(areas, dtype=np.float64).reshape(-1))
        N = coords.shape[0]
        if nrm.shape[0] != N:
            raise ValueError("normals and coordinates differ in length")
        self.pattern = P = Pattern(N, triangles, reorder=reorder, coordinates=coords)
        if ar.shape[0] != P.n_faces:
            raise ValueError("areas and triangles differ in length")
        self.n_vertices, self.n_faces, self.n_blocks = N, P.n_faces, P.n_blocks
        self.shape = (2 * N, 2 * N)
        self.triangles = np.ascontiguousarray(np.asarray(triangles))
        dev = self.device

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        self.d_level_desc = None
        # stage configuration of the persistent level kernel: four matrix blocks per row staged, two stages per warp.
        # The alternative (three blocks, three stages: one more item in flight per warp) measured 0.4 % slower at ico7
        # -- the sweeps sit at the memory system's rate for their access mix, not at a lack of items in flight -- and
        # stays selectable for the tests that run both (MOF_LEVEL_STAGE_BLOCKS=3).
        self.level_stage_blocks = 4
        import os
        if os.environ.get("MOF_LEVEL_STAGE_BLOCKS") in ("3", "4"):      # tests exercise both configurations on one mesh
            self.level_stage_blocks = int(os.environ["MOF_LEVEL_STAGE_BLOCKS"])
        with torch.cuda.device(dev):
            self.d_perm, self.d_rowptr, self.d_col, self.d_diag = up(P.perm), up(P.rowptr), up(P.col), up(P.diag)
            self.d_cptr, self.d_centry, self.d_tri = up(P.cptr), up(P.centry), up(P.tri)
            d_coords = up(coords[P.perm])
            d_normals = up(nrm[P.perm])
            self.d_areas = up(ar)
            self.d_e = torch.empty((N, 2, 3), dtype=torch.float64, device=dev)
            self.d_grad_w = torch.empty((P.n_faces, 3, 3), dtype=torch.float64, device=dev)
            self.d_integral = torch.empty((P.n_faces, 2), dtype=torch.float64, device=dev)
            self.d_a2v = torch.empty((P.n_blocks, 4), dtype=torch.float64, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mof_geom_basis(N, d_normals.data_ptr(), self.d_e.data_ptr(), st))
            _lib.check(lib.mof_geom_gradw(P.n_faces, d_coords.data_ptr(), self.d_tri.data_ptr(), self.d_areas.data_ptr(),
                                          self.d_grad_w.data_ptr(), self.d_integral.data_ptr(), st))
            _lib.check(lib.mof_geom_a2(ctypes.byref(self.struct()), self.d_a2v.data_ptr(), st))
            if P.n_levels:           # row descriptors of the persistent level-scheduled SSOR kernel
                desc = torch.empty((2, N, 8), dtype=torch.int32, device=dev)
                if _lib.check(lib.mof_level_desc_build(ctypes.byref(self.struct()), desc.data_ptr(), st), allow_positive=True) == 0:
                    self.d_level_desc = desc
            # host copies in the reference's vertex order (what the reference returns, :97)
            e_int = self.d_e.cpu().numpy()
            self.e = np.empty_like(e_int)
            self.e[P.perm] = e_int
            self.grad_w = self.d_grad_w.cpu().numpy()
            self.integral_wi_wj = self.d_integral.cpu().numpy()
            self.areas = ar
        self._geom_ids = (id(self.grad_w), id(self.e), id(self.integral_wi_wj), id(self.areas))
        self.build_seconds = time.time() - t0

    # -- C descriptor ---------------------------------------------------------------
    def struct(self):
        P = self.pattern
        return _lib.MeshDev(
            P.n_vertices, P.n_faces, P.n_blocks, P.n_contrib,
            self.d_perm.data_ptr(), self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.d_diag.data_ptr(),
            self.d_cptr.data_ptr(), self.d_centry.data_ptr(), self.d_tri.data_ptr(),
            self.d_e.data_ptr(), self.d_grad_w.data_ptr(), self.d_integral.data_ptr(), self.d_areas.data_ptr(),
            self.d_a2v.data_ptr(), P.n_colors,
            (ctypes.c_int32 * (_lib.MAX_COLORS + 1))(*([int(x) for x in P.color_tile_ptr] + [0] * (_lib.MAX_COLORS - P.n_colors))),
            P.n_levels, self.level_stage_blocks, P.level_ptr.ctypes.data if P.n_levels else None,
            self.d_level_desc.data_ptr() if self.d_level_desc is not None else None)

    # -- reference-compatible views ----------------------------------------------------
    def tocsr(self):
        """a2 as scipy CSR (2N,2N) in the reference's indexing."""
        return self.pattern.block_values_to_csr(self.d_a2v.cpu().numpy())

    def toarray(self):
        return self.tocsr().toarray()

    def use_geometry(self, grad_w, e, integral_wi_wj, areas):
        """The reference's worker uses whatever grad_w / e / integral_wi_wj / areas it is
        handed (compute_optical_flow.py:100-101).  If the caller passes arrays other than
        the ones this handle returned, upload them (a2 values stay as they are)."""
        import torch
        P = self.pattern
        if grad_w is not None and grad_w is not self.grad_w and not np.array_equal(grad_w, self.grad_w):
            self.grad_w = _as_f64(grad_w, "grad_w", (3, 3))
            self.d_grad_w.copy_(torch.from_numpy(self.grad_w))
        if e is not None and e is not self.e and not np.array_equal(np.asarray(e).reshape(self.e.shape), self.e):
            self.e = _as_f64(np.asarray(e).reshape(-1, 2, 3), "e", (2, 3))
            self.d_e.copy_(torch.from_numpy(np.ascontiguousarray(self.e[P.perm])))
        if integral_wi_wj is not None and integral_wi_wj is not self.integral_wi_wj \
                and not np.array_equal(integral_wi_wj, self.integral_wi_wj):
            self.integral_wi_wj = _as_f64(integral_wi_wj, "integral_wi_wj", (2,))
            self.d_integral.copy_(torch.from_numpy(self.integral_wi_wj))
        if areas is not None and areas is not self.areas:
            ar = np.ascontiguousarray(np.asarray(areas, dtype=np.float64).reshape(-1))
            if not np.array_equal(ar, self.areas):
                self.areas = ar
                self.d_areas.copy_(torch.from_numpy(ar))

    @classmethod
    def from_reference_a2(cls, a2, coordinates, normals, triangles, areas, **kw):
        """Build a handle whose a2 block values are taken from a matrix produced by the
        reference's own compute_geometrical_quantities (scipy sparse, reference indexing)."""
        import torch
        op = cls(coordinates, normals, triangles, areas, **kw)
        vals = op.pattern.csr_to_block_values(a2)
        op.d_a2v.copy_(torch.from_numpy(np.ascontiguousarray(vals)))
        return op

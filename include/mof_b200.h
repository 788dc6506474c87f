/*
 * mof_b200.h -- C ABI of libmof_b200.so: the per-frame velocity-field solve of
 * SEU-dynamical-models/Manifold-based-optical-flow-method on NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI (it is pure Python); each entry
 * point below names the reference function (file:line under /root/reference) whose
 * arithmetic it replaces.  The Python package manifold_based_optical_flow_method_b200
 * binds these symbols with ctypes and re-exposes the reference's own signatures
 * (utils/compute_optical_flow.py, utils/find_singularity_point.py); INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every array argument is a raw pointer + sizes.
 *   - "host" pointers are ordinary CPU memory, "device" pointers are CUDA global
 *     memory of the current device.  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Kernels are asynchronous on that stream
 *     unless stated.
 *   - return value: 0 = OK, <0 = CUDA/runtime/argument error (text available from
 *     mof_last_error_string()), >0 = numerical status (see each function).
 *   - fp64 arithmetic, int32 indices.  The library keeps no global state besides a
 *     thread-local error string and a per-call ticket buffer inside the batch.
 *
 * Internal data layout
 *   - vertices are renumbered (breadth-first / Cuthill-McKee) for gather locality:
 *     internal vertex v  <->  reference vertex perm[v].
 *   - the 2N x 2N system matrix (row index i + N*alpha, compute_optical_flow.py:83)
 *     is stored as 2x2 blocks over the vertex-adjacency pattern: block b of row v
 *     couples internal vertices (v, col[b]); n_blocks = N + 2E.
 *   - frames are processed in groups of MOF_GROUP (=32) frames; every per-frame
 *     quantity is stored frame-minor: value[group][slot][MOF_GROUP], so a warp
 *     (lane = frame) reads/writes 256 contiguous bytes.
 */
#ifndef MOF_B200_H
#define MOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOF_GROUP 32          /* frames per group (= one warp, lane = frame)      */
#define MOF_TILE_ROWS 64      /* block rows per CTA tile in the SpMV/PCG kernels  */
#define MOF_MAX_COLORS 16     /* patch colours of the block-multicolour ordering   */
#define MOF_SCAL_SLOTS 20      /* per-frame scalar slots in mof_batch_dev.scal      */

/* numerical status per frame, written by mof_pcg_solve_batch */
#define MOF_STATUS_CONVERGED 0
#define MOF_STATUS_MAXITER 1      /* relres > tol after max_iter iterations        */
#define MOF_STATUS_BREAKDOWN 2    /* p'Ap <= 0 or NaN/Inf met (e.g. NaN input)      */
#define MOF_STATUS_ZERO_RHS 3     /* f == 0: V = 0 returned without iterating       */

const char* mof_last_error_string(void);
int mof_version(void);

/* ------------------------------------------------------------------------- *
 * Sparsity pattern (host, once per mesh).
 * Replaces the implicit pattern that lil_matrix element updates create in
 * compute_geometrical_quantities (compute_optical_flow.py:49,78-93) and worker
 * (:106,127-141).
 * ------------------------------------------------------------------------- */
typedef struct mof_pattern mof_pattern;     /* opaque host object */

/* triangles: host (n_faces,3) int64 vertex ids of the reference mesh.
 * reorder: 0 = identity (tests), 1 = Cuthill-McKee renumbering (block-Jacobi PCG),
 *          2 = block multicolour: patches of MOF_TILE_ROWS vertices (recursive coordinate
 *              bisection of `coords`, host (n_vertices,3) double, required), patch graph
 *              greedily coloured, numbering colour-major (SSOR sweeps run colour by colour,
 *              sequentially inside a patch).  coords may be NULL for reorder 0/1/3.
 *          3 = level-scheduled natural order: the Cuthill-McKee order regrouped by dependency level
 *              (level(v) = 1 + max level of the neighbours before v), numbering level-major.  Same
 *              Gauss-Seidel splitting as Cuthill-McKee, but the rows of a level are independent and
 *              contiguous: the SSOR sweeps run level by level, one warp per row.
 * Errors: vertex id out of range, a face with a repeated vertex. */
int mof_pattern_create(int64_t n_vertices, int64_t n_faces, const int64_t* triangles,
                       int reorder, const double* coords, mof_pattern** out);
/* n_colors (0 unless reorder = 2) and color_tile_ptr[n_colors+1]: tiles (= patches of
 * MOF_TILE_ROWS consecutive internal rows) [ptr[c], ptr[c+1]) carry colour c. */
int mof_pattern_colors(const mof_pattern* p, int32_t* n_colors, int32_t* color_tile_ptr);
/* Number of dependency levels (0 unless reorder = 3); level_ptr (may be NULL), n_levels+1 entries:
 * internal rows [ptr[l], ptr[l+1]) form level l. */
int32_t mof_pattern_levels(const mof_pattern* p, int32_t* level_ptr);
void mof_pattern_destroy(mof_pattern* p);
int64_t mof_pattern_num_blocks(const mof_pattern* p);    /* N + 2E                         */
int64_t mof_pattern_num_contrib(const mof_pattern* p);   /* 9 F (ordered vertex pairs)      */
int64_t mof_pattern_max_row_blocks(const mof_pattern* p);
int64_t mof_pattern_bandwidth(const mof_pattern* p);     /* max |v - col| after renumbering */
/* Copy the pattern into caller-allocated host arrays:
 *   perm   [N]      internal -> reference vertex id
 *   rowptr [N+1]    block-row pointers
 *   col    [nb]     internal column vertex of each block, ascending inside a row
 *   diag   [N]      index of the diagonal block of each row
 *   cptr   [nb+1]   per-block list of contributing faces ...
 *   centry [9F]     ... packed (face << 4 | m << 2 | n): face contributes its local
 *                   pair (m,n) (positions of the row/col vertex inside the face) to
 *                   that block; faces ascending, i.e. the reference's accumulation
 *                   order (compute_optical_flow.py:60,113)
 *   tri    [F*3]    triangles in internal vertex ids (face order unchanged)        */
int mof_pattern_export(const mof_pattern* p, int32_t* perm, int32_t* rowptr, int32_t* col,
                       int32_t* diag, int32_t* cptr, int32_t* centry, int32_t* tri);

/* ------------------------------------------------------------------------- *
 * Device-side descriptors (all pointers are device pointers).
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t n_vertices, n_faces, n_blocks, n_contrib;
    const int32_t *perm, *rowptr, *col, *diag, *cptr, *centry, *tri;
    const double* e;         /* (N,2,3)  tangent basis, internal vertex order           */
    const double* grad_w;    /* (F,3,3)                                                 */
    const double* integral;  /* (F,2)    A/6, A/12                                      */
    const double* areas;     /* (F,)                                                    */
    const double* a2v;       /* (nb,4)   a2 block values [2*alpha+beta], frame-shared   */
    /* block-multicolour ordering (reorder = 2), host-side launch metadata; n_colors = 0 otherwise */
    int32_t n_colors;
    int32_t color_tile_ptr[MOF_MAX_COLORS + 1];
    /* level-scheduled ordering (reorder = 3); n_levels = 0 otherwise.
     * level_ptr is the ONE HOST POINTER of this struct: n_levels+1 row offsets read by the host side of
     * mof_pcg_solve_batch only (launch ranges of the per-level fallback path); kernels never dereference it. */
    int32_t n_levels;
    int32_t level_stage_blocks;   /* persistent kernel: 3 = stage three matrix blocks per row in shared memory and keep three
                                     stages per warp (meshes where nearly every row has <= 3 blocks on either side of its
                                     diagonal); anything else = four blocks, two stages */
    const int32_t* level_ptr;
    /* device, [2][N][8] int32, filled by mof_level_desc_build (may be NULL: the solver then falls back to one
     * launch per dependency level): per row {first block, block count, first six columns} of the strictly
     * upper ([0], backward sweep) and strictly lower ([1], forward sweep) part of the row. */
    const int32_t* level_desc;
} mof_mesh_dev;

typedef struct {
    int32_t n_groups;        /* G                                                       */
    int32_t n_frames;        /* valid frames (<= G*MOF_GROUP); the rest is zero padding  */
    double* It;              /* [G][N][32]       I(t_k) per vertex                      */
    double* dIt;             /* [G][N][32]       (I(t_k+1) - I(t_k)) / dt               */
    double* vals;            /* [G][nb][4][32]   a1 + lambda*a2 block values            */
    double* rhs;             /* [G][N][2][32]    f                                      */
    double* minv;            /* [G][N][3][32]    block Jacobi: inverse of the diagonal 2x2 blocks;
                                                 SSOR: S = D^-1/2, the symmetric scaling applied to vals / rhs */
    double* x;               /* [G][N][2][32]    solution (tangent coefficients)        */
    double* r;               /* [G][N][2][32]                                           */
    double* z;               /* [G][N][2][32]                                           */
    double* p;               /* [G][N][2][32]                                           */
    double* ap;              /* [G][N][2][32]    A p (block-Jacobi) / forward-sweep result w (SSOR) */
    double* t;               /* [G][N][2][32]    SSOR only (may be NULL otherwise): backward-sweep result */
    double* partial;         /* [G][n_tiles][2][32] per-tile partial dot products       */
    double* scal;            /* [G][MOF_SCAL_SLOTS][32] per-frame scalars (r'z, p'Ap, r'r, ..., see csrc/mof_common.cuh) */
    int32_t* state;          /* [G][4][32]  active, iters, status, spare ; then [G] group_done, [G] tickets, groups_active, frames_active */
    int32_t* ready;          /* [G][N]      level-scheduled SSOR, persistent kernel: stamp of the last sweep that finished the row
                                            (may be NULL: per-level launches are used instead) */
} mof_batch_dev;

int64_t mof_num_tiles(int64_t n_vertices);                      /* ceil(N / MOF_TILE_ROWS) */
int64_t mof_state_ints(int32_t n_groups);                       /* size of `state` in int32 */

/* Optional sampled timing of the PCG kernels (caller-owned accumulator, may be NULL): one
 * iteration per check interval is bracketed with CUDA events on the solver's stream. */
typedef struct {
    double ms_spmv, ms_update, ms_pupdate;  /* summed device time of the sampled launches; SSOR:
                                               ms_spmv = backward sweeps, ms_pupdate = forward sweeps */
    int64_t samples;                        /* sampled iterations (one launch of each kernel)   */
    int64_t group_launches;                 /* sum over samples of groups still iterating       */
    int64_t frame_launches;                 /* sum over samples of frames still iterating       */
    int64_t iterations_total;               /* iterations launched by the call(s)               */
    int64_t launches_total;                 /* kernel launches issued by the call(s)            */
    /* level-scheduled SSOR, persistent kernel (level_iter_kernel: check_every whole iterations per launch) */
    double ms_iter;                         /* summed CUDA-event duration of the level_iter_kernel launches */
    int64_t iter_launches;                  /* number of those launches                          */
    double phase_ns[4];                     /* in-kernel %globaltimer split of ms_iter: backward sweeps, forward
                                               sweeps, p'Ap reduction, r update (each incl. its grid barrier) */
} mof_pcg_profile;

/* ------------------------------------------------------------------------- *
 * K0: geometry, once per mesh (compute_geometrical_quantities,
 * compute_optical_flow.py:27-97).
 * ------------------------------------------------------------------------- */
/* e[v] from normals[v] (compute_orthonormal_basis, :210-235); internal order in/out. */
int mof_geom_basis(int64_t n_vertices, const double* normals, double* e, void* stream);
/* grad_w (compute_gradient_w, :238-255, argument orders of :63-68) and
 * integral_wi_wj (:73-75) per face; coords (N,3) internal order, tri internal ids. */
int mof_geom_gradw(int64_t n_faces, const double* coords, const int32_t* tri,
                   const double* areas, double* grad_w, double* integral, void* stream);
/* a2 block values (compute_a2 :258-270 accumulated per :78-93) into mesh->a2v
 * (cast away const); uses mesh->e, grad_w, areas and the contributor lists. */
int mof_geom_a2(const mof_mesh_dev* mesh, double* a2v, void* stream);

/* ------------------------------------------------------------------------- *
 * K1: per-frame assembly for a batch (worker face loop, :113-141, and a = a1 +
 * lambda_*a2, :144).
 * ------------------------------------------------------------------------- */
/* Transpose/renumber frames into the frame-minor layout.  I_now/I_next: device,
 * row k = frame k of this batch, row stride ld (elements), reference vertex order.
 * Frame k uses I_now[k] for the gradient and as the old value and I_next[k] as the
 * new value (:174-175); dt[k] = t_k[k+1]-t_k[k] (:125), device (n_frames,). */
int mof_pack_frames(const mof_mesh_dev* mesh, const mof_batch_dev* batch, const double* I_now,
                    const double* I_next, int64_t ld, const double* dt, void* stream);
/* omega = 0 (block Jacobi): vals = a1(I) + lambda*a2, rhs = f, minv = inverse of the diagonal
 * 2x2 blocks.  omega in (0,2) (SSOR): the system is additionally scaled symmetrically with
 * S = D^-1/2 (D = diagonal blocks): vals = S (a1 + lambda*a2) S (identity diagonal blocks),
 * rhs = S f, minv = S; mof_pcg_solve_batch undoes the scaling (x = S xh) and judges convergence
 * on the unscaled residual. */
int mof_assemble_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double lambda_,
                       double omega, void* stream);

/* ------------------------------------------------------------------------- *
 * K2/K3: batched block-Jacobi PCG (replaces spsolve, :147).
 * ------------------------------------------------------------------------- */
/* Row descriptors of the level-scheduled ordering for the persistent SSOR kernel: fills desc (device,
 * 2*N*8 int32) from mesh->rowptr/col/diag; store the pointer in mesh->level_desc afterwards.  Synchronous.
 * Returns 1 (and leaves desc unusable: keep level_desc NULL) if a row has more than 32 blocks on one side
 * of its diagonal. */
int mof_level_desc_build(const mof_mesh_dev* mesh, int32_t* desc, void* stream);
/* y = A x for every group of the batch (x, y in the [G][N][2][32] layout).  Test and
 * roofline hook for the SpMV kernel that the solver uses. */
int mof_spmv_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, const double* x,
                   double* y, void* stream);
/* Solve A x = rhs for all frames of the batch, x0 = 0, until ||r||/||b|| <= tol
 * (recurrence residual, confirmed on the true residual b - A x; at most
 * max_restarts restarts from the true residual).
 * omega = 0: block-Jacobi PCG (SpMV + two vector kernels per iteration).
 * omega in (0,2): SSOR PCG in Eisenstat's form (needs batch->minv assembled with the same omega): per
 * iteration one backward and one forward block-triangular sweep plus one vector pass.  With a mesh
 * built with reorder = 3 (mesh->n_levels > 0), mesh->level_desc and batch->ready set, ONE persistent cooperative
 * kernel runs check_every whole iterations per launch (row-level dataflow inside the sweeps, bulk-async prefetch;
 * environment MOF_LEVEL_PERSIST=0, a missing descriptor / stamp buffer or a device without cooperative launch:
 * one launch per dependency level, replayed as a CUDA graph -- same bits); with reorder = 2 (mesh->n_colors > 0)
 * the sweeps run colour by colour, one warp per patch.  ~3x (multicolour) to ~8x (level-scheduled) fewer
 * iterations than block Jacobi at the same bytes per iteration.  Synchronous: returns after the
 * stream has drained.  Host outputs (each MOF_GROUP*n_groups long, may be NULL):
 * iters, relres (true residual), status (MOF_STATUS_*).  Return value: 0 if every
 * valid frame converged (or had a zero rhs), else the largest status met. */
int mof_pcg_solve_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double tol,
                        double omega, int32_t max_iter, int32_t check_every, int32_t max_restarts,
                        int32_t* iters, double* relres, int32_t* status, mof_pcg_profile* prof,
                        void* stream);
/* Which code path the calling thread's last mof_pcg_solve_batch took (tests assert that the default is the
 * persistent kernel and not a silent fallback).  Returns MOF_PATH_*; info (may be NULL) receives
 * {path, grid size of the iteration kernel, its CTAs per SM, reason of a fallback to per-level launches:
 * 0 none, 1 mesh->level_desc NULL, 2 batch->ready NULL, 3 too many groups, 4 switched off or N*G too large,
 * 5 no cooperative launch, 6 kernel does not fit an SM}. */
#define MOF_PATH_JACOBI 0
#define MOF_PATH_MULTICOLOUR 1
#define MOF_PATH_LEVEL_LAUNCHES 2
#define MOF_PATH_LEVEL_PERSISTENT 3
int mof_pcg_last_path(int32_t* info);
/* x -> V[k][i + N*alpha] (reference order and layout, :149); V: device, row stride ld. */
int mof_unpack_solution(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double* V,
                        int64_t ld, void* stream);

/* ------------------------------------------------------------------------- *
 * K4: tangent coefficients -> xyz (process_V_k, find_singularity_point.py:28-69),
 * speed |V| (S3_compute_v_and_detection_singularity.py:130-132) and per-frame
 * v_length_max (find_singularity_point.py:161-162).
 * V (n_frames, 2N) row stride ldV; e (N,2,3) REFERENCE vertex order;
 * Vxyz (n_frames,N,3); speed (n_frames,N) or NULL; vmax (n_frames,) or NULL.
 * ------------------------------------------------------------------------- */
int mof_tangent_to_xyz(int64_t n_vertices, int64_t n_frames, const double* V, int64_t ldV,
                       const double* e, double* Vxyz, double* speed, double* vmax, void* stream);
/* vmax only, from an existing (n_frames,N,3) field. */
int mof_vmax(int64_t n_vertices, int64_t n_frames, const double* Vxyz, double* vmax, void* stream);

/* ------------------------------------------------------------------------- *
 * K5: singularity detection (find_singularity_points,
 * find_singularity_point.py:140-189).  Two calls with an exact-size allocation by
 * the caller in between.
 * ------------------------------------------------------------------------- */
#define MOF_DETECT_CHUNK 1024
/* Pass 1: vflag[k][i] = ||V_i/vmax|| <= eps (:72-90,:165-167); fflag[k][t] = face t
 * holds an interior zero (:93-137,:170-180) and has no singular vertex (:171).
 * Also writes per-chunk counts: vcnt [n_frames][ceil(N/1024)], fcnt
 * [n_frames][ceil(F/1024)] and their per-frame exclusive scans in place (so that
 * on return vcnt/fcnt hold chunk offsets) and totals[k] = {n_vertices_k, n_faces_k}.
 * coords (N,3), tri (F,3) int32: REFERENCE vertex ids. */
int mof_singularity_flags(int64_t n_vertices, int64_t n_faces, int64_t n_frames,
                          const double* coords, const int32_t* tri, const double* Vxyz,
                          const double* vmax, double eps, uint8_t* vflag, uint8_t* fflag,
                          int32_t* vcnt, int32_t* fcnt, int32_t* totals, void* stream);
/* Pass 2: ordered compaction.  voff/foff (n_frames,) int64: start of frame k inside
 * the output lists (exclusive scan of totals, computed by the caller).
 * Outputs: vertex_idx [sum nv]; face_idx [sum nf]; lam_mu [sum nf][2];
 * P [sum nf][3] (:181-182); index [sum nf] int8 per-face Poincare index (+1/-1). */
int mof_singularity_compact(int64_t n_vertices, int64_t n_faces, int64_t n_frames,
                            const double* coords, const int32_t* tri, const double* Vxyz,
                            const double* vmax, const uint8_t* vflag, const uint8_t* fflag,
                            const int32_t* vcnt, const int32_t* fcnt, const int64_t* voff,
                            const int64_t* foff, int32_t* vertex_idx, int32_t* face_idx,
                            double* lam_mu, double* P, int8_t* index, void* stream);

/* ------------------------------------------------------------------------- *
 * K5b ("next" row 2 of SURVEY 8f): Jacobian classification of the detected critical points
 * (compute_jacobian_matrix_for_vertex / _for_interior, classify_critical_point,
 * find_singularity_point.py:355-498) for the output of mof_singularity_compact.
 * e (N,2,3) REFERENCE vertex order; ring_ptr/ring_idx: ascending 1-ring neighbour lists
 * (pyvista point_neighbors, fsp:375); face_nbr (F,3): face across edge AB / BC / CA of each face,
 * -1 on the boundary ("the other triangle on the nearest edge", fsp:432-438); voff/foff
 * (n_frames+1) offsets of each frame inside the point lists.  Outputs per point: the 2x2
 * matrix the reference accumulates (row-major) and its class 0 Node, 1 Focus, 2 Saddle,
 * 3 Indeterminate.
 * ------------------------------------------------------------------------- */
int mof_classify_singularities(int64_t n_vertices, int64_t n_faces, int64_t n_frames, const double* coords,
                               const int32_t* tri, const double* Vxyz, const double* vmax, const double* e,
                               const int32_t* ring_ptr, const int32_t* ring_idx, const int32_t* face_nbr,
                               const int64_t* voff, const int64_t* foff, int64_t n_vertex_points,
                               int64_t n_face_points, const int32_t* vertex_idx, const int32_t* face_idx,
                               const double* P, double* jac_v, int8_t* cls_v, double* jac_f, int8_t* cls_f,
                               void* stream);

/* ------------------------------------------------------------------------- *
 * K7 ("next" row 3 of SURVEY 8f): multi-ring winding numbers of singular points
 * (calculate_winding_numbers, S7_winding_line.py:120-165; angle_between_vectors / winding_number
 * :59-87; polar ordering :93-102).  For each of the n_points points (points (n,3);
 * frame_of_point (n,) selects the velocity field Vxyz[frame] of shape (N,3)): the closest mesh
 * vertex (S7:130), then for ring level 0 .. max_level-1 around it (breadth-first topological
 * rings over ring_ptr / ring_idx, pyvista point_neighbors_levels, S7:131) the winding number of
 * the velocity field along the ring ordered by polar angle in the vertex's tangent basis
 * e (N,2,3).  counts = number of consecutive rings accepted by the reference's rule (first ring
 * within 0.01 of +1 or -1 fixes the type, later rings within 0.001 of it); types = +1 / -1 / 0;
 * winding (n, max_level), optional: the winding numbers evaluated, NaN after the stop;
 * status: 0 ok, 2 a ring held more than 1024 vertices (count stops there).  Where the mesh runs
 * out of rings the count stops (the reference raises IndexError).  All REFERENCE vertex order.
 * ------------------------------------------------------------------------- */
int mof_winding_numbers(int64_t n_vertices, int64_t n_frames, const double* coords, const double* Vxyz,
                        const double* e, const int32_t* ring_ptr, const int32_t* ring_idx, int64_t n_points,
                        const double* points, const int32_t* frame_of_point, int max_level, int32_t* closest,
                        int32_t* counts, int8_t* types, double* winding, int32_t* status, void* stream);

/* ------------------------------------------------------------------------- *
 * K8 ("next" row 4 of SURVEY 8f, the producer of the hot path's input): radial-basis interpolation
 * of electrode signals onto the mesh vertices -- `interpolation`, S2_interpolate.py:22-53 and
 * S2_interpolate_phases.py:22-56, i.e. scipy.interpolate.Rbf(x, y, z, d)(vx, vy, vz) with its
 * defaults: multiquadric phi(r) = sqrt((r/epsilon)^2 + 1), smooth 0, Euclidean norm.  epsilon is
 * the caller's (the scipy default is (prod(bounding-box edges)/m)^(1/n_edges)).
 *
 * mof_rbf_fit: builds the (m,m) matrix phi(|c_i - c_j|) into lu, factorises it in place (LU, partial
 * pivoting, piv (m,) LAPACK-style) and solves it for n_rhs right-hand sides data (n_rhs, m) row
 * stride ld into weights (m, n_rhs) (right-hand side minor).  info (device int32): 0, or k+1 if
 * pivot k is exactly zero (singular: duplicated electrodes).  Complex data: pass real parts as
 * rows 0..T-1 and imaginary parts as rows T..2T-1 (n_rhs = 2T).
 * mof_rbf_evaluate: out (n_frames, N) row stride ld, out[t][n] = sum_j weights[j][t] phi(|v_n - c_j|);
 * phase_mode 1: weights hold 2*n_frames columns and out = atan2(imaginary, real) (np.angle,
 * S2_interpolate_phases.py:52).
 * ------------------------------------------------------------------------- */
int mof_rbf_fit(int64_t n_centres, const double* centres, double epsilon, int64_t n_rhs, const double* data,
                int64_t ld, double* lu, int32_t* piv, double* weights, int32_t* info, void* stream);
int mof_rbf_evaluate(int64_t n_vertices, int64_t n_centres, int64_t n_frames, const double* vertices,
                     const double* centres, double epsilon, const double* weights, int phase_mode, double* out,
                     int64_t ld, void* stream);

/* ------------------------------------------------------------------------- *
 * K6 ("next" row of SURVEY 8f): wave speed of S5_compute_wave_v.py.
 * I: device (n_frames, N) row stride ld, REFERENCE vertex order (phases in (-pi,pi] or
 * potentials); dt = 1/SF.  phase_mode 1: wave_velocity_phase (S5:79-123, wrapped time
 * differences S5:60-77); 0: wave_velocity_amplitude (S5:14-58, np.gradient edge_order=2,
 * needs n_frames >= 3).  Outputs (either may be NULL), reference vertex order:
 * grad_point (n_frames,N,3) = compute_grad_M_I (S5:136-171); wave (n_frames,N) = time
 * derivative / |projected gradient| (S5:121), signed and unscaled like the functions'
 * return value (the script divides by 1000 and takes abs afterwards, S5:312-313).
 * The derivative couples neighbouring frames, so a call names the rows it was given and the rows it must
 * produce: I holds n_rows consecutive frames of a trial of T_trial frames, the first being frame t_first;
 * outputs are written for rows out0 .. out0+n_out-1 of the call only (grad_point (n_out,N,3), wave (n_out,N));
 * the other rows are halo (one frame either side; two at a trial end in amplitude mode, where np.gradient's
 * one-sided formula applies).  Whole trial on one GPU: n_rows = n_out = T_trial, out0 = t_first = 0.  Frames
 * sharded over GPUs (config 5): every rank passes its range plus the halo -- no collective on the data path.
 * work: 16-byte aligned device scratch of mof_wave_work_doubles(mesh, n_rows, grad_point != NULL, wave != NULL) doubles: the
 * signal in the frame-minor layout [group][internal vertex][32 frames], the two time neighbours of every group per
 * vertex, and the coefficient rows (CSR-aligned and padded to eight slots per vertex) (two / three
 * doubles per block of the mesh pattern) that turn the per-frame work into one sparse row product per vertex
 * (csrc/wave.cu); the results go straight into grad_point / wave.  The kernels work for any mesh ordering; the
 * transposes are fully coalesced when the mesh was built with reorder = 0 (perm = identity), which is what
 * S5_compute_wave_v.py does.
 * ------------------------------------------------------------------------- */
int64_t mof_wave_work_doubles(const mof_mesh_dev* mesh, int64_t n_rows, int want_grad, int want_wave);
int mof_wave_speed(const mof_mesh_dev* mesh, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                   int64_t T_trial, const double* I, int64_t ld, double dt, int phase_mode, double* grad_point,
                   double* wave, double* work, void* stream);
/* The row kernel(s) alone on a work buffer that a previous mof_wave_speed call with the same mesh, n_rows and
 * outputs has filled (roofline hook of bench.py, like mof_spmv_batch). */
int mof_wave_stencil(const mof_mesh_dev* mesh, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                     int64_t T_trial, double dt, int phase_mode, double* grad_point, double* wave, double* work,
                     void* stream);

/* Tuning knob: variant of the wave-speed row kernel = (32-frame groups per CTA pass, CTAs per SM compiled for):
 * 0: (1,4)  1: (2,3)  2: (1,5)  3: (2,4)  4: (2,5)  5: (3,4)  6: (4,3).  The environment variable MOF_WAVE_VARIANT sets
 * it at first use; the default is the fastest measured at config 5.  Results are bit-identical. */
int mof_wave_set_variant(int variant);
int mof_wave_get_variant(void);

/* ------------------------------------------------------------------------- *
 * On-disk formats either side of the path ("next" row 4 of SURVEY 8f), host only, multi-threaded.
 * The pandas CSV dialect of load_potentials (pd.read_csv(..., header='infer', index_col=0),
 * compute_optical_flow.py:203-207) and reshape_and_save_data (pd.DataFrame(a).to_csv(path),
 * :314-320): header ",0,1,...", rows "r,v0,v1,...", doubles printed like Python's repr (shortest
 * round trip), NaN as an empty field.  n_threads <= 0: all hardware threads.
 * ------------------------------------------------------------------------- */
int mof_csv_write(const char* path, const double* data, int64_t rows, int64_t cols, int n_threads);
int mof_csv_dims(const char* path, int64_t* rows, int64_t* cols);    /* data rows, value columns */
int mof_csv_read(const char* path, double* data, int64_t rows, int64_t cols, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* MOF_B200_H */

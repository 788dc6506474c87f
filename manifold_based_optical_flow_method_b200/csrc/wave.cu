// K6 -- wave speed from phase / amplitude maps (S5_compute_wave_v.py): per vertex and frame
//   grad_point = area-weighted mean of the per-face surface gradients of the incident faces
//                (compute_grad_M_I, S5:136-171, same hat-function "gradients" as cof:238-255),
//   its tangent-plane projection expressed in the basis (e1, e2) and the norm of that
//   (project_vector_to_plane S5:173-180, express_vector_on_basis S5:182-191, S5:96-117),
//   the time derivative: wrapped differences for phases (compute_temporal_gradient_phase,
//   S5:60-77, angle_subtract S5:224-233) or np.gradient(edge_order=2) for amplitudes (S5:24),
//   wave_velocity = time derivative / norm (S5:121 / S5:56).
//
// Formulation.  Everything between the signal and (alpha, beta) is LINEAR in the signal and depends on the mesh
// only: grad_point[v] = sum over the 1-ring u (v included) of cg[v,u] I[u], with
//     cg[v,u] = (1 / sum of A_f over the faces of v) * sum over the faces f that hold v and u of A_f grad_w[f][position of u]
// and (alpha, beta)[v] = sum_u cw[v,u] I[u], cw[v,u] = the tangent coefficients of the projected cg[v,u].  The pairs
// (v,u) are exactly the blocks of the solver's block-CSR pattern (rowptr / col / per-block face lists), so the
// coefficients are two or three doubles per block, computed once per call by wave_coef_kernel (faces ascending, the
// order of S5:161-166), and the per-frame work is one sparse row product: 7 entries x 2 FMAs per vertex and frame
// instead of six face gradients, a projection and two basis coefficients (~120 fp64 instructions in round 2's first
// kernel, which made it fp64-bound at 0.08 of the HBM peak).  The sums are re-associated with respect to the
// reference (per ring vertex instead of per face); measured difference from the unmodified reference 5e-14 rel-L2 at
// most on the parity cases (tests/test_wave_speed.py asserts <= 1e-12 / 1e-13).
//
// Layout.  The (T,N) signal is transposed once into the frame-minor layout of the solver (It[group][vertex][32 frames],
// wave_pack_kernel, which also leaves every group's two time neighbours per vertex in Ih).  wave_rows_kernel: a CTA
// owns 32 consecutive vertices of one 32-frame group; it stages the padded column indices and coefficients of its 32
// block rows in shared memory with one coalesced pass, then every warp walks four vertices with lane = frame -- each
// ring value is one 256-byte line that 32 frames share, the time neighbours are the adjacent lanes -- and the tile is
// transposed through shared memory so that the result is written straight into the caller's (T,N[,3]) array as
// 256-byte (768-byte) row pieces: no frame-minor result, no transpose kernel behind it.
// Algorithmic HBM bytes: 8 N read + 8 N written per frame (the pack moves another 16 N); the ring re-reads are served
// by L1 (inside the CTA's tile) and L2 (a group's It is N x 256 bytes = 42 MB at 164k vertices).
// The wave operator keeps the mesh in REFERENCE vertex order (S5_compute_wave_v._operator: reorder = 0), which
// makes both transposes fully coalesced; the kernels honour mesh->perm all the same.
// A call may cover a SHARD of a trial: rows outside [out0, out0 + n_out) are halo for the time derivative, and the
// one-sided end formulas apply at the trial's ends only (t_first, T_trial).
// Measured at config 5 (1000 frames x 163,842 vertices, profiles/r2_wave_probe.json, r2_ncu_summary.md section 5):
// pack 0.63 ms, row kernel 1.63 ms (round 2's first version: stencil 4.89 + transpose out 1.23 ms); with one group per
// pass the row kernel issues 238 instructions and 56 L1 wavefronts per (vertex, 32 frames) and runs with the issue slots
// and the L1 data pipe both about 60 % busy -- neither DRAM (26 % active) nor L2 limits it; three groups per pass
// (the default) share the shared-memory reads of the block rows and shave another 7 %.
#include <stdlib.h>

#include "mof_common.cuh"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kTileVerts = 32;        // vertices per CTA of wave_rows_kernel
constexpr int kSlots = 8;            // block-row entries per vertex staged in shared memory (valence <= 7; longer rows continue in global)

// (rows, N) row-major, reference vertex order -> It[g][v][32], internal order (rows padded with 0), and the time halo
// of every group Ih[g][v] = {row 32 g - 1, row 32 g + 32} (0 outside the rows of the call): the neighbours in time of a
// group's first and last frame, which the row kernel would otherwise fetch as two 32-byte sectors of other groups' lines
__global__ void __launch_bounds__(256) wave_pack_kernel(int64_t N, int64_t n_rows, const int32_t* __restrict__ perm,
                                                        const double* __restrict__ I, int64_t ld, double* __restrict__ It,
                                                        double* __restrict__ Ih) {
    __shared__ double s[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t g = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    const int64_t v = v0 + tx;
    const int64_t o = v < N ? perm[v] : 0;
    for (int fr = ty; fr < 32; fr += 8) {
        const int64_t k = g * 32 + fr;
        s[fr][tx] = (k < n_rows && v < N) ? __ldcs(I + k * ld + o) : 0.0;          // read once: evict first
    }
    if (ty < 2 && v < N) {
        const int64_t k = ty == 0 ? g * 32 - 1 : g * 32 + 32;
        Ih[(g * N + v) * 2 + ty] = (k >= 0 && k < n_rows) ? I[k * ld + o] : 0.0;
    }
    __syncthreads();
    for (int vv = ty; vv < 32; vv += 8) {
        const int64_t w = v0 + vv;
        if (w < N) It[mof_ix_sca(N, g, w) + tx] = s[tx][vv];
    }
}

// Coefficient rows aligned with the block pattern: cg[j][3] and / or cw[j][2] for block j = (row v, column u), and the
// same rows padded to kSlots slots per vertex for the row kernel's one-pass staging: pcol[v][8], pw[v][8][2], pg[v][8][4]
// (a slot past the end of a row: the vertex itself, zero coefficients; entries beyond kSlots stay in the CSR arrays).
__global__ void __launch_bounds__(128) wave_coef_kernel(mof_mesh_dev M, double* __restrict__ cw, double* __restrict__ cg,
                                                        int32_t* __restrict__ pcol, double* __restrict__ pw, double* __restrict__ pg) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M.n_vertices) return;
    mof_wave_coef_row_body(M, v, cw, cg);
    const int32_t j0 = M.rowptr[v], cnt = M.rowptr[v + 1] - j0;
    for (int k = 0; k < kSlots; ++k) {
        const bool on = k < cnt;
        const size_t j = (size_t)j0 + k, p = (size_t)v * kSlots + k;
        pcol[p] = on ? M.col[j] : (int32_t)v;
        if (pw) { pw[2 * p] = on ? cw[2 * j] : 0.0; pw[2 * p + 1] = on ? cw[2 * j + 1] : 0.0; }
        if (pg) { pg[4 * p] = on ? cg[3 * j] : 0.0; pg[4 * p + 1] = on ? cg[3 * j + 1] : 0.0; pg[4 * p + 2] = on ? cg[3 * j + 2] : 0.0; pg[4 * p + 3] = 0.0; }
    }
}

// One sparse row product per (vertex, frame).  C = 2: coefficients (alpha, beta) -> wave speed (T,N);
// C = 3: coefficients of grad_point -> (T,N,3).  GP = 32-frame groups a CTA takes in one pass (blockIdx.y counts
// passes): the column indices and coefficients read from shared memory then serve GP groups.  MINB: CTAs per SM the
// kernel is compiled for.
// The tile's block rows arrive as kSlots padded slots per vertex (wave_coef_kernel wrote them that way: thread =
// (row, slot), one coalesced pass of 16-byte loads that depends on nothing): a slot past the end of a row holds the
// vertex itself with zero coefficients, so the eight products of a vertex are straight-line code with no selects; the
// entries of a row beyond kSlots (valence > 7) are read from the CSR arrays behind them, in the same ascending column
// order.
template <int C, int GP, int MINB>
__global__ void __launch_bounds__(256, MINB) wave_rows_kernel(
    int64_t N, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
    const double* __restrict__ coef, const int32_t* __restrict__ pcol, const double* __restrict__ pcoef, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first, int64_t T_trial,
    const double* __restrict__ It, const double* __restrict__ Ih, double inv_dt, int phase_mode, double* __restrict__ out) {
    constexpr int OC = C == 3 ? 3 : 1;                 // doubles written per (vertex, frame)
    constexpr int LD = kTileVerts * OC + 1;            // odd row length: both sides of the transpose are conflict-free
    constexpr int CP = C == 3 ? 4 : 2;                 // doubles per staged slot: one (two) 16-byte shared loads
    __shared__ double s_out[GP * 32 * LD];
    __shared__ __align__(16) double s_coef[kTileVerts * kSlots * CP];
    __shared__ __align__(16) int32_t s_col[kTileVerts * kSlots];
    __shared__ double s_halo[GP * kTileVerts * 2];
    __shared__ int32_t s_j0[kTileVerts], s_cnt[kTileVerts];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t v0 = (int64_t)blockIdx.x * kTileVerts;
    const int64_t g0 = (int64_t)blockIdx.y * GP;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    // a pass that reaches past the last group computes its first group twice and writes it once
    int64_t grp[GP];
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) grp[gp] = g0 + gp < G ? g0 + gp : g0;
    {   // ---- stage the tile's padded block rows (thread = (row, slot), 16-byte loads, nothing depends on anything) ...
        const size_t p = (size_t)v0 * kSlots + tid;
        const bool on = p < (size_t)N * kSlots;
        s_col[tid] = on ? pcol[p] : 0;
        const double2* src = reinterpret_cast<const double2*>(pcoef + p * CP);
        double2* dst = reinterpret_cast<double2*>(s_coef + tid * CP);
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) dst[q] = on ? src[q] : make_double2(0.0, 0.0);
        // ... the row lengths (a row longer than kSlots continues in the CSR arrays) ...
        if (tid < kTileVerts) {
            const int64_t v = v0 + tid;
            const int32_t j0 = v < N ? rowptr[v] : 0;
            s_j0[tid] = j0;
            s_cnt[tid] = v < N ? rowptr[v + 1] - j0 : 0;
        }
        // ... and the time halo of its vertices (C = 2 only reads it)
        if (C == 2 && tid < GP * kTileVerts * 2) {
            const int gp = tid / (kTileVerts * 2), idx = tid - gp * kTileVerts * 2;
            const int64_t gg = g0 + gp < G ? g0 + gp : g0;
            s_halo[tid] = v0 + (idx >> 1) < N ? Ih[(gg * N + v0) * 2 + idx] : 0.0;
        }
    }
    // position of this thread's vertex column in the caller's array (transposed write below), requested early
    const int64_t o_out = v0 + lane < N ? perm[v0 + lane] : 0;
    __syncthreads();

    const double* It_l[GP];                         // this lane's frame of group g0 + gp: vertex u at It_l[gp][u * 32]
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) It_l[gp] = It + mof_ix_sca(N, grp[gp], 0) + lane;

    // the ring lines (and, for the time derivative, the vertex's own line) of tile vertex vv
    auto request = [&](int vv, double (&val)[GP][kSlots], double (&cur)[GP]) {
        if (v0 + vv >= N) return;                                           // warp-uniform
        const int4 ca = *reinterpret_cast<const int4*>(s_col + vv * kSlots), cb = *reinterpret_cast<const int4*>(s_col + vv * kSlots + 4);
        const int32_t us[kSlots] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int k = 0; k < kSlots; ++k)
#pragma unroll
            for (int gp = 0; gp < GP; ++gp) val[gp][k] = It_l[gp][(size_t)us[k] * MOF_W];
        if (C == 2) {
#pragma unroll
            for (int gp = 0; gp < GP; ++gp) cur[gp] = It_l[gp][(size_t)(v0 + vv) * MOF_W];
        }
    };
    auto compute = [&](int vv, const double (&val)[GP][kSlots], const double (&cur)[GP]) {
        const int64_t v = v0 + vv;
        if (v >= N) return;                                                 // warp-uniform
        double acc[GP][C];
#pragma unroll
        for (int gp = 0; gp < GP; ++gp)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[gp][c] = 0.0;
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const double* wk = s_coef + (vv * kSlots + k) * CP;
            double w[C];
            if (C == 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(wk);
                w[0] = t2.x; w[1] = t2.y;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) w[c] = wk[c];
            }
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int gp = 0; gp < GP; ++gp) acc[gp][c] = fma(w[c], val[gp][k], acc[gp][c]);
        }
        const int cnt = s_cnt[vv];
        if (cnt > kSlots) {                                                 // valence > 7: the rest of the row, same order
            const int32_t j0 = s_j0[vv];
#pragma unroll 1
            for (int k = kSlots; k < cnt; ++k) {
                const size_t u = (size_t)col[j0 + k] * MOF_W;
#pragma unroll
                for (int gp = 0; gp < GP; ++gp) {
                    const double x = It_l[gp][u];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[gp][c] = fma(coef[(size_t)(j0 + k) * C + c], x, acc[gp][c]);
                }
            }
        }
#pragma unroll
        for (int gp = 0; gp < GP; ++gp) {
            double* so = s_out + gp * 32 * LD + lane * LD + vv * OC;
            if (C == 3) {
#pragma unroll
                for (int c = 0; c < OC; ++c) so[c] = acc[gp][c];
            } else {
                // time derivative: the neighbours in time are the adjacent lanes, the staged halo at the ends of the group
                const int64_t r = (g0 + gp) * 32 + lane;                    // row of the call
                const int64_t t = t_first + r;                              // frame of the trial
                const bool first = t == 0, last_t = t == T_trial - 1;
                double prev = __shfl_up_sync(kFullMask, cur[gp], 1), next = __shfl_down_sync(kFullMask, cur[gp], 1);
                if (lane == 0) prev = s_halo[(gp * kTileVerts + vv) * 2];
                if (lane == 31) next = s_halo[(gp * kTileVerts + vv) * 2 + 1];
                double far2 = 0.0;                                          // np.gradient's one-sided ends reach two frames in
                if (!phase_mode && r < n_rows && (first || (last_t && r >= 2))) {
                    const int64_t row = first ? r + 2 : r - 2;
                    far2 = It[mof_ix_sca(N, row >> 5, v) + (row & 31)];
                }
                so[0] = mof_wave_speed_body(mof_wave_td_body(phase_mode, first, last_t, T_trial, cur[gp], prev, next, far2, inv_dt),
                                            acc[gp][0], acc[gp][1 % C]);
            }
        }
    };
    {
        const int vb = warp * 4;                                            // a warp walks four vertices of the tile
        double va[GP][kSlots], ca[GP];
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            request(vb + i, va, ca);
            compute(vb + i, va, ca);
        }
    }
    __syncthreads();
    // ---- transposed write: thread column tx walks the tile's doubles of one frame row, 8 rows per pass.  Streaming
    // stores: the result is not read again, and a group's It lines (N x 256 bytes) should stay in L2 for the ring
    // re-reads of the tiles that follow.
    const int tx = lane, ty = warp;
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) {
        if (g0 + gp >= G) break;
#pragma unroll
        for (int fr = ty; fr < 32; fr += 8) {
            const int64_t k = (g0 + gp) * 32 + fr - out0;
            if (k < 0 || k >= n_out) continue;
            if (C == 2) {
                if (v0 + tx < N) __stcs(out + (size_t)k * N + o_out, s_out[gp * 32 * LD + fr * LD + tx]);
            } else {
#pragma unroll
                for (int c = 0; c < OC; ++c) {
                    const int idx = tx + 32 * c;                            // position inside the tile's row piece
                    const int vv = idx / OC;
                    if (v0 + vv < N) __stcs(out + ((size_t)k * N + perm[v0 + vv]) * OC + (idx - vv * OC), s_out[gp * 32 * LD + fr * LD + idx]);
                }
            }
        }
    }
}

struct wave_work {
    double *It, *Ih, *cw, *cg, *pw, *pg;
    int32_t* pcol;
    int64_t total;
};

// work = the packed signal It[G][N][32], its time halo Ih[G][N][2], the padded column indices pcol[N][8] (int32), then
// per output asked for the CSR-aligned and the padded coefficient rows: cw[nb][2], pw[N][8][2] (wave speed),
// cg[nb][3] (+1 pad), pg[N][8][4] (grad_point).  Every piece starts on a 16-byte boundary.
wave_work wave_layout(const mof_mesh_dev* mesh, int64_t n_rows, bool want_grad, bool want_wave, double* work) {
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W, N = mesh->n_vertices, nb = mesh->n_blocks;
    int64_t off = 0;
    auto take = [&](int64_t doubles, bool on) {
        double* p = on && work ? work + off : nullptr;
        if (on) off += (doubles + 1) & ~int64_t(1);
        return p;
    };
    wave_work w;
    w.It = take(G * N * MOF_W, true);
    w.Ih = take(G * N * 2, true);
    w.pcol = reinterpret_cast<int32_t*>(take(N * kSlots / 2, true));
    w.cw = take(2 * nb, want_wave);
    w.pw = take(N * kSlots * 2, want_wave);
    w.cg = take(3 * nb, want_grad);
    w.pg = take(N * kSlots * 4, want_grad);
    w.total = off;
    return w;
}

// Variants of the wave-speed row kernel = (32-frame groups per CTA pass, CTAs per SM it is compiled for):
//   0: (1, 4)   1: (2, 3)   2: (1, 5)   3: (2, 4)   4: (2, 5)   5: (3, 4)   6: (4, 3)
// More groups per pass: fewer instructions and shared-memory reads per (vertex, frame), more ring lines in flight per warp,
// a larger L2 working set; more CTAs per SM: fewer registers per thread.  MOF_WAVE_VARIANT or mof_wave_set_variant()
// select one (results are bit-identical); the default is the fastest measured at config 5 (profiles/r2_wave_probe.json).
constexpr int kWaveVariants = 7;
constexpr int kWaveVariantDefault = 5;      // row kernel, 1000 frames x 163,842 vertices: 1.93 1.83 1.75 1.67 1.76 1.63 1.66 ms for variants 0 .. 6
int g_wave_variant = -1;
int wave_variant() {
    if (g_wave_variant < 0) {
        const char* e = getenv("MOF_WAVE_VARIANT");
        g_wave_variant = e && e[0] >= '0' && e[0] < '0' + kWaveVariants && !e[1] ? e[0] - '0' : kWaveVariantDefault;
    }
    return g_wave_variant;
}

int wave_rows_launch(const mof_mesh_dev* mesh, const wave_work& w, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                     int64_t T_trial, double dt, int phase_mode, double* grad_point, double* wave, cudaStream_t st) {
    const int64_t N = mesh->n_vertices;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    auto grid = [&](int gp) { return dim3(mof_cdiv(N, kTileVerts), (unsigned)((G + gp - 1) / gp)); };
#define MOF_WAVE_ARGS(coef_, pcoef_, out_) \
    N, mesh->rowptr, mesh->col, mesh->perm, coef_, w.pcol, pcoef_, n_rows, out0, n_out, t_first, T_trial, w.It, w.Ih, 1.0 / dt, phase_mode, out_
    if (grad_point) {
        wave_rows_kernel<3, 1, 4><<<grid(1), 256, 0, st>>>(MOF_WAVE_ARGS(w.cg, w.pg, grad_point));
        MOF_LAUNCH_CHECK("wave_rows_kernel<3,1,4>");
    }
    if (wave) {
        switch (wave_variant()) {
        case 0: wave_rows_kernel<2, 1, 4><<<grid(1), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        case 1: wave_rows_kernel<2, 2, 3><<<grid(2), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        case 2: wave_rows_kernel<2, 1, 5><<<grid(1), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        case 4: wave_rows_kernel<2, 2, 5><<<grid(2), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        case 5: wave_rows_kernel<2, 3, 4><<<grid(3), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        case 6: wave_rows_kernel<2, 4, 3><<<grid(4), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        default: wave_rows_kernel<2, 2, 4><<<grid(2), 256, 0, st>>>(MOF_WAVE_ARGS(w.cw, w.pw, wave)); break;
        }
        MOF_LAUNCH_CHECK("wave_rows_kernel<2,*,*>");
    }
#undef MOF_WAVE_ARGS
    return 0;
}

}  // namespace

extern "C" int mof_wave_get_variant(void) { return wave_variant(); }

extern "C" int mof_wave_set_variant(int variant) {
    MOF_REQUIRE(variant >= 0 && variant < kWaveVariants, "0 .. 6");
    g_wave_variant = variant;
    return 0;
}

extern "C" int64_t mof_wave_work_doubles(const mof_mesh_dev* mesh, int64_t n_rows, int want_grad, int want_wave) {
    if (!mesh || n_rows < 0) return -1;
    return wave_layout(mesh, n_rows, want_grad != 0, want_wave != 0, nullptr).total;
}

extern "C" int mof_wave_speed(const mof_mesh_dev* mesh, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                              int64_t T_trial, const double* I, int64_t ld, double dt, int phase_mode, double* grad_point,
                              double* wave, double* work, void* stream) {
    MOF_REQUIRE(mesh && I && work && n_rows >= 0 && ld >= mesh->n_vertices && dt != 0.0, "bad arguments");
    MOF_REQUIRE(grad_point || wave, "nothing to compute");
    MOF_REQUIRE((reinterpret_cast<uintptr_t>(work) & 15) == 0, "work must be 16-byte aligned");
    MOF_REQUIRE(out0 >= 0 && n_out >= 0 && out0 + n_out <= n_rows, "output rows outside the rows passed in");
    MOF_REQUIRE(t_first >= 0 && t_first + n_rows <= T_trial, "rows outside the trial");
    MOF_REQUIRE(phase_mode || !wave || T_trial >= 3, "np.gradient(edge_order=2) needs at least 3 frames");
    if (n_out == 0) return 0;
    // halo the time derivative needs around the output rows: one row, two at the trial's ends in amplitude mode
    if (wave) {
        const int64_t a = t_first + out0, b = a + n_out - 1;               // first / last output frame of the trial
        const int64_t need_lo = a == 0 ? 0 : a - 1, need_hi_plain = b == T_trial - 1 ? b : b + 1;
        int64_t lo = need_lo, hi = need_hi_plain;
        if (!phase_mode && a == 0) hi = hi > 2 ? hi : 2;
        if (!phase_mode && b == T_trial - 1) lo = lo < T_trial - 3 ? lo : T_trial - 3;
        if (phase_mode && T_trial == 1) { lo = 0; hi = 0; }
        MOF_REQUIRE(t_first <= lo && hi <= t_first + n_rows - 1, "the rows passed in lack the time-derivative halo of the output rows");
    }
    const int64_t N = mesh->n_vertices;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    MOF_REQUIRE(G <= 65535, "at most 65535 x 32 rows per call");
    cudaStream_t st = mof_stream(stream);
    const wave_work w = wave_layout(mesh, n_rows, grad_point != nullptr, wave != nullptr, work);
    wave_coef_kernel<<<mof_cdiv(N, 128), 128, 0, st>>>(*mesh, w.cw, w.cg, w.pcol, w.pw, w.pg);
    MOF_LAUNCH_CHECK("wave_coef_kernel");
    wave_pack_kernel<<<dim3(mof_cdiv(N, 32), (unsigned)G), dim3(32, 8), 0, st>>>(N, n_rows, mesh->perm, I, ld, w.It, w.Ih);
    MOF_LAUNCH_CHECK("wave_pack_kernel");
    return wave_rows_launch(mesh, w, n_rows, out0, n_out, t_first, T_trial, dt, phase_mode, grad_point, wave, st);
}

// Roofline hook (bench.py): the row kernel(s) alone on a work buffer that a mof_wave_speed call with the same mesh,
// n_rows and outputs has already filled (packed signal + coefficient rows).
extern "C" int mof_wave_stencil(const mof_mesh_dev* mesh, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                                int64_t T_trial, double dt, int phase_mode, double* grad_point, double* wave, double* work,
                                void* stream) {
    MOF_REQUIRE(mesh && work && n_rows > 0 && (grad_point || wave) && dt != 0.0, "bad arguments");
    MOF_REQUIRE((reinterpret_cast<uintptr_t>(work) & 15) == 0, "work must be 16-byte aligned");
    MOF_REQUIRE(out0 >= 0 && n_out >= 0 && out0 + n_out <= n_rows, "output rows outside the rows passed in");
    MOF_REQUIRE((n_rows + MOF_W - 1) / MOF_W <= 65535, "at most 65535 x 32 rows per call");
    const wave_work w = wave_layout(mesh, n_rows, grad_point != nullptr, wave != nullptr, work);
    return wave_rows_launch(mesh, w, n_rows, out0, n_out, t_first, T_trial, dt, phase_mode, grad_point, wave, mof_stream(stream));
}

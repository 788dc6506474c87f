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
// wave_pack_kernel).  wave_rows_kernel: a CTA owns 32 consecutive vertices of one 32-frame group; it stages the
// column indices and coefficients of its 32 block rows in shared memory with one coalesced pass, then every warp walks
// four vertices with lane = frame -- each ring value is one 256-byte line that 32 frames share, the time neighbours are
// the adjacent lanes -- and the tile is transposed through shared memory so that the result is written straight into
// the caller's (T,N[,3]) array as 256-byte (768-byte) row pieces: no frame-minor result, no transpose kernel behind it.
// Algorithmic HBM bytes: 8 N read + 8 N written per frame (the pack moves another 16 N); the ring re-reads are served
// by L1 (inside the CTA's tile) and L2 (a group's It is N x 256 bytes = 42 MB at 164k vertices).
// The wave operator keeps the mesh in REFERENCE vertex order (S5_compute_wave_v._operator: reorder = 0), which
// makes both transposes fully coalesced; the kernels honour mesh->perm all the same.
// A call may cover a SHARD of a trial: rows outside [out0, out0 + n_out) are halo for the time derivative, and the
// one-sided end formulas apply at the trial's ends only (t_first, T_trial).
#include <stdlib.h>

#include "mof_common.cuh"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kTileVerts = 32;        // vertices per CTA of wave_rows_kernel
constexpr int kSlots = 8;            // block-row entries per vertex staged in shared memory (valence <= 7; longer rows continue in global)

// (rows, N) row-major, reference vertex order -> It[g][v][32], internal order (rows padded with 0)
__global__ void __launch_bounds__(256) wave_pack_kernel(int64_t N, int64_t n_rows, const int32_t* __restrict__ perm,
                                                        const double* __restrict__ I, int64_t ld, double* __restrict__ It) {
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
    __syncthreads();
    for (int vv = ty; vv < 32; vv += 8) {
        const int64_t w = v0 + vv;
        if (w < N) It[mof_ix_sca(N, g, w) + tx] = s[tx][vv];
    }
}

// Coefficient rows aligned with the block pattern: cg[j][3] and / or cw[j][2] for block j = (row v, column u).
__global__ void __launch_bounds__(128) wave_coef_kernel(mof_mesh_dev M, double* __restrict__ cw, double* __restrict__ cg) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < M.n_vertices) mof_wave_coef_row_body(M, v, cw, cg);
}

// One sparse row product per (vertex, frame).  C = 2: coefficients (alpha, beta) -> wave speed (T,N);
// C = 3: coefficients of grad_point -> (T,N,3).  GP = 32-frame groups a CTA takes in one pass (blockIdx.y counts
// passes): the column indices and coefficients read from shared memory serve GP groups, and a warp has GP x 8 ring
// lines in flight per vertex.
// The tile's block rows are staged as kSlots padded slots per vertex (thread = (row, slot), one coalesced pass): a slot
// past the end of a row holds the vertex itself with zero coefficients, so the eight products of a vertex are straight-
// line code with no selects; the entries of a row beyond kSlots (valence > 7) are read from global memory behind them,
// in the same ascending column order.
template <int C, int GP>
__global__ void __launch_bounds__(256, GP == 1 ? 4 : 3) wave_rows_kernel(
    int64_t N, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
    const double* __restrict__ coef, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first, int64_t T_trial,
    const double* __restrict__ It, double inv_dt, int phase_mode, double* __restrict__ out) {
    constexpr int OC = C == 3 ? 3 : 1;                 // doubles written per (vertex, frame)
    constexpr int LD = kTileVerts * OC + 1;            // odd row length: both sides of the transpose are conflict-free
    constexpr int CP = C == 3 ? 4 : 2;                 // doubles per staged slot: one (two) 16-byte shared loads
    __shared__ double s_out[GP * 32 * LD];
    __shared__ __align__(16) double s_coef[kTileVerts * kSlots * CP];
    __shared__ __align__(16) int32_t s_col[kTileVerts * kSlots];
    __shared__ int32_t s_j0[kTileVerts], s_cnt[kTileVerts];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t v0 = (int64_t)blockIdx.x * kTileVerts;
    const int64_t g0 = (int64_t)blockIdx.y * GP;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    {   // ---- stage the tile's block rows: thread = (row, slot)
        const int row = tid >> 3, slot = tid & (kSlots - 1);
        const int64_t v = v0 + row;
        int32_t j0 = 0, cnt = 0;
        if (v < N) { j0 = rowptr[v]; cnt = rowptr[v + 1] - j0; }
        if (slot == 0) { s_j0[row] = j0; s_cnt[row] = cnt; }
        int32_t u = v < N ? (int32_t)v : 0;
        double w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = 0.0;
        if (slot < cnt) {
            u = col[j0 + slot];
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] = coef[(size_t)(j0 + slot) * C + c];
        }
        s_col[tid] = u;
#pragma unroll
        for (int c = 0; c < C; ++c) s_coef[tid * CP + c] = w[c];
    }
    __syncthreads();

    // Offsets, relative to a vertex's first line It[0][v][0], of this lane's frame in group g0 + gp and of its neighbours
    // in time (the previous / next frame of lane 0 / 31 lives in the neighbouring group's line).  A pass that reaches past
    // the last group computes its first group twice and writes it once.
    const int64_t group_stride = N * MOF_W;
    int64_t o_cur[GP], o_prev[GP], o_next[GP];
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) {
        const int64_t g = g0 + gp < G ? g0 + gp : g0;
        const int64_t r = g * 32 + lane;
        auto off = [&](int64_t row) { return (row >> 5) * group_stride + (row & 31); };
        o_cur[gp] = off(r);
        o_prev[gp] = off(r > 0 ? r - 1 : r);
        o_next[gp] = off(r + 1 < G * 32 ? r + 1 : r);
    }
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int vv = warp * 4 + i;
        const int64_t v = v0 + vv;
        if (v >= N) break;                                                  // warp-uniform
        const int4 ca = *reinterpret_cast<const int4*>(s_col + vv * kSlots), cb = *reinterpret_cast<const int4*>(s_col + vv * kSlots + 4);
        const int32_t us[kSlots] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        double val[GP][kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const double* line = It + (size_t)us[k] * MOF_W;
#pragma unroll
            for (int gp = 0; gp < GP; ++gp) val[gp][k] = line[o_cur[gp]];
        }
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
                const double* line = It + (size_t)col[j0 + k] * MOF_W;
#pragma unroll
                for (int gp = 0; gp < GP; ++gp) {
                    const double x = line[o_cur[gp]];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[gp][c] = fma(coef[(size_t)(j0 + k) * C + c], x, acc[gp][c]);
                }
            }
        }
        const double* Iv = It + (size_t)v * MOF_W;
#pragma unroll
        for (int gp = 0; gp < GP; ++gp) {
            double* so = s_out + gp * 32 * LD + lane * LD + vv * OC;
            if (C == 3) {
#pragma unroll
                for (int c = 0; c < OC; ++c) so[c] = acc[gp][c];
            } else {
                const int64_t r = (g0 + gp) * 32 + lane;                    // row of the call
                const int64_t t = t_first + r;                              // frame of the trial
                const bool first = t == 0, last_t = t == T_trial - 1;
                const double cur = Iv[o_cur[gp]], prev = Iv[o_prev[gp]], next = Iv[o_next[gp]];
                double far2 = 0.0;                                          // np.gradient's one-sided ends reach two frames in
                if (!phase_mode && r < n_rows && (first || (last_t && r >= 2))) {
                    const int64_t row = first ? r + 2 : r - 2;
                    far2 = Iv[(row >> 5) * group_stride + (row & 31)];
                }
                so[0] = mof_wave_speed_body(mof_wave_td_body(phase_mode, first, last_t, T_trial, cur, prev, next, far2, inv_dt),
                                            acc[gp][0], acc[gp][1 % C]);
            }
        }
    }
    __syncthreads();
    // ---- transposed write: thread column tx walks the tile's doubles of one frame row, 8 rows per pass.  Streaming
    // stores: the result is not read again, and a group's It lines (N x 256 bytes) should stay in L2 for the ring
    // re-reads of the tiles that follow.
    const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
    for (int gp = 0; gp < GP; ++gp) {
        if (g0 + gp >= G) break;
        for (int fr = ty; fr < 32; fr += 8) {
            const int64_t k = (g0 + gp) * 32 + fr - out0;
            if (k < 0 || k >= n_out) continue;
#pragma unroll
            for (int c = 0; c < OC; ++c) {
                const int idx = tx + 32 * c;                                // position inside the tile's row piece
                const int vv = idx / OC;
                if (v0 + vv < N) __stcs(out + ((size_t)k * N + perm[v0 + vv]) * OC + (idx - vv * OC), s_out[gp * 32 * LD + fr * LD + idx]);
            }
        }
    }
}

struct wave_work {
    double *It, *cw, *cg;
    int64_t total;
};

// work = the packed signal It[G][N][32], then cw[nb][2] (wave speed asked for), then cg[nb][3] (grad_point asked for)
wave_work wave_layout(const mof_mesh_dev* mesh, int64_t n_rows, bool want_grad, bool want_wave, double* work) {
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    const int64_t off_cw = G * mesh->n_vertices * MOF_W;
    const int64_t off_cg = off_cw + (want_wave ? 2 * mesh->n_blocks : 0);
    wave_work w;
    w.total = off_cg + (want_grad ? 3 * mesh->n_blocks : 0);
    w.It = work;
    w.cw = work && want_wave ? work + off_cw : nullptr;
    w.cg = work && want_grad ? work + off_cg : nullptr;
    return w;
}

// groups per pass of the wave-speed kernel: 2 by default, MOF_WAVE_GROUPS=1 or mof_wave_set_groups_per_pass(1) selects
// the one-group variant (kept for the comparison in profiles/)
int g_wave_groups = 0;
int wave_groups_per_pass() {
    if (g_wave_groups == 0) {
        const char* e = getenv("MOF_WAVE_GROUPS");
        g_wave_groups = e && e[0] == '1' ? 1 : 2;
    }
    return g_wave_groups;
}

int wave_rows_launch(const mof_mesh_dev* mesh, const wave_work& w, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                     int64_t T_trial, double dt, int phase_mode, double* grad_point, double* wave, cudaStream_t st) {
    const int64_t N = mesh->n_vertices;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    if (grad_point) {
        wave_rows_kernel<3, 1><<<dim3(mof_cdiv(N, kTileVerts), (unsigned)G), 256, 0, st>>>(
            N, mesh->rowptr, mesh->col, mesh->perm, w.cg, n_rows, out0, n_out, t_first, T_trial, w.It, 1.0 / dt, phase_mode, grad_point);
        MOF_LAUNCH_CHECK("wave_rows_kernel<3,1>");
    }
    if (wave) {
        if (wave_groups_per_pass() == 2)
            wave_rows_kernel<2, 2><<<dim3(mof_cdiv(N, kTileVerts), (unsigned)((G + 1) / 2)), 256, 0, st>>>(
                N, mesh->rowptr, mesh->col, mesh->perm, w.cw, n_rows, out0, n_out, t_first, T_trial, w.It, 1.0 / dt, phase_mode, wave);
        else
            wave_rows_kernel<2, 1><<<dim3(mof_cdiv(N, kTileVerts), (unsigned)G), 256, 0, st>>>(
                N, mesh->rowptr, mesh->col, mesh->perm, w.cw, n_rows, out0, n_out, t_first, T_trial, w.It, 1.0 / dt, phase_mode, wave);
        MOF_LAUNCH_CHECK("wave_rows_kernel<2,*>");
    }
    return 0;
}

}  // namespace

extern "C" int mof_wave_set_groups_per_pass(int groups) {
    MOF_REQUIRE(groups == 1 || groups == 2, "1 or 2");
    g_wave_groups = groups;
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
    wave_coef_kernel<<<mof_cdiv(N, 128), 128, 0, st>>>(*mesh, w.cw, w.cg);
    MOF_LAUNCH_CHECK("wave_coef_kernel");
    wave_pack_kernel<<<dim3(mof_cdiv(N, 32), (unsigned)G), dim3(32, 8), 0, st>>>(N, n_rows, mesh->perm, I, ld, w.It);
    MOF_LAUNCH_CHECK("wave_pack_kernel");
    return wave_rows_launch(mesh, w, n_rows, out0, n_out, t_first, T_trial, dt, phase_mode, grad_point, wave, st);
}

// Roofline hook (bench.py): the row kernel(s) alone on a work buffer that a mof_wave_speed call with the same mesh,
// n_rows and outputs has already filled (packed signal + coefficient rows).
extern "C" int mof_wave_stencil(const mof_mesh_dev* mesh, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first,
                                int64_t T_trial, double dt, int phase_mode, double* grad_point, double* wave, double* work,
                                void* stream) {
    MOF_REQUIRE(mesh && work && n_rows > 0 && (grad_point || wave) && dt != 0.0, "bad arguments");
    MOF_REQUIRE(out0 >= 0 && n_out >= 0 && out0 + n_out <= n_rows, "output rows outside the rows passed in");
    MOF_REQUIRE((n_rows + MOF_W - 1) / MOF_W <= 65535, "at most 65535 x 32 rows per call");
    const wave_work w = wave_layout(mesh, n_rows, grad_point != nullptr, wave != nullptr, work);
    return wave_rows_launch(mesh, w, n_rows, out0, n_out, t_first, T_trial, dt, phase_mode, grad_point, wave, mof_stream(stream));
}

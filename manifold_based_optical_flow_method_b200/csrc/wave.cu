// K6 -- wave speed from phase / amplitude maps (S5_compute_wave_v.py): per vertex and frame
//   grad_point = area-weighted mean of the per-face surface gradients of the incident faces
//                (compute_grad_M_I, S5:136-171, same hat-function "gradients" as cof:238-255),
//   its tangent-plane projection expressed in the basis (e1, e2) and the norm of that
//   (project_vector_to_plane S5:173-180, express_vector_on_basis S5:182-191, S5:96-117),
//   the time derivative: wrapped differences for phases (compute_temporal_gradient_phase,
//   S5:60-77, angle_subtract S5:224-233) or np.gradient(edge_order=2) for amplitudes (S5:24),
//   wave_velocity = time derivative / norm (S5:121 / S5:56).
//
// Layout.  Round 1 ran one thread per (vertex, frame) on the (T,N) arrays as they are: every value of
// the 1-ring was an 8-byte load through two index indirections and the result an 8-byte scattered
// store.  Now the signal is first transposed into the frame-minor layout of the solver
// (It[group][internal vertex][32 frames], wave_pack_kernel), the stencil runs with one WARP per
// vertex and lane = frame -- every value of the 1-ring is one 256-byte line that 32 frames share,
// the face geometry is a broadcast, the time neighbours are the adjacent lanes -- and the result is
// transposed back (wave_unpack_kernel).  Algorithmic HBM bytes of the stencil: 8 N read + 8 N written
// per frame; the 1-ring re-reads are served by L1/L2 (internal numbering = breadth-first).
// A call may cover a SHARD of a trial: rows outside [out0, out0 + n_out) are halo for the time
// derivative, and the one-sided end formulas apply at the trial's ends only (t_first, T_trial).
#include "mof_common.cuh"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double angle_subtract(double a, double b) {
    // np.mod(f1 - f2 + pi, 2 pi) - pi : result in [-pi, pi)            (S5:230)
    const double kPi = 3.141592653589793, kTwoPi = 2.0 * 3.141592653589793;
    double d = a - b + kPi;
    double m = fmod(d, kTwoPi);
    if (m != 0.0 && m < 0.0) m += kTwoPi;      // numpy's mod takes the sign of the divisor
    return m - kPi;
}

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
        s[fr][tx] = (k < n_rows && v < N) ? I[k * ld + o] : 0.0;
    }
    __syncthreads();
    for (int vv = ty; vv < 32; vv += 8) {
        const int64_t w = v0 + vv;
        if (w < N) It[mof_ix_sca(N, g, w) + tx] = s[tx][vv];
    }
}

// frame-minor result [g][v][C][32] -> out (n_out, N, C) reference order, rows out0 .. out0+n_out-1 of the call
template <int C>
__global__ void __launch_bounds__(256) wave_unpack_kernel(int64_t N, int64_t out0, int64_t n_out, const int32_t* __restrict__ perm,
                                                          const double* __restrict__ Wt, double* __restrict__ out) {
    __shared__ double s[C][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t g = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    for (int vv = ty; vv < 32; vv += 8) {
        const int64_t v = v0 + vv;
        if (v < N)
#pragma unroll
            for (int c = 0; c < C; ++c) s[c][vv][tx] = Wt[((size_t)(g * N + v) * C + c) * MOF_W + tx];
    }
    __syncthreads();
    const int64_t v = v0 + tx;
    if (v >= N) return;
    const int64_t o = perm[v];
    for (int fr = ty; fr < 32; fr += 8) {
        const int64_t k = g * 32 + fr - out0;
        if (k >= 0 && k < n_out)
#pragma unroll
            for (int c = 0; c < C; ++c) out[((size_t)k * N + o) * C + c] = s[c][tx][fr];
    }
}

// Ring record of a vertex: its incident faces (ascending, the order of S5:161-166) as {face, v0, v1, v2},
// kRingFaces slots of four int32 (face = -1: empty; a vertex with more faces keeps the rest in the contributor
// list and sets slot kRingFaces-1's face to -2).  One coalesced 128-byte load gives a warp every index it
// needs, so the gathers of a vertex are all independent instead of a centry -> tri -> value chain per face.
constexpr int kRingFaces = 8;

// ... and the vertex constants the per-frame arithmetic would otherwise recompute in every lane: 1 / area sum of the
// ring (S5:169), the plane normal e1 x e2 and 1 / |n|^2 (S5:177-179), 1 / |e1|^2, 1 / |e2|^2 (S5:189-190).  fp64
// divisions cost ~25 instructions each; as reciprocals computed once per vertex they are a multiplication per frame
// (one rounding more than the reference's division, 1 ulp, inside the 1e-12 parity tolerance).
constexpr int kVertexConsts = 8;      // inv_area_sum, n[3], inv_nn, inv_e1e1, inv_e2e2, pad

__global__ void wave_ring_kernel(mof_mesh_dev M, int32_t* __restrict__ ring, double* __restrict__ vc) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M.n_vertices) return;
    const int32_t bd = M.diag[v];
    const int32_t q0 = M.cptr[bd], q1 = M.cptr[bd + 1];
    {
        double asum = 0.0;
        for (int32_t q = q0; q < q1; ++q) asum += M.areas[M.centry[q] >> 4];      // same order as the reference's sum
        const double* e1 = M.e + 6 * v;
        const double* e2 = e1 + 3;
        const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
        double* c = vc + (size_t)v * kVertexConsts;
        c[0] = 1.0 / asum;
        c[1] = nx; c[2] = ny; c[3] = nz;
        c[4] = 1.0 / (nx * nx + ny * ny + nz * nz);
        c[5] = 1.0 / (e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        c[6] = 1.0 / (e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
        c[7] = 0.0;
    }
    int32_t* r = ring + (size_t)v * kRingFaces * 4;
    for (int k = 0; k < kRingFaces; ++k) {
        const int32_t q = q0 + k;
        const bool on = q < q1;
        const int32_t f = on ? M.centry[q] >> 4 : -1;
        r[4 * k] = (k == kRingFaces - 1 && q1 - q0 > kRingFaces) ? -2 : f;
        r[4 * k + 1] = on ? M.tri[3 * f] : 0;
        r[4 * k + 2] = on ? M.tri[3 * f + 1] : 0;
        r[4 * k + 3] = on ? M.tri[3 * f + 2] : 0;
    }
}

// One warp per internal vertex, lane = frame (row 32 g + lane of the call).
__global__ void __launch_bounds__(256) wave_stencil_kernel(mof_mesh_dev M, const int32_t* __restrict__ ring,
                                                           const double* __restrict__ vc, int64_t n_rows,
                                                           int64_t t_first, int64_t T_trial, const double* __restrict__ It,
                                                           double dt, int phase_mode, double* __restrict__ Gt,
                                                           double* __restrict__ Wt) {
    const int64_t N = M.n_vertices;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v = (int64_t)blockIdx.x * 8 + warp;
    const int64_t g = blockIdx.y;
    if (v >= N) return;
    const double* It_l = It + mof_ix_sca(N, g, 0) + lane;
    // area-weighted mean of the face gradients, faces ascending (S5:161-169)
    double gx = 0.0, gy = 0.0, gz = 0.0;
    const int32_t rec = ring[(size_t)v * kRingFaces * 4 + lane];
    const double* kc = vc + (size_t)v * kVertexConsts;
    auto add_face = [&](int64_t f, int64_t t0, int64_t t1, int64_t t2) {
        const double* gw = M.grad_w + 9 * f;
        const double I0 = It_l[(size_t)t0 * MOF_W], I1 = It_l[(size_t)t1 * MOF_W], I2 = It_l[(size_t)t2 * MOF_W];
        const double A = M.areas[f];
        gx += (I0 * gw[0] + I1 * gw[3] + I2 * gw[6]) * A;                   // S5:154-158,165
        gy += (I0 * gw[1] + I1 * gw[4] + I2 * gw[7]) * A;
        gz += (I0 * gw[2] + I1 * gw[5] + I2 * gw[8]) * A;
    };
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < kRingFaces; ++k) {
        const int32_t f = __shfl_sync(kFullMask, rec, 4 * k);
        const int32_t t0 = __shfl_sync(kFullMask, rec, 4 * k + 1), t1 = __shfl_sync(kFullMask, rec, 4 * k + 2),
                      t2 = __shfl_sync(kFullMask, rec, 4 * k + 3);
        if (f >= 0) add_face(f, t0, t1, t2);
        else if (f == -2) overflow = true;
    }
    if (overflow) {                                  // more than kRingFaces faces: the rest straight from the contributor list
        const int32_t bd = M.diag[v];
        for (int32_t q = M.cptr[bd] + kRingFaces - 1; q < M.cptr[bd + 1]; ++q) {
            const int64_t f = M.centry[q] >> 4;
            add_face(f, M.tri[3 * f], M.tri[3 * f + 1], M.tri[3 * f + 2]);
        }
    }
    gx *= kc[0]; gy *= kc[0]; gz *= kc[0];                                     // S5:169
    if (Gt) {
        double* gp = Gt + ((size_t)(g * N + v) * 3) * MOF_W + lane;
        gp[0] = gx; gp[MOF_W] = gy; gp[2 * MOF_W] = gz;
    }
    if (!Wt) return;
    const double* e1 = M.e + 6 * v;
    const double* e2 = e1 + 3;
    // project_vector_to_plane (S5:173-180)
    const double nx = kc[1], ny = kc[2], nz = kc[3];
    const double s = (gx * nx + gy * ny + gz * nz) * kc[4];
    const double px = gx - s * nx, py = gy - s * ny, pz = gz - s * nz;
    // express_vector_on_basis (S5:182-191) and its norm (S5:117)
    const double al = (px * e1[0] + py * e1[1] + pz * e1[2]) * kc[5];
    const double be = (px * e2[0] + py * e2[1] + pz * e2[2]) * kc[6];
    const double dis = sqrt(al * al + be * be);
    // time derivative: the neighbours in time are the adjacent lanes; the lanes at the ends of the group fetch
    // theirs from the next / previous group
    const int64_t r = g * 32 + lane;                                        // row of the call
    const int64_t t = t_first + r;                                          // frame of the trial
    const double* Iv = It + (size_t)v * MOF_W;                              // It[gg][v][ll] = Iv[gg * N * 32 + ll]
    auto at = [&](int64_t row) { return Iv[(size_t)(row >> 5) * N * MOF_W + (row & 31)]; };
    const double c = It_l[(size_t)v * MOF_W];
    double prev = __shfl_up_sync(kFullMask, c, 1), next = __shfl_down_sync(kFullMask, c, 1);
    if (lane == 0 && r > 0) prev = at(r - 1);
    if (lane == 31 && r + 1 < n_rows) next = at(r + 1);
    double td = 0.0;
    if (r < n_rows) {
        if (phase_mode) {                                                   // S5:60-77
            if (T_trial == 1) td = 0.0;
            else if (t == 0) td = angle_subtract(next, c) / dt;
            else if (t == T_trial - 1) td = angle_subtract(c, prev) / dt;
            else td = angle_subtract(next, prev) / (2 * dt);
        } else {                                                            // np.gradient(axis=0, edge_order=2) / dt, S5:24
            if (t == 0) td = (-1.5 * c + 2.0 * next - 0.5 * at(r + 2)) / dt;
            else if (t == T_trial - 1) td = (1.5 * c - 2.0 * prev + 0.5 * at(r - 2)) / dt;
            else td = ((next - prev) / 2.0) / dt;
        }
    }
    Wt[mof_ix_sca(N, g, v) + lane] = td / dis;                              // S5:121
}

}  // namespace

extern "C" int64_t mof_wave_work_doubles(int64_t n_vertices, int64_t n_rows, int want_grad, int want_wave) {
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    // signal + results in the frame-minor layout, then the ring records (kRingFaces x 4 int32 per vertex)
    return G * n_vertices * MOF_W * (1 + (want_grad ? 3 : 0) + (want_wave ? 1 : 0)) + n_vertices * (kRingFaces * 4 / 2 + kVertexConsts);
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
    double* It = work;
    double* Gt = grad_point ? It + (size_t)G * N * MOF_W : nullptr;
    double* Wt = wave ? It + (size_t)G * N * MOF_W * (grad_point ? 4 : 1) : nullptr;
    dim3 tgrid(mof_cdiv(N, 32), (unsigned)G), tblock(32, 8);
    int32_t* ring = reinterpret_cast<int32_t*>(It + (size_t)G * N * MOF_W * (1 + (grad_point ? 3 : 0) + (wave ? 1 : 0)));
    double* vc = reinterpret_cast<double*>(ring + (size_t)N * kRingFaces * 4);
    wave_ring_kernel<<<mof_cdiv(N, 256), 256, 0, st>>>(*mesh, ring, vc);
    MOF_LAUNCH_CHECK("wave_ring_kernel");
    wave_pack_kernel<<<tgrid, tblock, 0, st>>>(N, n_rows, mesh->perm, I, ld, It);
    MOF_LAUNCH_CHECK("wave_pack_kernel");
    wave_stencil_kernel<<<dim3(mof_cdiv(N, 8), (unsigned)G), 256, 0, st>>>(*mesh, ring, vc, n_rows, t_first, T_trial, It, dt, phase_mode, Gt, Wt);
    MOF_LAUNCH_CHECK("wave_stencil_kernel");
    if (grad_point) {
        wave_unpack_kernel<3><<<tgrid, tblock, 0, st>>>(N, out0, n_out, mesh->perm, Gt, grad_point);
        MOF_LAUNCH_CHECK("wave_unpack_kernel<3>");
    }
    if (wave) {
        wave_unpack_kernel<1><<<tgrid, tblock, 0, st>>>(N, out0, n_out, mesh->perm, Wt, wave);
        MOF_LAUNCH_CHECK("wave_unpack_kernel<1>");
    }
    return 0;
}

// Roofline hook (bench.py): the stencil alone on a work buffer that mof_wave_speed has already packed.
extern "C" int mof_wave_stencil(const mof_mesh_dev* mesh, int64_t n_rows, int64_t t_first, int64_t T_trial, double dt,
                                int phase_mode, int want_grad, int want_wave, double* work, void* stream) {
    MOF_REQUIRE(mesh && work && n_rows > 0 && (want_grad || want_wave), "bad arguments");
    const int64_t N = mesh->n_vertices;
    const int64_t G = (n_rows + MOF_W - 1) / MOF_W;
    MOF_REQUIRE(G <= 65535, "at most 65535 x 32 rows per call");
    double* It = work;
    double* Gt = want_grad ? It + (size_t)G * N * MOF_W : nullptr;
    double* Wt = want_wave ? It + (size_t)G * N * MOF_W * (want_grad ? 4 : 1) : nullptr;
    const int32_t* ring = reinterpret_cast<const int32_t*>(It + (size_t)G * N * MOF_W * (1 + (want_grad ? 3 : 0) + (want_wave ? 1 : 0)));
    const double* vc = reinterpret_cast<const double*>(ring + (size_t)N * kRingFaces * 4);
    wave_stencil_kernel<<<dim3(mof_cdiv(N, 8), (unsigned)G), 256, 0, mof_stream(stream)>>>(*mesh, ring, vc, n_rows, t_first, T_trial, It, dt,
                                                                                          phase_mode, Gt, Wt);
    MOF_LAUNCH_CHECK("wave_stencil_kernel");
    return 0;
}

// K8 -- radial-basis interpolation of electrode signals onto the mesh vertices, the producer of the
//       hot path's (T, N) input ("next" row 4 of SURVEY.md 8f): `interpolation`,
//       S2_interpolate.py:22-53 and S2_interpolate_phases.py:22-56, i.e. scipy.interpolate.Rbf with
//       its defaults (multiquadric phi(r) = sqrt((r/eps)^2 + 1), smooth 0, Euclidean norm), once per
//       frame in the reference: build and solve the same m x m system T times, then a dense
//       (N x m) kernel matrix times the weights.
//
// Here the m x m matrix is built and LU-factorised ONCE (partial pivoting, one CTA -- m is the
// number of electrodes, 10^1..10^3), all frames are solved against it (one thread per frame), and
// the evaluation is one GEMM-shaped kernel out(T, N) = W(T, m) . Phi(m, N) with Phi computed on
// the fly in shared memory (never stored in HBM: at 164k vertices x 128 electrodes it would be
// 168 MB).  fp64 FMA on the CUDA cores: 64 x 64 output tile, 4 x 4 per thread, centres in chunks
// of 32; up to ~390 electrodes a CTA keeps its whole Phi tile resident and walks all frame tiles.  Phase mode interpolates real and imaginary parts and writes
// atan2(im, re) (np.angle, S2_interpolate_phases.py:52).
#include "mof_common.cuh"

namespace {

constexpr int kLuThreads = 1024;

__global__ void rbf_matrix_kernel(int m, const double* __restrict__ c, double inv_eps, double* __restrict__ A) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * m) return;
    const int i = (int)(idx / m), j = (int)(idx % m);
    A[idx] = mof_rbf_phi_body(c + 3 * (size_t)i, c + 3 * (size_t)j, inv_eps);
}

// Right-looking LU with partial pivoting (first largest |entry|, like idamax), row-major A, LAPACK-style
// ipiv (row k was swapped with row piv[k]).  info = k+1 for an exactly zero pivot.
__global__ void __launch_bounds__(kLuThreads) rbf_lu_kernel(int m, double* __restrict__ A, int32_t* __restrict__ piv,
                                                            int32_t* __restrict__ info) {
    __shared__ double s_val[kLuThreads / 32];
    __shared__ int s_idx[kLuThreads / 32];
    __shared__ int s_p;
    const int tid = threadIdx.x;
    if (tid == 0) *info = 0;
    for (int k = 0; k < m; ++k) {
        double best = -1.0;
        int arg = 0x7fffffff;
        for (int i = k + tid; i < m; i += kLuThreads) {
            const double a = fabs(A[(size_t)i * m + k]);
            if (a > best) { best = a; arg = i; }
        }
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = arg; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kLuThreads / 32; ++w)
                if (s_val[w] > best || (s_val[w] == best && s_idx[w] < arg)) { best = s_val[w]; arg = s_idx[w]; }
            if (arg == 0x7fffffff) arg = k;                       // column of NaNs
            if (!(best > 0.0) && *info == 0) *info = k + 1;
            piv[k] = arg;
            s_p = arg;
        }
        __syncthreads();
        const int p = s_p;
        if (p != k)
            for (int j = tid; j < m; j += kLuThreads) {
                const double a = A[(size_t)k * m + j];
                A[(size_t)k * m + j] = A[(size_t)p * m + j];
                A[(size_t)p * m + j] = a;
            }
        __syncthreads();
        const double d = A[(size_t)k * m + k];
        for (int i = k + 1 + tid; i < m; i += kLuThreads) A[(size_t)i * m + k] /= d;
        __syncthreads();
        const int w = m - k - 1;
        for (int64_t q = tid; q < (int64_t)w * w; q += kLuThreads) {
            const int i = k + 1 + (int)(q / w), j = k + 1 + (int)(q % w);
            A[(size_t)i * m + j] = fma(-A[(size_t)i * m + k], A[(size_t)k * m + j], A[(size_t)i * m + j]);
        }
        __syncthreads();
    }
}

// One thread per right-hand side t: W[:, t] = A^-1 data[t, :]; W is (m, n_rhs), t minor.
__global__ void rbf_solve_kernel(int m, int64_t n_rhs, const double* __restrict__ LU, const int32_t* __restrict__ piv,
                                 const double* __restrict__ data, int64_t ld, double* __restrict__ W) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rhs) return;
    for (int j = 0; j < m; ++j) W[(size_t)j * n_rhs + t] = data[(size_t)t * ld + j];
    for (int k = 0; k < m; ++k) {
        const int p = piv[k];
        if (p != k) {
            const double a = W[(size_t)k * n_rhs + t];
            W[(size_t)k * n_rhs + t] = W[(size_t)p * n_rhs + t];
            W[(size_t)p * n_rhs + t] = a;
        }
    }
    for (int i = 1; i < m; ++i) {                                  // L y = P b (unit lower)
        double s = W[(size_t)i * n_rhs + t];
        for (int j = 0; j < i; ++j) s = fma(-LU[(size_t)i * m + j], W[(size_t)j * n_rhs + t], s);
        W[(size_t)i * n_rhs + t] = s;
    }
    for (int i = m - 1; i >= 0; --i) {                             // U x = y
        double s = W[(size_t)i * n_rhs + t];
        for (int j = i + 1; j < m; ++j) s = fma(-LU[(size_t)i * m + j], W[(size_t)j * n_rhs + t], s);
        W[(size_t)i * n_rhs + t] = s / LU[(size_t)i * m + i];
    }
}

constexpr int kTile = 64;      // output tile: 64 frames x 64 vertices per CTA, 4 x 4 per thread
constexpr int kChunk = 32;     // centres per shared-memory stage

template <int PHASE>
__global__ void __launch_bounds__(256) rbf_eval_kernel(int64_t N, int m, int64_t T, int64_t n_rhs,
                                                       const double* __restrict__ vertices, const double* __restrict__ centres,
                                                       double inv_eps, const double* __restrict__ W, double* __restrict__ out,
                                                       int64_t ld) {
    __shared__ __align__(16) double Ps[kChunk][kTile];
    __shared__ __align__(16) double Wr[kChunk][kTile];
    __shared__ __align__(16) double Wi[PHASE ? kChunk : 1][kTile];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t n0 = (int64_t)blockIdx.x * kTile, t0 = (int64_t)blockIdx.y * kTile;
    const int lane64 = tid & 63, row4 = tid >> 6;                  // staging: column lane64, rows row4 + 4q
    const int64_t nv = n0 + lane64;
    double x[3] = {0.0, 0.0, 0.0};
    if (nv < N) { x[0] = vertices[3 * nv]; x[1] = vertices[3 * nv + 1]; x[2] = vertices[3 * nv + 2]; }
    const int64_t tf = t0 + lane64;
    double acc[4][4], aci[PHASE ? 4 : 1][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b] = 0.0; if constexpr (PHASE != 0) aci[a][b] = 0.0; }
    for (int k0 = 0; k0 < m; k0 += kChunk) {
#pragma unroll
        for (int q = 0; q < kChunk / 4; ++q) {
            const int kk = row4 + 4 * q, k = k0 + kk;
            const bool kin = k < m;
            Ps[kk][lane64] = kin ? mof_rbf_phi_body(x, centres + 3 * (size_t)k, inv_eps) : 0.0;
            Wr[kk][lane64] = (kin && tf < T) ? W[(size_t)k * n_rhs + tf] : 0.0;
            if constexpr (PHASE != 0) Wi[kk][lane64] = (kin && tf < T) ? W[(size_t)k * n_rhs + T + tf] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kChunk; ++kk) {
            // a thread's four vertices are (2tx, 2tx+1, 32+2tx, 33+2tx): each 16-lane LDS.128 is contiguous
            const double2 p01 = *reinterpret_cast<const double2*>(&Ps[kk][2 * tx]);
            const double2 p23 = *reinterpret_cast<const double2*>(&Ps[kk][kTile / 2 + 2 * tx]);
            const double p[4] = {p01.x, p01.y, p23.x, p23.y};
            const double2 w01 = *reinterpret_cast<const double2*>(&Wr[kk][4 * ty]);
            const double2 w23 = *reinterpret_cast<const double2*>(&Wr[kk][4 * ty + 2]);
            const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(w[a], p[b], acc[a][b]);
            if constexpr (PHASE != 0) {
                const double2 v01 = *reinterpret_cast<const double2*>(&Wi[kk][4 * ty]);
                const double2 v23 = *reinterpret_cast<const double2*>(&Wi[kk][4 * ty + 2]);
                const double v[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) aci[a][b] = fma(v[a], p[b], aci[a][b]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t t = t0 + 4 * ty + a;
        if (t >= T) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t n = n0 + (b >> 1) * (kTile / 2) + 2 * tx + (b & 1);
            if (n >= N) continue;
            if constexpr (PHASE != 0) out[(size_t)t * ld + n] = atan2(aci[a][b], acc[a][b]);
            else out[(size_t)t * ld + n] = acc[a][b];
        }
    }
}

// Same product with the kernel-matrix tile Phi[m][64] of the CTA's 64 vertices RESIDENT in shared memory:
// it is computed once (two fp64 square roots per entry) and reused by every frame tile, instead of once per
// frame tile as above -- at T = 1000 that removes 15/16 of the square roots, which cost as much as the
// product itself.  Weights stream through a 32 x 64 stage, prefetched into registers one chunk ahead.
// Needs m * 512 B + 16 (32) KB of shared memory: m <= ~390 (~360 in phase mode); larger m use the kernel above.
template <int PHASE>
__global__ void __launch_bounds__(256, 2) rbf_eval_resident_kernel(int64_t N, int m, int64_t T, int64_t n_rhs,
                                                                const double* __restrict__ vertices,
                                                                const double* __restrict__ centres, double inv_eps,
                                                                const double* __restrict__ W, double* __restrict__ out,
                                                                int64_t ld) {
    extern __shared__ __align__(16) double rbf_smem[];
    double* Phi = rbf_smem;                                     // [m][64]
    double* Wr = Phi + (size_t)m * kTile;                       // [32][64]
    double* Wi = Wr + kChunk * kTile;                           // [32][64], phase mode only
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int lane64 = tid & 63, row4 = tid >> 6;
    const int64_t n0 = (int64_t)blockIdx.x * kTile;
    {
        const int64_t nv = n0 + lane64;
        double x[3] = {0.0, 0.0, 0.0};
        if (nv < N) { x[0] = vertices[3 * nv]; x[1] = vertices[3 * nv + 1]; x[2] = vertices[3 * nv + 2]; }
        for (int k = row4; k < m; k += 4) Phi[(size_t)k * kTile + lane64] = mof_rbf_phi_body(x, centres + 3 * (size_t)k, inv_eps);
    }
    __syncthreads();
    for (int64_t t0 = 0; t0 < T; t0 += kTile) {
        const int64_t tf = t0 + lane64;
        double wr[kChunk / 4], wi[PHASE ? kChunk / 4 : 1];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int q = 0; q < kChunk / 4; ++q) {
                const int k = k0 + row4 + 4 * q;
                const bool in = k < m && tf < T;
                wr[q] = in ? W[(size_t)k * n_rhs + tf] : 0.0;
                if constexpr (PHASE != 0) wi[q] = in ? W[(size_t)k * n_rhs + T + tf] : 0.0;
            }
        };
        double acc[4][4], aci[PHASE ? 4 : 1][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { acc[a][b] = 0.0; if constexpr (PHASE != 0) aci[a][b] = 0.0; }
        fetch(0);
        for (int k0 = 0; k0 < m; k0 += kChunk) {
#pragma unroll
            for (int q = 0; q < kChunk / 4; ++q) {
                Wr[(row4 + 4 * q) * kTile + lane64] = wr[q];
                if constexpr (PHASE != 0) Wi[(row4 + 4 * q) * kTile + lane64] = wi[q];
            }
            __syncthreads();
            if (k0 + kChunk < m) fetch(k0 + kChunk);
            const int kmax = min(kChunk, m - k0);
            const double* ph = Phi + (size_t)k0 * kTile;
#pragma unroll 4
            for (int kk = 0; kk < kmax; ++kk) {
                const double2 p01 = *reinterpret_cast<const double2*>(ph + kk * kTile + 2 * tx);
                const double2 p23 = *reinterpret_cast<const double2*>(ph + kk * kTile + kTile / 2 + 2 * tx);
                const double p[4] = {p01.x, p01.y, p23.x, p23.y};
                const double2 w01 = *reinterpret_cast<const double2*>(Wr + kk * kTile + 4 * ty);
                const double2 w23 = *reinterpret_cast<const double2*>(Wr + kk * kTile + 4 * ty + 2);
                const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(w[a], p[b], acc[a][b]);
                if constexpr (PHASE != 0) {
                    const double2 v01 = *reinterpret_cast<const double2*>(Wi + kk * kTile + 4 * ty);
                    const double2 v23 = *reinterpret_cast<const double2*>(Wi + kk * kTile + 4 * ty + 2);
                    const double v[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) aci[a][b] = fma(v[a], p[b], aci[a][b]);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int64_t t = t0 + 4 * ty + a;
            if (t >= T) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t n = n0 + h * (kTile / 2) + 2 * tx;
                double v0, v1;
                if constexpr (PHASE != 0) { v0 = atan2(aci[a][2 * h], acc[a][2 * h]); v1 = atan2(aci[a][2 * h + 1], acc[a][2 * h + 1]); }
                else { v0 = acc[a][2 * h]; v1 = acc[a][2 * h + 1]; }
                double* o = out + (size_t)t * ld + n;
                if (n + 1 < N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
                else { if (n < N) o[0] = v0; if (n + 1 < N) o[1] = v1; }
            }
        }
    }
}

}  // namespace

extern "C" int mof_rbf_fit(int64_t m, const double* centres, double epsilon, int64_t n_rhs, const double* data, int64_t ld,
                           double* lu, int32_t* piv, double* weights, int32_t* info, void* stream) {
    MOF_REQUIRE(m > 0 && m <= 8192 && n_rhs >= 0 && ld >= m, "bad sizes (1 <= n_centres <= 8192, ld >= n_centres)");
    MOF_REQUIRE(centres && lu && piv && info && (n_rhs == 0 || (data && weights)), "null argument");
    MOF_REQUIRE(epsilon > 0.0 && epsilon < INFINITY, "epsilon must be positive and finite");
    cudaStream_t st = mof_stream(stream);
    rbf_matrix_kernel<<<mof_cdiv(m * m, 256), 256, 0, st>>>((int)m, centres, 1.0 / epsilon, lu);
    MOF_LAUNCH_CHECK("rbf_matrix_kernel");
    rbf_lu_kernel<<<1, kLuThreads, 0, st>>>((int)m, lu, piv, info);
    MOF_LAUNCH_CHECK("rbf_lu_kernel");
    if (n_rhs > 0) {
        rbf_solve_kernel<<<mof_cdiv(n_rhs, 64), 64, 0, st>>>((int)m, n_rhs, lu, piv, data, ld, weights);
        MOF_LAUNCH_CHECK("rbf_solve_kernel");
    }
    return 0;
}

extern "C" int mof_rbf_evaluate(int64_t N, int64_t m, int64_t n_frames, const double* vertices, const double* centres,
                                double epsilon, const double* weights, int phase_mode, double* out, int64_t ld, void* stream) {
    MOF_REQUIRE(N > 0 && m > 0 && n_frames >= 0 && ld >= N, "bad sizes");
    if (n_frames == 0) return 0;
    MOF_REQUIRE(vertices && centres && weights && out, "null argument");
    MOF_REQUIRE(epsilon > 0.0 && epsilon < INFINITY, "epsilon must be positive and finite");
    MOF_REQUIRE(mof_cdiv(n_frames, kTile) <= 65535, "at most 4,194,240 frames per call");
    cudaStream_t st = mof_stream(stream);
    const size_t resident = ((size_t)m * kTile + (phase_mode ? 2 : 1) * kChunk * kTile) * sizeof(double);
    if (resident <= 200 * 1024) {
        if (phase_mode) {
            MOF_CUDA_TRY(cudaFuncSetAttribute(rbf_eval_resident_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resident));
            rbf_eval_resident_kernel<1><<<mof_cdiv(N, kTile), 256, resident, st>>>(N, (int)m, n_frames, 2 * n_frames, vertices, centres,
                                                                                1.0 / epsilon, weights, out, ld);
        } else {
            MOF_CUDA_TRY(cudaFuncSetAttribute(rbf_eval_resident_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resident));
            rbf_eval_resident_kernel<0><<<mof_cdiv(N, kTile), 256, resident, st>>>(N, (int)m, n_frames, n_frames, vertices, centres,
                                                                                1.0 / epsilon, weights, out, ld);
        }
        MOF_LAUNCH_CHECK("rbf_eval_resident_kernel");
        return 0;
    }
    dim3 grid(mof_cdiv(N, kTile), mof_cdiv(n_frames, kTile));
    if (phase_mode)
        rbf_eval_kernel<1><<<grid, 256, 0, st>>>(N, (int)m, n_frames, 2 * n_frames, vertices, centres, 1.0 / epsilon, weights, out, ld);
    else
        rbf_eval_kernel<0><<<grid, 256, 0, st>>>(N, (int)m, n_frames, n_frames, vertices, centres, 1.0 / epsilon, weights, out, ld);
    MOF_LAUNCH_CHECK("rbf_eval_kernel");
    return 0;
}

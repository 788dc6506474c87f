// K2/K3 -- batched fp64 preconditioned conjugate gradients, two preconditioners, one driver.
// Replaces scipy.sparse.linalg.spsolve (SuperLU) in worker,
// utils/compute_optical_flow.py:147, for a whole batch of frames at once.
//
// Layout and mapping.  The matrix of every frame shares one block pattern (rowptr/col over
// the renumbered vertex adjacency).  Values and vectors are frame-minor,
// [group][slot][32 frames]: a warp works on block rows, lane = frame, so each value /
// vector access of a warp is one 256-byte line and the column index is read once for 32
// frames.  Vector kernels and the SpMV: a CTA (8 warps) owns a tile of MOF_TILE_ROWS
// consecutive rows, grid = (tiles, groups).  SSOR sweeps: a warp owns one patch (= tile) of
// the launch's colour and walks its rows in sequence, 8 patches per CTA.
//
// Reductions.  Dot products are private to a lane (a frame) while a warp walks its rows;
// the 8 warps of a CTA are combined through shared memory in warp order, the per-tile
// partials are written to `partial`, and the last CTA of a group to finish (ticket
// counter) adds the tiles in index order.  No floating-point atomics: results are
// bit-reproducible and independent of how frames are batched or sharded over GPUs.
// That last CTA also does the scalar step (alpha, beta, convergence test per frame).
//
// omega = 0, block Jacobi (the north-star design), per iteration:
//   pupdate_kernel (p = zs z + beta p), spmv_kernel (ap = A p, p'Ap, alpha),
//   update_kernel<false> (x += alpha p, r -= alpha ap, z = D^-1 r, r'z, r'r, beta, convergence).
//   HBM bytes per frame-iteration: spmv 32 nb + 32 N, update 136 N, pupdate 48 N  (72.1 MB at ico7)
// omega in (0,2), block-multicolour SSOR in Eisenstat's form on the D^-1/2-scaled system
// (identity diagonal blocks, see mof_bodies.h), per iteration:
//   sweep_back_kernel<0> per colour, last colour first (x += alpha p of the previous step and
//                        p = zs r/omega + beta p fused; t = (Dt+U)^-1 p),
//   sweep_fwd_kernel<0> per colour (w = (Dt+L)^-1 (p - ((2-omega)/omega) t); p'(t+w); alpha),
//   update_kernel<true> (r -= alpha (t+w), r'r, beta, convergence).
//   HBM bytes per frame-iteration: sweeps 2 x 16 (nb-N) + 144 N, update 64 N  (65.5 MB at ico7),
//   ~3.1x fewer iterations than block Jacobi.
// The same SSOR on the level-scheduled natural ordering (mesh built with reorder = 3; the default of
// the Python layer), ~2.6x fewer iterations again (143 vs 372 at ico7): ONE persistent cooperative kernel
// (level_iter_kernel) runs check_every whole iterations per launch -- row-level dataflow inside the sweeps,
// bulk-async prefetch, p'Ap accumulated exactly in fixed point, the r update as a further phase; fallback
// (bit-identical): level_back_kernel<0> / level_fwd_kernel<0> per dependency level, one warp per row,
// level_alpha_kernel, update_kernel<true>, ~2050 small launches per iteration replayed as a CUDA graph.
// A frame that meets its threshold is frozen with (alpha, beta, zs) = (0, 1, 0) and can resume
// exactly; a group whose frames are all frozen makes its CTAs return at once.
#include <cooperative_groups.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "mof_common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kRowsPerWarp = MOF_TILE_ROWS / kWarps;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kPrefetchRows = 4;      // SSOR sweeps: matrix values are prefetched to L2 this many rows ahead

__device__ __forceinline__ double* scal_ptr(double* scal, int64_t g, int which) {
    return scal + ((size_t)g * MOF_S_COUNT + which) * MOF_W;
}
inline double* scal_host_ptr(double* scal, int64_t g, int which) { return scal + ((size_t)g * MOF_S_COUNT + which) * MOF_W; }
__device__ __forceinline__ int32_t* state_ptr(int32_t* st, int64_t g, int which) {
    return st + ((size_t)g * MOF_I_COUNT + which) * MOF_W;
}
__device__ __forceinline__ int32_t* group_done_ptr(int32_t* st, int G) { return st + (size_t)G * MOF_I_COUNT * MOF_W; }
__device__ __forceinline__ int32_t* ticket_ptr(int32_t* st, int G) { return group_done_ptr(st, G) + G; }
__device__ __forceinline__ int32_t* groups_active_ptr(int32_t* st, int G) { return group_done_ptr(st, G) + 2 * G; }
__device__ __forceinline__ int32_t* lanes_active_ptr(int32_t* st, int G) { return group_done_ptr(st, G) + 2 * G + 1; }

// Sum the per-tile partials of group g in tile-index order (8 warps take strided subsets, then
// are combined in warp order).  Called by every thread of the last CTA; `tot` valid in warp 0.
template <int NV>
__device__ __forceinline__ void reduce_all_tiles(const double* __restrict__ partial_g, int ntiles, double (&tot)[NV],
                                                 double (*red)[NV][MOF_W]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kBatch = 8;                          // loads in flight per warp; the additions keep their order
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double acc = 0.0;
        int t = warp;
        for (; t + (kBatch - 1) * kWarps < ntiles; t += kBatch * kWarps) {
            double v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) v[u] = __ldcg(partial_g + ((size_t)(t + u * kWarps) * 2 + k) * MOF_W + lane);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) acc += v[u];
        }
        for (; t < ntiles; t += kWarps) acc += __ldcg(partial_g + ((size_t)t * 2 + k) * MOF_W + lane);
        red[warp][k][lane] = acc;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = red[0][k][lane];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s += red[w][k][lane];
            tot[k] = s;
        }
    }
}

// Deterministic CTA + cross-tile reduction of NV per-lane values.  Returns true (for every
// thread of the CTA) in the last CTA of group g; there `tot` holds the totals in warp 0.
template <int NV>
__device__ __forceinline__ bool tile_reduce(const double (&val)[NV], double* __restrict__ partial_g, int ntiles,
                                            int tile, int32_t* ticket_g, double (&tot)[NV]) {
    __shared__ double red[kWarps][NV][MOF_W];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) red[warp][k][lane] = val[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = red[0][k][lane];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s += red[w][k][lane];
            partial_g[((size_t)tile * 2 + k) * MOF_W + lane] = s;
        }
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket_g, 1) == ntiles - 1);
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    reduce_all_tiles<NV>(partial_g, ntiles, tot, red);
    if (threadIdx.x == 0) *ticket_g = 0;
    return true;
}

// alpha = r'z / p'Ap for the frames still iterating (0 for the others); p'Ap <= 0 / NaN -> breakdown
__device__ __forceinline__ void finalize_alpha(double pap, double* scal, int32_t* state, int64_t g, int G, int lane) {
    int32_t* active = state_ptr(state, g, MOF_I_ACTIVE);
    double alpha = 0.0;
    if (active[lane]) {
        const double rz = scal_ptr(scal, g, MOF_S_RZ)[lane];
        if (!(pap > 0.0) || isinf(pap)) {          // not SPD / NaN / overflow
            state_ptr(state, g, MOF_I_STATUS)[lane] = MOF_STATUS_BREAKDOWN;
            active[lane] = 0;
            atomicSub(lanes_active_ptr(state, G), 1);
        } else {
            alpha = rz / pap;
        }
    }
    scal_ptr(scal, g, MOF_S_PAP)[lane] = pap;
    scal_ptr(scal, g, MOF_S_ALPHA)[lane] = alpha;
}

// ---------------------------------------------------------------------------------
// K2: y = A x (block CSR, frame-minor), optionally x'y and the alpha step.
// ---------------------------------------------------------------------------------
template <bool SOLVER>
__global__ void __launch_bounds__(256) spmv_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                   const double* __restrict__ vals, const double* __restrict__ x,
                                                   double* __restrict__ y, int64_t N, int64_t nb, int ntiles,
                                                   double* __restrict__ partial, double* __restrict__ scal,
                                                   int32_t* __restrict__ state, int G) {
    const int64_t g = blockIdx.y;
    if (SOLVER && group_done_ptr(state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x;
    const int64_t row0 = (int64_t)tile * MOF_TILE_ROWS + warp * kRowsPerWarp;
    const double* __restrict__ vg = vals + (size_t)g * nb * 4 * MOF_W + lane;
    const double* __restrict__ xg = x + (size_t)g * N * 2 * MOF_W + lane;
    double* __restrict__ yg = y + (size_t)g * N * 2 * MOF_W + lane;

    int32_t rp = 0;
    if (lane <= kRowsPerWarp && row0 + lane <= N) rp = rowptr[row0 + lane];
    double dot = 0.0;
    // software-pipelined column indices: the first <=32 indices of the next row are
    // fetched while the current row is processed
    int32_t cj_next = 0;
    {
        int32_t bs = __shfl_sync(kFull, rp, 0), be = __shfl_sync(kFull, rp, 1);
        if (row0 < N && lane < be - bs) cj_next = col[bs + lane];
    }
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int64_t v = row0 + rr;
        if (v >= N) break;
        const int32_t bs = __shfl_sync(kFull, rp, rr), be = __shfl_sync(kFull, rp, rr + 1);
        int32_t cj = cj_next;
        if (rr + 1 < kRowsPerWarp && v + 1 < N) {
            int32_t be2 = __shfl_sync(kFull, rp, rr + 2 > kRowsPerWarp ? kRowsPerWarp : rr + 2);
            if (lane < be2 - be) cj_next = col[be + lane];
        }
        double y0 = 0.0, y1 = 0.0;
        for (int32_t base = bs; base < be; base += 32) {
            const int cnt = min(32, be - base);
            if (base != bs) cj = lane < cnt ? col[base + lane] : 0;
            const double* __restrict__ a = vg + (size_t)base * 4 * MOF_W;
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const int32_t j = __shfl_sync(kFull, cj, k);
                const double a00 = __ldcs(a + (size_t)k * 4 * MOF_W);
                const double a01 = __ldcs(a + (size_t)k * 4 * MOF_W + MOF_W);
                const double a10 = __ldcs(a + (size_t)k * 4 * MOF_W + 2 * MOF_W);
                const double a11 = __ldcs(a + (size_t)k * 4 * MOF_W + 3 * MOF_W);
                const double x0 = __ldg(xg + (size_t)j * 2 * MOF_W);
                const double x1 = __ldg(xg + (size_t)j * 2 * MOF_W + MOF_W);
                y0 = fma(a00, x0, y0);
                y0 = fma(a01, x1, y0);
                y1 = fma(a10, x0, y1);
                y1 = fma(a11, x1, y1);
            }
        }
        yg[(size_t)v * 2 * MOF_W] = y0;
        yg[(size_t)v * 2 * MOF_W + MOF_W] = y1;
        if (SOLVER) {
            const double xi0 = __ldg(xg + (size_t)v * 2 * MOF_W), xi1 = __ldg(xg + (size_t)v * 2 * MOF_W + MOF_W);
            dot = fma(xi0, y0, dot);
            dot = fma(xi1, y1, dot);
        }
    }
    if (!SOLVER) return;

    double val[1] = {dot}, tot[1];
    if (!tile_reduce<1>(val, partial + (size_t)g * ntiles * 2 * MOF_W, ntiles, tile, ticket_ptr(state, G) + g, tot))
        return;
    if (warp == 0) finalize_alpha(tot[0], scal, state, g, G, lane);
}

// ---------------------------------------------------------------------------------
// K3a: x += alpha p ; r -= alpha q ; z = M r ; r'z, r'r ; beta and convergence.
//   block Jacobi: q = ap = A p,            M = D^-1  (both in B.ap / B.minv)
//   SSOR:         q = t + w = At p (B.t + B.ap), M = Dt = D/omega
// A frame that meets its threshold is frozen with (alpha, beta, zscale) = (0, 1, 0), which
// leaves x, r, z, p and r'z untouched, so that it can resume exactly where it stopped if the
// true-residual check (init_kernel, MODE_VERIFY) asks for more iterations.
// ---------------------------------------------------------------------------------
// beta and the convergence test of group g from (r'z, r'r); called by warp 0 of one CTA per group.
__device__ __forceinline__ void update_scalar_step(const mof_batch_dev& B, int64_t g, const double (&tot)[2]) {
    const int G = B.n_groups;
    const int lane = threadIdx.x & 31;
    int32_t* active = state_ptr(B.state, g, MOF_I_ACTIVE);
    double beta = 1.0, zs = 0.0;                    // frozen: p stays as it is
    int act = active[lane];
    const int was = act;
    if (act) {
        const double rz_old = scal_ptr(B.scal, g, MOF_S_RZ)[lane];
        const double bb = scal_ptr(B.scal, g, MOF_S_BB)[lane];
        const double thr = scal_ptr(B.scal, g, MOF_S_THR)[lane];
        state_ptr(B.state, g, MOF_I_ITERS)[lane] += 1;
        scal_ptr(B.scal, g, MOF_S_RZ)[lane] = tot[0];
        scal_ptr(B.scal, g, MOF_S_RR)[lane] = tot[1];
        if (!isfinite(tot[0]) || !isfinite(tot[1])) {
            state_ptr(B.state, g, MOF_I_STATUS)[lane] = MOF_STATUS_BREAKDOWN;
            act = 0;
        } else if (tot[1] <= thr * bb) {
            state_ptr(B.state, g, MOF_I_STATUS)[lane] = MOF_STATUS_CONVERGED;
            scal_ptr(B.scal, g, MOF_S_BETA_SAVED)[lane] = tot[0] / rz_old;   // used if the frame resumes
            act = 0;
        } else {
            beta = tot[0] / rz_old;
            zs = 1.0;
        }
        active[lane] = act;
    }
    scal_ptr(B.scal, g, MOF_S_BETA)[lane] = beta;
    scal_ptr(B.scal, g, MOF_S_ZS)[lane] = zs;
    const int any = __any_sync(kFull, act);
    const int dropped = __popc(__ballot_sync(kFull, was && !act));
    if (lane == 0) {
        if (dropped) atomicSub(lanes_active_ptr(B.state, G), dropped);
        if (!any) {
            group_done_ptr(B.state, G)[g] = 1;
            atomicSub(groups_active_ptr(B.state, G), 1);
        }
    }
}

// Tile-level end of the vector update: deterministic reduction of (r'z, r'r), then beta and the convergence
// test in the last CTA of the group.
template <bool SSOR>
__device__ __forceinline__ void update_finish(const mof_batch_dev& B, int ntiles, double inv_omega, int tile, int64_t g,
                                              double rz, double rr) {
    const int G = B.n_groups;
    const int warp = threadIdx.x >> 5;
    if (SSOR) rz = rr * inv_omega;                     // r'z with z = r / omega (linear, so per-lane partials add up)
    double val[2] = {rz, rr}, tot[2];
    if (!tile_reduce<2>(val, B.partial + (size_t)g * ntiles * 2 * MOF_W, ntiles, tile, ticket_ptr(B.state, G) + g, tot))
        return;
    if (warp == 0) update_scalar_step(B, g, tot);
}

template <bool SSOR>
__device__ __forceinline__ void update_body(const mof_batch_dev& B, int64_t N, int ntiles, double inv_omega, int tile, int64_t g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)tile * MOF_TILE_ROWS + warp * kRowsPerWarp;
    const double alpha = scal_ptr(B.scal, g, MOF_S_ALPHA)[lane];
    double rz = 0.0, rr = 0.0;
#pragma unroll 2
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        double a0 = __ldcs(B.ap + i0), a1 = __ldcs(B.ap + i1);
        if (SSOR) { a0 += __ldcs(B.t + i0); a1 += __ldcs(B.t + i1); }
        double r0 = B.r[i0], r1 = B.r[i1];
        if (!SSOR) {                                   // SSOR: x += alpha p is folded into the next backward sweep
            B.x[i0] = fma(alpha, B.p[i0], B.x[i0]);
            B.x[i1] = fma(alpha, B.p[i1], B.x[i1]);
        }
        r0 = fma(-alpha, a0, r0);
        r1 = fma(-alpha, a1, r1);
        B.r[i0] = r0; B.r[i1] = r1;
        if (!SSOR) {                                   // z = D^-1 r ; SSOR: z = r / omega is never stored
            const size_t im = mof_ix_minv(N, g, v, 0) + lane;
            const double m0 = B.minv[im], m1 = B.minv[im + MOF_W], m2 = B.minv[im + 2 * MOF_W];
            const double z0 = m0 * r0 + m1 * r1;
            const double z1 = m1 * r0 + m2 * r1;
            B.z[i0] = z0; B.z[i1] = z1;
            rz = fma(r0, z0, rz); rz = fma(r1, z1, rz);
        }
        rr = fma(r0, r0, rr); rr = fma(r1, r1, rr);
    }
    update_finish<SSOR>(B, ntiles, inv_omega, tile, g, rz, rr);
}

template <bool SSOR>
__global__ void __launch_bounds__(256) update_kernel(mof_batch_dev B, int64_t N, int ntiles, double inv_omega) {
    if (group_done_ptr(B.state, B.n_groups)[blockIdx.y]) return;
    update_body<SSOR>(B, N, ntiles, inv_omega, blockIdx.x, blockIdx.y);
}

// K3b (block Jacobi): p = zs z + beta p at the start of an iteration
__global__ void __launch_bounds__(256) pupdate_kernel(mof_batch_dev B, int64_t N) {
    const int64_t g = blockIdx.y;
    if (group_done_ptr(B.state, B.n_groups)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    const double beta = scal_ptr(B.scal, g, MOF_S_BETA)[lane];
    const double zs = scal_ptr(B.scal, g, MOF_S_ZS)[lane];
#pragma unroll 4
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        const double z0 = __ldcs(B.z + i0), z1 = __ldcs(B.z + i1);
        B.p[i0] = fma(beta, B.p[i0], zs * z0);
        B.p[i1] = fma(beta, B.p[i1], zs * z1);
    }
}

// ---------------------------------------------------------------------------------
// SSOR sweeps: one warp = one patch (tile) of the launch's colour, lane = frame.  The rows of
// a patch are solved sequentially by the warp (mof_sweep_*_body); patches of one colour do not
// touch each other, earlier colours are complete because they ran in earlier launches.
// ---------------------------------------------------------------------------------
// Pull the 4 x 256-byte lines of matrix block b of this group towards L2 (one address per lane
// covers the line pair of a component).
__device__ __forceinline__ void prefetch_block_l2(const double* __restrict__ vals_l, int32_t b) {
    const double* p = vals_l + (size_t)b * 4 * MOF_W;
#pragma unroll
    for (int c = 0; c < 4; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + c * MOF_W));
}

// Row arithmetic of the SSOR sweeps with every rounding spelled out (no compiler-chosen contraction), so
// that the patch sweeps, the per-level launches and the persistent level kernel produce bit-identical
// rows whatever code surrounds them.
__device__ __forceinline__ void row_sub_block(double& a0, double& a1, double A0, double A1, double A2, double A3,
                                              double v0, double v1) {
    a0 = __dsub_rn(a0, __fma_rn(A1, v1, __dmul_rn(A0, v0)));
    a1 = __dsub_rn(a1, __fma_rn(A3, v1, __dmul_rn(A2, v0)));
}
__device__ __forceinline__ double row_dir(double zsw, double r, double beta, double p) {       // zsw r + beta p
    return __fma_rn(zsw, r, __dmul_rn(beta, p));
}
__device__ __forceinline__ double row_pdot(double p0, double p1, double q0, double q1) {       // p0 q0 + p1 q1
    return __fma_rn(p1, q1, __dmul_rn(p0, q0));
}

// Validity bit of the level-scheduled sweeps.  The entries of t (backward sweep) and w (forward sweep) that
// other rows gather carry, in the least significant mantissa bit, the parity of the frame's iteration counter
// (state ITERS) at the time they were written: an active frame rewrites every entry once per iteration with
// the opposite parity, so a reader that knows the frame's counter tells this iteration's value from last
// iteration's by looking at the value itself -- one 8-byte load is both the "is it ready" poll and the gather,
// and the writer needs no fence and no flag (a naturally aligned 8-byte store is single-copy atomic).  A frozen
// frame keeps its counter, p and therefore t and w: for it the old and the new entry are the same number, so
// either may be read.  The bit costs one ulp of t / w inside the preconditioner application, is a function of
// the frame's own iteration count only (results stay independent of batching) and is applied by every kernel
// of the level path, per-level launches included, so the paths stay bit-identical.
__device__ __forceinline__ double with_parity(double v, int par) {
    return __longlong_as_double((__double_as_longlong(v) & ~1LL) | (long long)par);
}
__device__ __forceinline__ bool has_parity(double v, int par) { return ((int)__double_as_longlong(v) & 1) == par; }
__device__ __forceinline__ double ld_strong(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_strong(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Exact, order-free accumulation of p'Ap on the level path.  A row's share p_i.(t_i + w_i) is converted to a
// 128-bit fixed-point integer (value x 2^k, floor) and integers are added -- associative, so warps may add the
// shares of whatever rows they happen to own, in any order, and flush into one accumulator per frame with two
// 64-bit atomics: the sum is independent of grid size, batch composition and GPU, like the fixed-order tile
// reduction it replaces, but needs no per-row buffer (16 N bytes per frame-iteration less), no read-back pass
// and no barrier of its own.  k = 90 - exponent(previous p'Ap, or r'z before the first iteration): shares up to
// 2^37 times the reference fit, resolution 2^-90 of it (a double carries 2^-53).  A share that is not finite or
// does not fit sets the frame's flag and the frame ends as a breakdown.
struct Fx128 {
    unsigned long long lo;
    long long hi;
};
__device__ __forceinline__ bool fx_from_double(double v, int k, Fx128& out) {
    const double t = scalbn(v, k - 64);                  // integer part = bits 64..127 of v 2^k
    if (!(fabs(t) < 9.0e18)) { out.lo = 0; out.hi = 0; return false; }
    const double fl = floor(t);
    out.hi = (long long)fl;
    out.lo = __double2ull_rz(scalbn(t - fl, 64));        // t - fl in [0, 1): exact, below 2^64
    return true;
}
__device__ __forceinline__ void fx_add(Fx128& acc, const Fx128& x) {
    acc.lo += x.lo;
    acc.hi += x.hi + (acc.lo < x.lo ? 1 : 0);
}
__device__ __forceinline__ double fx_to_double(unsigned long long lo, long long hi, int k) {
    const bool neg = hi < 0;
    if (neg) { lo = ~lo + 1ull; hi = ~hi + (lo == 0ull ? 1 : 0); }
    const double d = scalbn((double)(unsigned long long)hi, 64) + (double)lo;
    return scalbn(neg ? -d : d, -k);
}
__device__ __forceinline__ unsigned long long* fx_word(double* scal, int64_t g, int which) {
    return reinterpret_cast<unsigned long long*>(scal_ptr(scal, g, which));
}
// add a warp's partial sum (per lane) into the frame accumulators of group g
__device__ __forceinline__ void fx_flush(double* scal, int64_t g, int lane, Fx128& acc, bool& bad) {
    if (acc.lo | (unsigned long long)acc.hi) {
        const unsigned long long old = atomicAdd(fx_word(scal, g, MOF_S_FX_LO) + lane, acc.lo);
        const unsigned long long carry = (old + acc.lo) < acc.lo ? 1ull : 0ull;
        atomicAdd(fx_word(scal, g, MOF_S_FX_HI) + lane, (unsigned long long)acc.hi + carry);
    }
    if (bad) atomicOr(fx_word(scal, g, MOF_S_FX_BAD) + lane, 1ull);
    acc.lo = 0; acc.hi = 0; bad = false;
}
__device__ __forceinline__ int fx_scale_for(double ref) {          // k for the NEXT accumulation given a positive reference
    return (ref > 0.0 && isfinite(ref)) ? 90 - ilogb(ref) : 0;
}
// p'Ap of group g from the accumulators -> alpha; accumulators cleared, scale set for the next iteration (warp 0 of one CTA)
__device__ __forceinline__ void fx_finalize_alpha(double* scal, int32_t* state, int64_t g, int G, int lane) {
    unsigned long long* lo = fx_word(scal, g, MOF_S_FX_LO) + lane;
    unsigned long long* hi = fx_word(scal, g, MOF_S_FX_HI) + lane;
    unsigned long long* kk = fx_word(scal, g, MOF_S_FX_K) + lane;
    unsigned long long* bad = fx_word(scal, g, MOF_S_FX_BAD) + lane;
    const int k = (int)(long long)__ldcg(kk);
    double pap = fx_to_double(__ldcg(lo), (long long)__ldcg(hi), k);
    if (__ldcg(bad)) pap = nan("");
    *lo = 0ull; *hi = 0ull; *bad = 0ull;
    if (pap > 0.0 && isfinite(pap)) *kk = (unsigned long long)(long long)fx_scale_for(pap);
    finalize_alpha(pap, scal, state, g, G, lane);
}

// r'r of group g from the accumulators -> r'z = r'r / omega, beta and the convergence test; accumulators cleared,
// scale set for the next iteration (warp 0 of one CTA).
__device__ __forceinline__ void fx_finalize_beta(const mof_batch_dev& B, int64_t g, int lane, double inv_omega) {
    unsigned long long* lo = fx_word(B.scal, g, MOF_S_FX_LO) + lane;
    unsigned long long* hi = fx_word(B.scal, g, MOF_S_FX_HI) + lane;
    unsigned long long* kk = fx_word(B.scal, g, MOF_S_FX_K_RR) + lane;
    unsigned long long* bad = fx_word(B.scal, g, MOF_S_FX_BAD) + lane;
    const int k = (int)(long long)__ldcg(kk);
    double rr = fx_to_double(__ldcg(lo), (long long)__ldcg(hi), k);
    if (__ldcg(bad)) rr = nan("");
    *lo = 0ull; *hi = 0ull; *bad = 0ull;
    if (rr > 0.0 && isfinite(rr)) *kk = (unsigned long long)(long long)fx_scale_for(rr);
    const double tot[2] = {rr * inv_omega, rr};          // r'z with z = r / omega
    update_scalar_step(B, g, tot);
}

// Level path, per-level launches: the SSOR update with r'r accumulated in fixed point like the persistent kernel
// (a warp's eight rows are summed in order in fp64, that partial goes into the accumulator), then level_beta_kernel.
__global__ void __launch_bounds__(256) level_update_kernel(mof_batch_dev B, int64_t N) {
    const int64_t g = blockIdx.y;
    if (group_done_ptr(B.state, B.n_groups)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    if (row0 >= N) return;
    const double alpha = scal_ptr(B.scal, g, MOF_S_ALPHA)[lane];
    double rr = 0.0;
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        const double a0 = __ldcs(B.ap + i0) + __ldcs(B.t + i0), a1 = __ldcs(B.ap + i1) + __ldcs(B.t + i1);
        const double r0 = fma(-alpha, a0, B.r[i0]), r1 = fma(-alpha, a1, B.r[i1]);
        B.r[i0] = r0; B.r[i1] = r1;
        rr = fma(r0, r0, rr); rr = fma(r1, r1, rr);
    }
    Fx128 x, acc = {0ull, 0ll};
    bool bad = !fx_from_double(rr, (int)(long long)fx_word(B.scal, g, MOF_S_FX_K_RR)[lane], x);
    fx_add(acc, x);
    fx_flush(B.scal, g, lane, acc, bad);
}

__global__ void __launch_bounds__(32) level_beta_kernel(mof_batch_dev B, double inv_omega) {
    const int64_t g = blockIdx.x;
    if (group_done_ptr(B.state, B.n_groups)[g]) return;
    fx_finalize_beta(B, g, threadIdx.x, inv_omega);
}

// What a sweep knows about a row one step before it is processed: its block range, the first
// four column indices and the row's entries of the two streamed vectors.  Loading these one row
// ahead takes the DRAM latency of the streamed vectors and the index latency off the per-row
// critical path (ncu: 32 % of the stall samples sat on the first use of r/p, 8 % on the indices).
struct RowPre {
    int32_t bs, be;
    int32_t c[4];
    double s[6];
};

// Off-diagonal part of one row of a sweep: acc -= sum_k A_k v[col_k] over blocks [bs, be), in two
// steps so that the caller can put independent work (the next row's loads, L2 prefetches) between
// issuing the loads and consuming them.  All loads of up to four blocks are issued before the
// first use: one memory round trip serves the row (the matrix values are independent of the
// sweep's recurrence; only the gathered v values may have been written by this thread a few rows
// earlier, hence plain loads for them).
struct RowBatch {
    double a[4][4], v[4][2];
};

__device__ __forceinline__ void sweep_row_issue(const double* __restrict__ vals_l, const double* v_l, const RowPre& R,
                                                RowBatch& Bt) {
    const int cnt = R.be - R.bs;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cnt) {
            const double* ap = vals_l + (size_t)(R.bs + k) * 4 * MOF_W;
            Bt.a[k][0] = __ldcs(ap);
            Bt.a[k][1] = __ldcs(ap + MOF_W);
            Bt.a[k][2] = __ldcs(ap + 2 * MOF_W);
            Bt.a[k][3] = __ldcs(ap + 3 * MOF_W);
            Bt.v[k][0] = v_l[(size_t)(2 * (int64_t)R.c[k]) * MOF_W];
            Bt.v[k][1] = v_l[(size_t)(2 * (int64_t)R.c[k] + 1) * MOF_W];
        }
    }
}

__device__ __forceinline__ void sweep_row_consume(const int32_t* __restrict__ col, const double* __restrict__ vals_l,
                                                  const double* v_l, const RowPre& R, const RowBatch& Bt, double& a0,
                                                  double& a1) {
    const int cnt = R.be - R.bs;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cnt) row_sub_block(a0, a1, Bt.a[k][0], Bt.a[k][1], Bt.a[k][2], Bt.a[k][3], Bt.v[k][0], Bt.v[k][1]);
    }
    for (int32_t b = R.bs + 4; b < R.be; ++b) {            // rows with more than four blocks in this half (rare)
        const int64_t j = col[b];
        const double* ap = vals_l + (size_t)b * 4 * MOF_W;
        const double v0 = v_l[(size_t)(2 * j) * MOF_W], v1 = v_l[(size_t)(2 * j + 1) * MOF_W];
        row_sub_block(a0, a1, __ldcs(ap), __ldcs(ap + MOF_W), __ldcs(ap + 2 * MOF_W), __ldcs(ap + 3 * MOF_W), v0, v1);
    }
}

// MODE 0: iteration (x += alpha p_old and p <- zs r/omega + beta p_old fused in, t = (Dt+U)^-1 p).
// MODE 1: back-transform t = (Dt+U)^-1 (x + alpha p), i.e. of the iterate including its pending step.
// Same arithmetic as mof_sweep_back_body (mof_bodies.h), with the loads of a row batched and the
// next row's streamed data fetched one row ahead.
template <int MODE>
__global__ void __launch_bounds__(256, 2) sweep_back_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            const int32_t* __restrict__ diag, mof_batch_dev B,
                                                            double* tout, int64_t N, int64_t nb, int tile0, int tile1,
                                                            double omega) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    if (MODE == 0 && group_done_ptr(B.state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = tile0 + blockIdx.x * kWarps + warp;
    if (tile >= tile1) return;
    const int64_t r0 = (int64_t)tile * MOF_TILE_ROWS;
    const int64_t r1 = min(N, r0 + (int64_t)MOF_TILE_ROWS);
    const int nrows = (int)(r1 - r0);
    double beta = 0.0, zsw = 0.0;
    // x += alpha p of the PREVIOUS iteration is applied here, where p is read anyway (the vector
    // kernel then never touches x or p): alpha is the step the last forward sweep computed; it is 0
    // once a frozen frame's last step has been applied
    const double alpha = scal_ptr(B.scal, g, MOF_S_ALPHA)[lane];
    if (MODE == 0) {
        beta = scal_ptr(B.scal, g, MOF_S_BETA)[lane];
        zsw = scal_ptr(B.scal, g, MOF_S_ZS)[lane] / omega;         // z = Dt r = r / omega
    }
    const double* __restrict__ vals_l = B.vals + (size_t)g * nb * 4 * MOF_W + lane;
    const double* __restrict__ r_l = B.r + (size_t)g * N * 2 * MOF_W + lane;
    double* p_l = B.p + (size_t)g * N * 2 * MOF_W + lane;
    double* x_l = B.x + (size_t)g * N * 2 * MOF_W + lane;
    double* t_l = tout + (size_t)g * N * 2 * MOF_W + lane;
    // row pointers of the patch: lane q holds rowptr[r0+q+1] and diag[r0+q] (two coalesced loads)
    int32_t rp_lo = 0, rp_hi = 0, dg_lo = 0, dg_hi = 0;
    if (r0 + lane < r1) { rp_lo = rowptr[r0 + lane + 1]; dg_lo = diag[r0 + lane]; }
    if (r0 + 32 + lane < r1) { rp_hi = rowptr[r0 + 32 + lane + 1]; dg_hi = diag[r0 + 32 + lane]; }
    auto load_row = [&](int q, RowPre& R) {
        R.be = q < 32 ? __shfl_sync(kFull, rp_lo, q) : __shfl_sync(kFull, rp_hi, q - 32);
        R.bs = (q < 32 ? __shfl_sync(kFull, dg_lo, q) : __shfl_sync(kFull, dg_hi, q - 32)) + 1;
        const size_t i = (size_t)(r0 + q);
        R.s[0] = p_l[(2 * i) * MOF_W];
        R.s[1] = p_l[(2 * i + 1) * MOF_W];
        if (MODE == 0) {
            R.s[2] = r_l[(2 * i) * MOF_W];
            R.s[3] = r_l[(2 * i + 1) * MOF_W];
        }
        R.s[4] = x_l[(2 * i) * MOF_W];
        R.s[5] = x_l[(2 * i + 1) * MOF_W];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (R.bs + k < R.be) R.c[k] = col[R.bs + k];
    };
    RowPre cur, nxt;
    load_row(nrows - 1, cur);
    for (int q = nrows - 1; q >= 0; --q) {
        const size_t i = (size_t)(r0 + q);
        RowBatch bt;
        sweep_row_issue(vals_l, t_l, cur, bt);                      // this row's matrix values and gathered t
        if (q > 0) load_row(q - 1, nxt);
        if (q >= kPrefetchRows) {                                   // matrix values of a row a few steps ahead -> L2
            const int qq = q - kPrefetchRows;
            const int32_t pe = qq < 32 ? __shfl_sync(kFull, rp_lo, qq) : __shfl_sync(kFull, rp_hi, qq - 32);
            const int32_t ps = (qq < 32 ? __shfl_sync(kFull, dg_lo, qq) : __shfl_sync(kFull, dg_hi, qq - 32)) + 1;
            for (int32_t b = ps; b < pe; ++b) prefetch_block_l2(vals_l, b);
        }
        // pending step of the previous iteration: x += alpha p_old
        const double x0 = fma(alpha, cur.s[0], cur.s[4]), x1 = fma(alpha, cur.s[1], cur.s[5]);
        double a0, a1;
        if (MODE == 0) {
            x_l[(2 * i) * MOF_W] = x0;
            x_l[(2 * i + 1) * MOF_W] = x1;
            a0 = row_dir(zsw, cur.s[2], beta, cur.s[0]);
            a1 = row_dir(zsw, cur.s[3], beta, cur.s[1]);
            p_l[(2 * i) * MOF_W] = a0;
            p_l[(2 * i + 1) * MOF_W] = a1;
        } else {                                                    // back-transform of the complete iterate
            a0 = x0;
            a1 = x1;
        }
        sweep_row_consume(col, vals_l, t_l, cur, bt, a0, a1);
        t_l[(2 * i) * MOF_W] = omega * a0;
        t_l[(2 * i + 1) * MOF_W] = omega * a1;
        cur = nxt;
    }
}

// MODE 0: iteration (w = (Dt+L)^-1 (p - ((2-omega)/omega) t), p'(t+w) -> alpha in the last CTA of the
// last colour).  MODE 1: wout = (Dt+L)^-1 pin.  Same arithmetic as mof_sweep_fwd_body.
template <int MODE>
__global__ void __launch_bounds__(256, 2) sweep_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                           const int32_t* __restrict__ diag, mof_batch_dev B, const double* pin,
                                                           double* wout, int64_t N, int64_t nb, int tile0, int tile1,
                                                           int ntiles, double omega) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    if (MODE == 0 && group_done_ptr(B.state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = tile0 + blockIdx.x * kWarps + warp;
    const bool valid = tile < tile1;
    double dot = 0.0;
    if (valid) {
        const int64_t r0 = (int64_t)tile * MOF_TILE_ROWS;
        const int64_t r1 = min(N, r0 + (int64_t)MOF_TILE_ROWS);
        const int nrows = (int)(r1 - r0);
        const double* __restrict__ vals_l = B.vals + (size_t)g * nb * 4 * MOF_W + lane;
        const double* __restrict__ p_l = pin + (size_t)g * N * 2 * MOF_W + lane;
        const double* __restrict__ t_l = B.t + (size_t)g * N * 2 * MOF_W + lane;
        double* w_l = wout + (size_t)g * N * 2 * MOF_W + lane;
        int32_t rp_lo = 0, rp_hi = 0, dg_lo = 0, dg_hi = 0;
        if (r0 + lane < r1) { rp_lo = rowptr[r0 + lane]; dg_lo = diag[r0 + lane]; }
        if (r0 + 32 + lane < r1) { rp_hi = rowptr[r0 + 32 + lane]; dg_hi = diag[r0 + 32 + lane]; }
        const double kscale = (2.0 - omega) / omega;
        auto load_row = [&](int q, RowPre& R) {
            R.bs = q < 32 ? __shfl_sync(kFull, rp_lo, q) : __shfl_sync(kFull, rp_hi, q - 32);
            R.be = q < 32 ? __shfl_sync(kFull, dg_lo, q) : __shfl_sync(kFull, dg_hi, q - 32);
            const size_t i = (size_t)(r0 + q);
            R.s[0] = p_l[(2 * i) * MOF_W];
            R.s[1] = p_l[(2 * i + 1) * MOF_W];
            if (MODE == 0) {
                R.s[2] = t_l[(2 * i) * MOF_W];
                R.s[3] = t_l[(2 * i + 1) * MOF_W];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (R.bs + k < R.be) R.c[k] = col[R.bs + k];
        };
        RowPre cur, nxt;
        load_row(0, cur);
        for (int q = 0; q < nrows; ++q) {
            const size_t i = (size_t)(r0 + q);
            RowBatch bt;
            sweep_row_issue(vals_l, w_l, cur, bt);                  // this row's matrix values and gathered w
            if (q + 1 < nrows) load_row(q + 1, nxt);
            if (q + kPrefetchRows < nrows) {
                const int qq = q + kPrefetchRows;
                const int32_t ps = qq < 32 ? __shfl_sync(kFull, rp_lo, qq) : __shfl_sync(kFull, rp_hi, qq - 32);
                const int32_t pe = qq < 32 ? __shfl_sync(kFull, dg_lo, qq) : __shfl_sync(kFull, dg_hi, qq - 32);
                for (int32_t b = ps; b < pe; ++b) prefetch_block_l2(vals_l, b);
            }
            const double p0 = cur.s[0], p1 = cur.s[1];
            double a0 = p0, a1 = p1, t0 = 0.0, t1 = 0.0;
            if (MODE == 0) {
                t0 = cur.s[2];
                t1 = cur.s[3];
                a0 = __fma_rn(-kscale, t0, a0);
                a1 = __fma_rn(-kscale, t1, a1);
            }
            sweep_row_consume(col, vals_l, w_l, cur, bt, a0, a1);
            const double o0 = omega * a0, o1 = omega * a1;
            w_l[(2 * i) * MOF_W] = o0;
            w_l[(2 * i + 1) * MOF_W] = o1;
            if (MODE == 0) dot += row_pdot(p0, p1, t0 + o0, t1 + o1);
            cur = nxt;
        }
    }
    if (MODE != 0) return;
    // one partial per patch; the CTA that completes the count over all colours reduces them
    __shared__ double red[kWarps][1][MOF_W];
    __shared__ int s_last;
    double* partial_g = B.partial + (size_t)g * ntiles * 2 * MOF_W;
    if (valid) partial_g[(size_t)tile * 2 * MOF_W + lane] = dot;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cnt = min(kWarps, tile1 - (tile0 + (int)blockIdx.x * kWarps));
        s_last = (atomicAdd(ticket_ptr(B.state, G) + g, cnt) + cnt == ntiles);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[1];
    reduce_all_tiles<1>(partial_g, ntiles, tot, red);
    if (threadIdx.x == 0) ticket_ptr(B.state, G)[g] = 0;
    if (warp == 0) finalize_alpha(tot[0], B.scal, B.state, g, G, lane);
}

// ---------------------------------------------------------------------------------
// Level-scheduled SSOR sweeps (mesh built with reorder = 3): the rows of one dependency level are
// independent, so a launch covers the rows [r_lo, r_hi) of one level with ONE WARP PER ROW (lane =
// frame) and no recurrence inside the launch; levels run in sequence (descending for the backward
// sweep, ascending for the forward one).  Same per-row arithmetic as the patch sweeps above.  The
// ordering is the Cuthill-McKee one regrouped by level, whose SSOR needs ~2.4x fewer iterations at
// ico7 than the block-multicolour ordering (148 vs 361 at omega 1.85 / 1.4).
// ---------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) level_back_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                         const int32_t* __restrict__ diag, mof_batch_dev B, double* tout,
                                                         int64_t N, int64_t nb, int r_lo, int r_hi, double omega) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    if (MODE == 0 && group_done_ptr(B.state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = r_lo + blockIdx.x * kWarps + warp;
    if (row >= r_hi) return;
    const size_t i = (size_t)row;
    const double* __restrict__ vals_l = B.vals + (size_t)g * nb * 4 * MOF_W + lane;
    const double* __restrict__ r_l = B.r + (size_t)g * N * 2 * MOF_W + lane;
    double* p_l = B.p + (size_t)g * N * 2 * MOF_W + lane;
    double* x_l = B.x + (size_t)g * N * 2 * MOF_W + lane;
    double* t_l = tout + (size_t)g * N * 2 * MOF_W + lane;
    RowPre R;
    R.be = rowptr[row + 1];
    R.bs = diag[row] + 1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (R.bs + k < R.be) R.c[k] = col[R.bs + k];
    RowBatch bt;
    sweep_row_issue(vals_l, t_l, R, bt);                            // upper blocks and t of the later levels
    const double p0 = p_l[(2 * i) * MOF_W], p1 = p_l[(2 * i + 1) * MOF_W];
    const double x0 = x_l[(2 * i) * MOF_W], x1 = x_l[(2 * i + 1) * MOF_W];
    const double alpha = scal_ptr(B.scal, g, MOF_S_ALPHA)[lane];
    const double y0 = fma(alpha, p0, x0), y1 = fma(alpha, p1, x1);  // pending step of the previous iteration
    double a0, a1;
    if (MODE == 0) {
        const double beta = scal_ptr(B.scal, g, MOF_S_BETA)[lane];
        const double zsw = scal_ptr(B.scal, g, MOF_S_ZS)[lane] / omega;
        const double r0 = r_l[(2 * i) * MOF_W], r1 = r_l[(2 * i + 1) * MOF_W];
        x_l[(2 * i) * MOF_W] = y0;
        x_l[(2 * i + 1) * MOF_W] = y1;
        a0 = row_dir(zsw, r0, beta, p0);
        a1 = row_dir(zsw, r1, beta, p1);
        p_l[(2 * i) * MOF_W] = a0;
        p_l[(2 * i + 1) * MOF_W] = a1;
    } else {
        a0 = y0;
        a1 = y1;
    }
    sweep_row_consume(col, vals_l, t_l, R, bt, a0, a1);
    // validity bit (see with_parity): iteration -> parity of the frame's counter, back-transform -> the parity of
    // the last iteration, so that a resumed frame's next sweep still finds the opposite one
    const int par = (state_ptr(B.state, g, MOF_I_ITERS)[lane] + MODE) & 1;
    t_l[(2 * i) * MOF_W] = with_parity(omega * a0, par);
    t_l[(2 * i + 1) * MOF_W] = with_parity(omega * a1, par);
}

// MODE 0 also adds the row's share of p'(t+w) to the frame's fixed-point accumulator; level_alpha_kernel turns it into alpha.
template <int MODE>
__global__ void __launch_bounds__(256) level_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                        const int32_t* __restrict__ diag, mof_batch_dev B, const double* pin,
                                                        double* wout, int64_t N, int64_t nb, int r_lo,
                                                        int r_hi, double omega) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    if (MODE == 0 && group_done_ptr(B.state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = r_lo + blockIdx.x * kWarps + warp;
    if (row >= r_hi) return;
    const size_t i = (size_t)row;
    const double* __restrict__ vals_l = B.vals + (size_t)g * nb * 4 * MOF_W + lane;
    const double* __restrict__ p_l = pin + (size_t)g * N * 2 * MOF_W + lane;
    const double* __restrict__ t_l = B.t + (size_t)g * N * 2 * MOF_W + lane;
    double* w_l = wout + (size_t)g * N * 2 * MOF_W + lane;
    RowPre R;
    R.bs = rowptr[row];
    R.be = diag[row];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (R.bs + k < R.be) R.c[k] = col[R.bs + k];
    RowBatch bt;
    sweep_row_issue(vals_l, w_l, R, bt);                            // lower blocks and w of the earlier levels
    const double p0 = p_l[(2 * i) * MOF_W], p1 = p_l[(2 * i + 1) * MOF_W];
    double a0 = p0, a1 = p1, t0 = 0.0, t1 = 0.0;
    if (MODE == 0) {
        const double kscale = (2.0 - omega) / omega;
        t0 = t_l[(2 * i) * MOF_W];
        t1 = t_l[(2 * i + 1) * MOF_W];
        a0 = __fma_rn(-kscale, t0, a0);
        a1 = __fma_rn(-kscale, t1, a1);
    }
    sweep_row_consume(col, vals_l, w_l, R, bt, a0, a1);
    double o0 = omega * a0, o1 = omega * a1;
    if (MODE == 0) {
        const int par = state_ptr(B.state, g, MOF_I_ITERS)[lane] & 1;
        o0 = with_parity(o0, par);
        o1 = with_parity(o1, par);
    }
    w_l[(2 * i) * MOF_W] = o0;
    w_l[(2 * i + 1) * MOF_W] = o1;
    if (MODE == 0) {                                 // the row's share of p'Ap, exact fixed-point accumulation
        Fx128 x, acc = {0ull, 0ll};
        const int k = (int)(long long)fx_word(B.scal, g, MOF_S_FX_K)[lane];
        bool bad = !fx_from_double(row_pdot(p0, p1, t0 + o0, t1 + o1), k, x);
        fx_add(acc, x);
        fx_flush(B.scal, g, lane, acc, bad);
    }
}

// per-level launches: p'Ap of every group from its accumulators -> alpha (one warp per group)
__global__ void __launch_bounds__(32) level_alpha_kernel(mof_batch_dev B) {
    const int64_t g = blockIdx.x;
    if (group_done_ptr(B.state, B.n_groups)[g]) return;
    fx_finalize_alpha(B.scal, B.state, g, B.n_groups, threadIdx.x);
}

// ---------------------------------------------------------------------------------
// Persistent level-scheduled SSOR iteration (the default of the level path).
//
// The per-level launches above cost a kernel boundary (a GPU-wide barrier plus a launch gap) per
// dependency level, 2 x 1024 per iteration at ico7, and leave every row's loads exposed behind it.  Here
// ONE cooperative launch runs `n_iter` whole PCG iterations.  Inside a sweep there is no barrier at all:
// the rows of all active groups form one static work list in dependency order (blocks of eight rows, group by
// group inside a block; backward: rows descending; A = groups still iterating), warp w of the grid owns items
// w, w + W, ... and a row waits only for the rows it actually reads:
//   * iteration sweeps: the gathered t / w entries carry a validity bit (with_parity): ONE 8-byte
//     ld.relaxed.gpu is both the "is it ready" poll and the gather; the writer needs neither fence nor flag;
//   * the three other sweeps of a solve (start transform, back-transforms): an int32 `ready` stamp per
//     (group, row) -- consumer ld.acquire.gpu, producer stores, fence, stamp.
// Every item depends on items earlier in the list only, each warp walks its items in list order and all
// CTAs are co-resident (cooperative launch), so the earliest unfinished item can always run: no deadlock.
// Everything of a row that does not depend on the sweep (its matrix blocks -- contiguous in the level-major
// numbering -- and its own entries of p, r, x / p, t) is fetched ahead by 1-D bulk async copies (cp.async.bulk
// -> UBLKCP, completion on an mbarrier) into a ring of shared-memory stages per warp (StageCfg), so the
// dependent chain of a row is gather -> FMAs -> store.
// The phases of an iteration (backward sweep, forward sweep, alpha, r update, beta and the convergence test)
// are separated by five grid barriers; p'Ap and r'r are accumulated exactly in 128-bit fixed point (Fx128), so
// no reduction has an order; alpha, beta and the group_done flags never leave the device.  Per-row arithmetic
// is that of the per-level kernels: results are bit-identical to them and independent of grid size, batch
// composition and GPU.
// ---------------------------------------------------------------------------------
constexpr int kStageVecBytes = 2 * MOF_W * 8;                         // one row of one vector: 512
constexpr int kSlotBytes = 2048;                                      // r update: three vectors x four rows = 3 x 2048 per piece
constexpr int kPieceBytes = 3 * kSlotBytes;
// Stage configuration of the sweeps: SB matrix blocks of a row are staged in shared memory (further blocks of a
// longer row come through ordinary loads), NS stages per warp = NS - 1 items in flight behind the one being
// computed.  A sweep's rate follows its items in flight (DRAM latency under load is several times one item's
// dependent chain), and shared memory bounds them: <4, 2> (5.5 KB stages) suits meshes where many rows have four
// blocks on one side of the diagonal, <3, 3> (4.5 KB stages, 216 KB per SM) the regular ones (ico7: 99.6 % of the
// rows have three); the host picks by mesh->level_stage_blocks.
template <int SB, int NS>
struct StageCfg {
    static constexpr int kBlocks = SB, kStages = NS;
    static constexpr int kValBytes = SB * 4 * MOF_W * 8;
    static constexpr int kBytes = kValBytes + 3 * kStageVecBytes;
    static constexpr int kWarpBytes = NS * kBytes > 2 * kPieceBytes ? NS * kBytes : 2 * kPieceBytes;   // the r update rides two 6 KB pieces
};
constexpr int kPersistMaxGroups = 1024;                               // active-group list kept in shared memory (uint16)
constexpr int kDescInts = 8;                                          // per row and direction: bs, cnt, col[0..5]
constexpr int kDescCols = kDescInts - 2;
constexpr uint32_t kRowBlock = 8;                                     // rows per block of the sweeps' work list (= warps per CTA)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Every wait inside the persistent kernels is bounded: a protocol error must end in a trap (a CUDA error the
// host reports), never in a hung GPU.  The bounds are seconds of spinning, far beyond any legitimate wait.
constexpr uint32_t kSpinLimit = 1u << 26;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        if (++spins > kSpinLimit) __trap();
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), streamed through L2
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int32_t ld_acquire(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int32_t* p, int32_t v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Row descriptors of the level path, built once per mesh: desc[dir][row] = {first block, block count,
// first six column indices} of the strictly upper (dir 0, backward sweep) / lower (dir 1, forward sweep)
// part of the row.  One 32-byte load per item replaces the rowptr/diag -> col chain.
__global__ void level_desc_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  const int32_t* __restrict__ diag, int64_t N, int32_t* __restrict__ desc,
                                  int32_t* __restrict__ max_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * N) return;
    const int dir = i >= N;
    const int64_t row = dir ? i - N : i;
    const int32_t bs = dir ? rowptr[row] : diag[row] + 1;
    const int32_t be = dir ? diag[row] : rowptr[row + 1];
    int32_t* d = desc + (size_t)i * kDescInts;
    d[0] = bs;
    d[1] = be - bs;
    for (int k = 0; k < kDescCols; ++k) d[2 + k] = bs + k < be ? col[bs + k] : -1;
    if (be - bs > 32) atomicMax(max_cnt, be - bs);
}

struct LevelArgs {
    const int32_t* desc;          // [2][N][8]
    const int32_t* col;
    mof_batch_dev B;
    int64_t N, nb;
    int ntiles;
    double omega, inv_omega;
    unsigned long long* probe;    // development probe (MOF_LEVEL_PROBE=1), else NULL: [0..15] summed SM cycles per item
                                  // segment and item counts, [16 + dir*N + row] = %globaltimer when group act[0] finished the row
};

// One warp, one sweep.  DIR 0: backward (rows descending, upper blocks), DIR 1: forward.
//   DIR 0 MODE 0: x += alpha p_old ; p <- zs r/omega + beta p_old ; t = (Dt+U)^-1 p        (vout = B.t)
//   DIR 0 MODE 1: vout = (Dt+U)^-1 (x + alpha p)                                           (back-transform)
//   DIR 1 MODE 0: w = (Dt+L)^-1 (p - ((2-omega)/omega) t) ; p'Ap += p.(t + w)              (vin = B.p, vout = B.ap)
//   DIR 1 MODE 1: vout = (Dt+L)^-1 vin
// act == nullptr: all groups (A = n_groups).
template <int DIR, int MODE, class CFG, int PROBE = 0>     // PROBE = 1 (MOF_LEVEL_PROBE=1): cycle counters per item segment, same results
__device__ __forceinline__ void level_sweep_phase(const LevelArgs& a, const double* vin, double* vout,
                                                  const uint16_t* act, int A, int32_t stamp, unsigned char* stage,
                                                  uint32_t bar, uint32_t& parity, uint64_t policy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t N = (uint32_t)a.N;
    // Work list: blocks of kRowBlock consecutive rows, inside a block group by group, inside a group row by
    // row -- the eight warps of a CTA take eight consecutive rows of ONE group, so a CTA's bulk copies walk
    // contiguous memory (matrix blocks and vector rows of neighbouring rows are neighbours in HBM).  Any order
    // that is monotone in the row within a group keeps "an item depends on earlier items only".
    const uint32_t nblk = (N + kRowBlock - 1) / kRowBlock;
    const uint32_t total = nblk * kRowBlock * (uint32_t)A;            // host guarantees (N + kRowBlock) * n_groups < 2^31
    const uint32_t W = gridDim.x * kWarps, uA = (uint32_t)A;
    uint32_t J = blockIdx.x * kWarps + warp;
    if (J >= total) return;
    const int32_t* __restrict__ desc = a.desc + (DIR ? (size_t)N * kDescInts : 0);
    const mof_batch_dev& B = a.B;
    const double omega = a.omega;
    constexpr int nvec = (DIR == 0) ? (MODE == 0 ? 3 : 2) : (MODE == 0 ? 2 : 1);
    constexpr int kStageBlocks = CFG::kBlocks, kStages = CFG::kStages, kStageBytes = CFG::kBytes, kStageValBytes = CFG::kValBytes;
    const uint32_t stage_s0 = smem_u32(stage);

    // item J -> (row, group); row >= N: padding of the last block (skipped)
    auto item = [&](uint32_t j, uint32_t& irow, uint32_t& ig) {
        const uint32_t blk = j / (kRowBlock * uA), rem = j - blk * (kRowBlock * uA);
        const uint32_t ia = rem / kRowBlock, q = blk * kRowBlock + (rem - ia * kRowBlock);
        irow = DIR ? q : nblk * kRowBlock - 1 - q;
        ig = act ? act[ia] : ia;
    };
    // next real item at or after j (j >= total: none)
    auto next_item = [&](uint32_t j, uint32_t& irow, uint32_t& ig) {
        while (j < total) {
            item(j, irow, ig);
            if (irow < N) break;
            j += W;
        }
        return j;
    };
    auto issue = [&](int slot, uint32_t irow, uint32_t ig, int32_t d) {
        const int32_t bs = __shfl_sync(kFull, d, 0), cnt = __shfl_sync(kFull, d, 1);
        if (lane == 0) {
            const uint32_t stage_s = stage_s0 + slot * kStageBytes, b = bar + 8u * slot;
            const int nstage = cnt < kStageBlocks ? cnt : kStageBlocks;
            const size_t vo = ((size_t)ig * N + irow) * 2 * MOF_W;
            mbar_expect_tx(b, (uint32_t)(nstage * 4 * MOF_W * 8 + nvec * kStageVecBytes));
            if (nstage > 0)
                bulk_g2s(stage_s, B.vals + ((size_t)ig * a.nb + bs) * 4 * MOF_W, (uint32_t)(nstage * 4 * MOF_W * 8), b, policy);
            if (DIR == 0) {
                bulk_g2s(stage_s + kStageValBytes, B.p + vo, kStageVecBytes, b, policy);
                bulk_g2s(stage_s + kStageValBytes + kStageVecBytes, B.x + vo, kStageVecBytes, b, policy);
                if (MODE == 0) bulk_g2s(stage_s + kStageValBytes + 2 * kStageVecBytes, B.r + vo, kStageVecBytes, b, policy);
            } else {
                bulk_g2s(stage_s + kStageValBytes, vin + vo, kStageVecBytes, b, policy);
                if (MODE == 0) bulk_g2s(stage_s + kStageValBytes + kStageVecBytes, B.t + vo, kStageVecBytes, b, policy);
            }
        }
    };
    // the warp's first kStages items go into flight: cur (being computed next), nx1 and -- with three stages -- nx2
    struct Item { uint32_t J, row, g; int32_t d; };
    auto fetch = [&](uint32_t j) {
        Item it{j, 0u, 0u, 0};
        it.J = next_item(j, it.row, it.g);
        if (it.J < total) it.d = __ldg(desc + (size_t)it.row * kDescInts + (lane & 7));
        return it;
    };
    Item cur = fetch(J);
    if (cur.J >= total) return;
    issue(0, cur.row, cur.g, cur.d);
    Item nx1 = fetch(cur.J + W), nx2{total, 0u, 0u, 0};
    if (nx1.J < total) issue(1, nx1.row, nx1.g, nx1.d);
    if (kStages > 2 && nx1.J < total) {
        nx2 = fetch(nx1.J + W);
        if (nx2.J < total) issue(2, nx2.row, nx2.g, nx2.d);
    }
    int slot = 0;
    uint32_t gprev = 0xffffffffu;
    // per-frame scalars of the item's group: re-read only when the group changes (with W a multiple of A, never)
    double s_alpha = 0.0, s_beta = 0.0, s_zsw = 0.0;
    int par = 0, fxk = 0;                      // fxk: binary scale of the group's p'Ap accumulator (forward, MODE 0)
    Fx128 fxacc = {0ull, 0ll};
    bool fxbad = false;
    long long probe_prev = PROBE == 1 ? clock64() : 0, pacc[5] = {0, 0, 0, 0, 0};

    for (;;) {
        const uint32_t row = cur.row, g = cur.g;
        const int32_t cd = cur.d;
        const bool more = nx1.J < total;
        // the item behind the ones in flight: its descriptor is fetched now, its data requested when this item's stage is free
        Item far{total, 0u, 0u, 0};
        {
            const uint32_t last = kStages > 2 ? nx2.J : nx1.J;
            if (last < total) far = fetch(last + W);
        }
        const uint32_t bslot = bar + 8u * slot;
        const double* sv = reinterpret_cast<const double*>(stage + slot * kStageBytes) + lane;
        const double* sx = reinterpret_cast<const double*>(stage + slot * kStageBytes + kStageValBytes) + lane;
        if (g != gprev) {                                // these loads overlap the gather below
            if (DIR == 1 && MODE == 0) {
                if (gprev != 0xffffffffu) fx_flush(B.scal, gprev, lane, fxacc, fxbad);
                fxk = (int)(long long)__ldcg(fx_word(B.scal, g, MOF_S_FX_K) + lane);
            }
            const int iters = __ldcg(state_ptr(B.state, g, MOF_I_ITERS) + lane);
            par = (MODE == 0 ? iters : iters + 1) & 1;
            if (DIR == 0) {
                s_alpha = __ldcg(scal_ptr(B.scal, g, MOF_S_ALPHA) + lane);
                if (MODE == 0) {
                    s_beta = __ldcg(scal_ptr(B.scal, g, MOF_S_BETA) + lane);
                    s_zsw = __ldcg(scal_ptr(B.scal, g, MOF_S_ZS) + lane) / omega;
                }
            }
            gprev = g;
        }
        const int32_t bs = __shfl_sync(kFull, cd, 0), cnt = __shfl_sync(kFull, cd, 1);
        double* v_l = vout + ((size_t)g * N) * 2 * MOF_W + lane;          // the sweep's own output, gathered from finished rows
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        if (PROBE == 1) tk0 = clock64();
        // MODE 0: the gathered entries carry their own validity bit (with_parity); MODE 1: ready stamps
        if (MODE == 1) {
            int32_t mycol = __shfl_sync(kFull, cd, 2 + (lane < kDescCols ? lane : 0));
            if (lane < cnt) {
                if (lane >= kDescCols) mycol = __ldg(a.col + bs + lane);
                const int32_t* f = B.ready + (size_t)g * N + mycol;
                uint32_t spins = 0;
                while (ld_acquire(f) < stamp)
                    if (++spins > kSpinLimit) __trap();
            }
            __syncwarp();
        }
        double gv[kStageBlocks][2];
        {
            uint32_t co[kStageBlocks];
#pragma unroll
            for (int k = 0; k < kStageBlocks; ++k) co[k] = (uint32_t)__shfl_sync(kFull, cd, 2 + k) * (2 * MOF_W);
            for (uint32_t spins = 0;; ++spins) {
                if (spins > kSpinLimit) __trap();
#pragma unroll
                for (int k = 0; k < kStageBlocks; ++k)
                    if (k < cnt) {
                        gv[k][0] = ld_strong(v_l + co[k]);
                        gv[k][1] = ld_strong(v_l + co[k] + MOF_W);
                    }
                if (MODE == 1) break;
                bool ok = true;
#pragma unroll
                for (int k = 0; k < kStageBlocks; ++k)
                    if (k < cnt) ok = ok && has_parity(gv[k][0], par) && has_parity(gv[k][1], par);
                if (__all_sync(kFull, ok)) break;
            }
        }
        if (PROBE == 1) tk1 = clock64();
        mbar_wait(bslot, (parity >> slot) & 1u);
        parity ^= 1u << slot;
        if (PROBE == 1) tk2 = clock64();
        v_l += (size_t)row * 2 * MOF_W;                                   // -> this row's entry of the output
        double a0, a1, q0 = 0.0, q1 = 0.0;                                // q: what the row's share of p'Ap needs (forward, MODE 0)
        if (DIR == 0) {
            const double p0 = sx[0], p1 = sx[MOF_W];
            const double y0 = fma(s_alpha, p0, sx[2 * MOF_W]), y1 = fma(s_alpha, p1, sx[3 * MOF_W]);   // pending step of the previous iteration
            if (MODE == 0) {
                const size_t vo = ((size_t)g * N + row) * 2 * MOF_W + lane;
                a0 = row_dir(s_zsw, sx[4 * MOF_W], s_beta, p0);
                a1 = row_dir(s_zsw, sx[5 * MOF_W], s_beta, p1);
                __stcs(B.x + vo, y0);
                __stcs(B.x + vo + MOF_W, y1);
                __stcs(B.p + vo, a0);
                __stcs(B.p + vo + MOF_W, a1);
            } else {
                a0 = y0;
                a1 = y1;
            }
        } else {
            a0 = sx[0];
            a1 = sx[MOF_W];
            if (MODE == 0) {
                const double kscale = (2.0 - omega) / omega;
                a0 = __fma_rn(-kscale, sx[2 * MOF_W], a0);
                a1 = __fma_rn(-kscale, sx[3 * MOF_W], a1);
            }
        }
#pragma unroll
        for (int k = 0; k < kStageBlocks; ++k)
            if (k < cnt)
                row_sub_block(a0, a1, sv[(4 * k) * MOF_W], sv[(4 * k + 1) * MOF_W], sv[(4 * k + 2) * MOF_W], sv[(4 * k + 3) * MOF_W],
                              gv[k][0], gv[k][1]);
        if (cnt > kStageBlocks) {                                        // rows with more blocks in this half (rare)
            const double* vals_l = B.vals + (size_t)g * a.nb * 4 * MOF_W + lane;
            const double* g_l = vout + ((size_t)g * N) * 2 * MOF_W + lane;
#pragma unroll 1
            for (int k = kStageBlocks; k < cnt; ++k) {
                const int32_t dk = __shfl_sync(kFull, cd, 2 + (k < kDescCols ? k : 0));
                const int64_t j = k < kDescCols ? dk : __ldg(a.col + bs + k);
                const double* ap = vals_l + (size_t)(bs + k) * 4 * MOF_W;
                double v0, v1;
                for (uint32_t spins = 0;; ++spins) {
                    if (spins > kSpinLimit) __trap();
                    v0 = ld_strong(g_l + (size_t)(2 * j) * MOF_W);
                    v1 = ld_strong(g_l + (size_t)(2 * j + 1) * MOF_W);
                    if (MODE == 1 || __all_sync(kFull, has_parity(v0, par) && has_parity(v1, par))) break;
                }
                row_sub_block(a0, a1, __ldcs(ap), __ldcs(ap + MOF_W), __ldcs(ap + 2 * MOF_W), __ldcs(ap + 3 * MOF_W), v0, v1);
            }
        }
        double o0 = omega * a0, o1 = omega * a1;
        if (MODE == 0 || DIR == 0) {
            o0 = with_parity(o0, par);
            o1 = with_parity(o1, par);
        }
        st_strong(v_l, o0);
        st_strong(v_l + MOF_W, o1);
        if (DIR == 1 && MODE == 0) {                                    // the row's share of p'Ap
            q0 = sx[2 * MOF_W] + o0;
            q1 = sx[3 * MOF_W] + o1;
            Fx128 x;
            if (!fx_from_double(row_pdot(sx[0], sx[MOF_W], q0, q1), fxk, x)) fxbad = true;
            fx_add(fxacc, x);
        }
        __syncwarp();                                                    // every lane is done with the stage
        if (MODE == 1) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            __syncwarp();
            if (lane == 0) st_relaxed(B.ready + (size_t)g * N + row, stamp);
        }
        if (PROBE == 1) tk3 = clock64();
        if (far.J < total) issue(slot, far.row, far.g, far.d);          // this stage is free again
        if (PROBE == 1) {
            pacc[0] += tk1 - tk0;                  // poll + gather
            pacc[1] += tk2 - tk1;                  // stage wait
            pacc[2] += tk3 - tk2;                  // arithmetic + result stores
            pacc[3] += tk0 - probe_prev;           // previous item's tail + loop head
            pacc[4] += 1;
            probe_prev = clock64();
            if (lane == 0 && act && g == act[0]) a.probe[16 + (size_t)DIR * N + row] = global_ns();
        }
        if (!more) break;
        cur = nx1;
        if (kStages > 2) { nx1 = nx2; nx2 = far; } else { nx1 = far; }
        slot = slot + 1 == kStages ? 0 : slot + 1;
    }
    if (DIR == 1 && MODE == 0 && gprev != 0xffffffffu) fx_flush(B.scal, gprev, lane, fxacc, fxbad);
    if (PROBE == 1 && lane == 0)
        for (int k = 0; k < 5; ++k) atomicAdd(a.probe + DIR * 8 + k, (unsigned long long)pacc[k]);
}

// Ordered list of the groups still iterating -> act[0..A) (shared memory); returns A to every thread.
__device__ __forceinline__ int build_active_list(const int32_t* group_done, int G, uint16_t* act, int* s_count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (warp == 0) {
        int n = 0;
        for (int g0 = 0; g0 < G; g0 += 32) {
            const int g = g0 + lane;
            const bool on = g < G && __ldcg(group_done + g) == 0;
            const unsigned m = __ballot_sync(kFull, on);
            if (on) act[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)g;
            n += __popc(m);
        }
        if (lane == 0) *s_count = n;
    }
    __syncthreads();
    return *s_count;
}

struct PersistShared {
    uint64_t bar[kWarps][3];               // [0]: sweeps and the r update, [0..2]: slots of the p'Ap phase
    uint16_t act[kPersistMaxGroups];
    int count;
    unsigned long long tprev, tacc[4];     // phase clock of CTA 0 (profile)
};

__device__ __forceinline__ void persist_setup(PersistShared& S) {
    if (threadIdx.x < kWarps * 3) mbar_init(smem_u32(&S.bar[threadIdx.x / 3][threadIdx.x % 3]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
}

// r -= alpha (t + w) ; r'r -> beta and the convergence test.  update_body<true>'s arithmetic; the three vectors
// of a warp's rows come in by bulk copies, four rows (3 x 2 KB) at a time, the next piece requested as soon as
// the stage has been read.
__device__ __noinline__ void level_update_phase(const LevelArgs& a, const uint16_t* act, int A, unsigned char* stage,
                                                   uint32_t bar, uint32_t& parity, uint64_t policy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const mof_batch_dev& B = a.B;
    const int64_t N = a.N;
    const int64_t items = (int64_t)a.ntiles * A;
    const uint32_t stage_s = smem_u32(stage);
    constexpr int kPiece = 4;                                             // rows per bulk step
    auto rows_of = [&](int64_t q, int h) {
        const int64_t row0 = (q / A) * MOF_TILE_ROWS + warp * kRowsPerWarp + h * kPiece;
        const int64_t n = N - row0;
        return (int)(n < 0 ? 0 : (n > kPiece ? kPiece : n));
    };
    // pieces of this CTA's items in order: piece i = half (i & 1) of the (i >> 1)-th item; piece i lives in slot
    // i % 2 of the warp's shared memory and piece i + 2 is requested as soon as piece i has been read
    constexpr int kStages = 2, kStageBytes = kPieceBytes;
    const int64_t q0 = blockIdx.x, gd = gridDim.x;
    const int64_t n_items = q0 < items ? (items - q0 + gd - 1) / gd : 0;
    const int64_t n_pieces = 2 * n_items;
    auto issue_piece = [&](int64_t i) {
        const int64_t q = q0 + (i >> 1) * gd;
        const int h = (int)(i & 1), slot = (int)(i % kStages);
        const int n = rows_of(q, h);
        if (lane == 0 && n > 0) {
            const int64_t row0 = (q / A) * MOF_TILE_ROWS + warp * kRowsPerWarp + h * kPiece;
            const size_t off = ((size_t)act[q % A] * N + row0) * 2 * MOF_W;
            const uint32_t bytes = (uint32_t)(n * 2 * MOF_W * 8);
            const uint32_t b = bar + 8u * slot, dst = stage_s + slot * kStageBytes;
            mbar_expect_tx(b, 3 * bytes);
            bulk_g2s(dst, B.ap + off, bytes, b, policy);
            bulk_g2s(dst + kSlotBytes, B.t + off, bytes, b, policy);
            bulk_g2s(dst + 2 * kSlotBytes, B.r + off, bytes, b, policy);
        }
    };
    for (int64_t i = 0; i < kStages && i < n_pieces; ++i) issue_piece(i);
    Fx128 fxacc = {0ull, 0ll};
    bool fxbad = false;
    int fxk = 0;
    int64_t gacc = -1;
    for (int64_t it = 0; it < n_items; ++it) {
        const int64_t q = q0 + it * gd;
        const int tile = (int)(q / A);
        const int64_t g = act[q % A];
        const double alpha = __ldcg(scal_ptr(B.scal, g, MOF_S_ALPHA) + lane);
        double rr = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = 2 * it + h;
            const int slot = (int)(i % kStages);
            const int n = rows_of(q, h);
            const double* sw = reinterpret_cast<const double*>(stage + slot * kStageBytes) + lane;
            const double* st = sw + kSlotBytes / 8;
            const double* sr = sw + 2 * kSlotBytes / 8;
            double rn[kPiece][2];
            if (n > 0) {
                mbar_wait(bar + 8u * slot, (parity >> slot) & 1u);
                parity ^= 1u << slot;
#pragma unroll
                for (int j = 0; j < kPiece; ++j)
                    if (j < n) {
                        const double a0 = sw[(2 * j) * MOF_W] + st[(2 * j) * MOF_W], a1 = sw[(2 * j + 1) * MOF_W] + st[(2 * j + 1) * MOF_W];
                        rn[j][0] = fma(-alpha, a0, sr[(2 * j) * MOF_W]);
                        rn[j][1] = fma(-alpha, a1, sr[(2 * j + 1) * MOF_W]);
                    }
            }
            __syncwarp();                                                 // the stage has been read
            if (i + kStages < n_pieces) issue_piece(i + kStages);
            if (n > 0) {
                double* r_l = B.r + ((size_t)g * N + (int64_t)tile * MOF_TILE_ROWS + warp * kRowsPerWarp + h * kPiece) * 2 * MOF_W + lane;
#pragma unroll
                for (int j = 0; j < kPiece; ++j)
                    if (j < n) {
                        r_l[(2 * j) * MOF_W] = rn[j][0];
                        r_l[(2 * j + 1) * MOF_W] = rn[j][1];
                        rr = fma(rn[j][0], rn[j][0], rr);
                        rr = fma(rn[j][1], rn[j][1], rr);
                    }
            }
        }
        // the warp's partial of r'r (its eight rows, in order) goes into the group's fixed-point accumulator
        if (g != gacc) {
            if (gacc >= 0) fx_flush(B.scal, gacc, lane, fxacc, fxbad);
            gacc = g;
            fxk = (int)(long long)__ldcg(fx_word(B.scal, g, MOF_S_FX_K_RR) + lane);
        }
        Fx128 x;
        if (!fx_from_double(rr, fxk, x)) fxbad = true;
        fx_add(fxacc, x);
    }
    if (gacc >= 0) fx_flush(B.scal, gacc, lane, fxacc, fxbad);
}

// After a grid barrier: one warp per group turns the fixed-point accumulator into the scalar step
// (STEP 0: p'Ap -> alpha; STEP 1: r'r -> beta and the convergence test).
template <int STEP>
__device__ __noinline__ void level_group_phase(const LevelArgs& a, const uint16_t* act, int A) {
    const mof_batch_dev& B = a.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp != 0) return;
    for (int q = blockIdx.x; q < A; q += gridDim.x) {
        const int64_t g = act[q];
        if constexpr (STEP == 0) fx_finalize_alpha(B.scal, B.state, g, B.n_groups, lane);
        else fx_finalize_beta(B, g, lane, a.inv_omega);
    }
}

// grid barrier + make what other SMs wrote with ordinary stores visible to this thread's bulk copies
__device__ __forceinline__ void phase_barrier(cooperative_groups::grid_group& grid) {
    asm volatile("fence.proxy.async;" ::: "memory");
    grid.sync();
    asm volatile("fence.proxy.async;" ::: "memory");
}

// n_iter PCG iterations (or fewer if every group converges earlier).  stamp0: value of the last stamp used
// in B.ready; the sweeps of iteration i use stamp0 + 2 i + 1 and stamp0 + 2 i + 2.
template <int PROBE, class CFG>   // two CTAs of eight warps per SM, <= 128 registers; CFG: stage configuration of the sweeps
__global__ void __launch_bounds__(256, 2) level_iter_kernel(LevelArgs a, int n_iter, int32_t stamp0, int timing) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ PersistShared S;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int warp = threadIdx.x >> 5;
    const mof_batch_dev& B = a.B;
    const int G = B.n_groups;
    persist_setup(S);
    unsigned char* stage = dyn_smem + (size_t)warp * CFG::kWarpBytes;
    const uint32_t bar = smem_u32(&S.bar[warp][0]);
    uint32_t parity = 0;                   // bit s: phase parity of this warp's mbarrier s (bit 0 is shared by all phases)
    const uint64_t policy = policy_evict_first();
    const bool clock = timing && blockIdx.x == 0 && threadIdx.x == 0;
    if (clock) {
        S.tprev = global_ns();
        for (int k = 0; k < 4; ++k) S.tacc[k] = 0;
    }
    auto lap = [&](int k) {
        if (clock) { const unsigned long long t = global_ns(); S.tacc[k] += t - S.tprev; S.tprev = t; }
    };
    for (int it = 0; it < n_iter; ++it) {
        const int A = build_active_list(group_done_ptr(B.state, G), G, S.act, &S.count);
        if (A == 0) break;
        level_sweep_phase<0, 0, CFG, PROBE>(a, nullptr, B.t, S.act, A, stamp0 + 2 * it + 1, stage, bar, parity, policy);
        phase_barrier(grid);
        lap(0);
        level_sweep_phase<1, 0, CFG, PROBE>(a, B.p, B.ap, S.act, A, stamp0 + 2 * it + 2, stage, bar, parity, policy);
        phase_barrier(grid);
        lap(1);
        level_group_phase<0>(a, S.act, A);                               // p'Ap accumulator -> alpha
        phase_barrier(grid);
        lap(2);
        level_update_phase(a, S.act, A, stage, bar, parity, policy);
        phase_barrier(grid);
        level_group_phase<1>(a, S.act, A);                               // r'r accumulator -> beta, convergence
        phase_barrier(grid);
        lap(3);
    }
    if (clock) {                                   // accumulated phase times (ns) of this solve: profile only
        double* acc = scal_ptr(B.scal, 0, MOF_S_SPARE);
        for (int k = 0; k < 4; ++k) acc[k] += (double)S.tacc[k];
    }
}

// A single sweep of all groups (start: r = (Dt+L)^-1 b ; end: xs = (Dt+U)^-1 (xhat + alpha p)).
template <int DIR, class CFG>
__global__ void __launch_bounds__(256, 2) level_sweep_kernel(LevelArgs a, const double* vin, double* vout, int32_t stamp) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ PersistShared S;
    const int warp = threadIdx.x >> 5;
    persist_setup(S);
    uint32_t parity = 0;
    level_sweep_phase<DIR, 1, CFG>(a, vin, vout, nullptr, a.B.n_groups, stamp, dyn_smem + (size_t)warp * CFG::kWarpBytes,
                              smem_u32(&S.bar[warp][0]), parity, policy_evict_first());
}

// ---------------------------------------------------------------------------------
// Start / verification kernel.  ssor != 0: the batch holds the scaled system (vals = S A S,
// rhs = S b, minv = S); norms of the ORIGINAL residual / rhs are obtained by applying S^-1.
//   MODE_START_JACOBI : x = 0, r = rhs, z = D^-1 r, p = 0 ; ||b||^2 ; per-frame bookkeeping
//   MODE_NORM         : ||S^-1 rhs||^2 = ||b||^2 -> BBT (SSOR path)
//   MODE_START_SSOR   : r (= (Dt+L)^-1 rhs, already in B.r) ; x = 0, p = 0 ; bookkeeping
//   MODE_CALIBRATE    : (SSOR, once, after the first check interval) true residual of the frames still iterating ->
//                       their threshold on the recurrence residual is set from the measured true / recurrence
//                       ratio, so that a frame freezes where its TRUE residual meets tol instead of freezing early
//                       (smooth signals: ratio 2.4, i.e. a failed verification, a resume and a second verification)
//                       or late (wrapped phases: ratio 0.13, i.e. ~0.7 digits = ~8 iterations too many)
//   MODE_VERIFY       : true residual ||b - A x||^2 from ap = A x (scaled space for SSOR) ; frames
//                       frozen on the recurrence residual that miss tol get a tighter threshold and resume
// ---------------------------------------------------------------------------------
enum { MODE_START_JACOBI = 0, MODE_NORM = 1, MODE_START_SSOR = 2, MODE_VERIFY = 3, MODE_CALIBRATE = 4 };

// fill_parity (level path, MODE_START_SSOR): t and w start with the validity bit of "one iteration ago".
// ax: A x of MODE_VERIFY.
__global__ void __launch_bounds__(256) init_kernel(mof_batch_dev B, int64_t N, int ntiles, int mode, double tol2,
                                                   int last_round, int ssor, double inv_omega, int fill_parity,
                                                   const double* __restrict__ ax, double margin = 0.5) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x;
    const int64_t row0 = (int64_t)tile * MOF_TILE_ROWS + warp * kRowsPerWarp;
    double rz = 0.0, rr = 0.0;
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        const size_t im = mof_ix_minv(N, g, v, 0) + lane;
        double r0, r1;
        if (mode == MODE_START_SSOR) { r0 = B.r[i0]; r1 = B.r[i1]; }
        else                         { r0 = B.rhs[i0]; r1 = B.rhs[i1]; }
        if (mode == MODE_VERIFY || mode == MODE_CALIBRATE) { r0 -= ax[i0]; r1 -= ax[i1]; }
        if (ssor && (mode == MODE_VERIFY || mode == MODE_NORM || mode == MODE_CALIBRATE)) {      // back to the unscaled system: S^-1 r
            const double s0 = B.minv[im], s1 = B.minv[im + MOF_W], s2 = B.minv[im + 2 * MOF_W];
            const double rdet = 1.0 / (s0 * s2 - s1 * s1);
            const double u0 = (s2 * r0 - s1 * r1) * rdet, u1 = (s0 * r1 - s1 * r0) * rdet;
            r0 = u0; r1 = u1;
        }
        if (mode == MODE_START_JACOBI) {
            const double m0 = B.minv[im], m1 = B.minv[im + MOF_W], m2 = B.minv[im + 2 * MOF_W];
            const double z0 = m0 * r0 + m1 * r1;
            const double z1 = m1 * r0 + m2 * r1;
            B.r[i0] = r0; B.r[i1] = r1;
            B.z[i0] = z0; B.z[i1] = z1;
            rz = fma(r0, z0, rz); rz = fma(r1, z1, rz);
        }
        if (mode == MODE_START_JACOBI || mode == MODE_START_SSOR) {
            B.x[i0] = 0.0; B.x[i1] = 0.0;
            B.p[i0] = 0.0; B.p[i1] = 0.0;
        }
        if (fill_parity) {
            const double stale = with_parity(0.0, 1);
            B.t[i0] = stale; B.t[i1] = stale;
            B.ap[i0] = stale; B.ap[i1] = stale;
        }
        rr = fma(r0, r0, rr); rr = fma(r1, r1, rr);
    }
    if (mode == MODE_START_SSOR) rz = rr * inv_omega;                   // z = r / omega
    double val[2] = {rz, rr}, tot[2];
    if (!tile_reduce<2>(val, B.partial + (size_t)g * ntiles * 2 * MOF_W, ntiles, tile, ticket_ptr(B.state, G) + g, tot))
        return;
    if (warp != 0) return;
    int32_t* active = state_ptr(B.state, g, MOF_I_ACTIVE);
    int32_t* status = state_ptr(B.state, g, MOF_I_STATUS);
    if (mode == MODE_NORM) {
        scal_ptr(B.scal, g, MOF_S_BBT)[lane] = tot[1];
        return;
    }
    if (mode == MODE_CALIBRATE) {
        // q = (true residual / ||b||)^2 / (recurrence residual / its start)^2, measured now; the frame shall freeze
        // when recurrence <= tol^2 / q, with the factor 0.5 of the verification's rescaling as margin.  A function of
        // the frame's own numbers only: results stay independent of batching.
        if (active[lane]) {
            const double bbt = scal_ptr(B.scal, g, MOF_S_BBT)[lane], bb = scal_ptr(B.scal, g, MOF_S_BB)[lane];
            const double rr = scal_ptr(B.scal, g, MOF_S_RR)[lane];
            const double q = (tot[1] / bbt) / (rr / bb);
            if (isfinite(q) && q > 1e-6 && q < 1e6) scal_ptr(B.scal, g, MOF_S_THR)[lane] = tol2 / q * margin;
        }
        return;
    }
    int act;
    const int was = (mode == MODE_VERIFY) ? active[lane] : 0;
    if (mode == MODE_VERIFY) {
        const double bbt = scal_ptr(B.scal, g, MOF_S_BBT)[lane];
        scal_ptr(B.scal, g, MOF_S_RRTRUE)[lane] = tot[1];
        act = was;
        if (status[lane] == MOF_STATUS_CONVERGED && !(tot[1] <= tol2 * bbt)) {
            if (last_round || !isfinite(tot[1])) {
                status[lane] = MOF_STATUS_MAXITER;          // still short of tol after the allowed rounds
            } else {
                // true / recurrence residual ratio is stable along a run: rescale the threshold
                double* thr = scal_ptr(B.scal, g, MOF_S_THR) + lane;
                *thr = *thr * (tol2 * bbt / tot[1]) * 0.5;
                scal_ptr(B.scal, g, MOF_S_BETA)[lane] = scal_ptr(B.scal, g, MOF_S_BETA_SAVED)[lane];
                scal_ptr(B.scal, g, MOF_S_ZS)[lane] = 1.0;
                status[lane] = MOF_STATUS_PENDING;
                act = 1;
            }
        }
    } else {
        const bool valid = g * MOF_W + lane < B.n_frames;
        if (mode == MODE_START_JACOBI) scal_ptr(B.scal, g, MOF_S_BBT)[lane] = tot[1];
        scal_ptr(B.scal, g, MOF_S_BB)[lane] = tot[1];
        scal_ptr(B.scal, g, MOF_S_RRTRUE)[lane] = scal_ptr(B.scal, g, MOF_S_BBT)[lane];
        scal_ptr(B.scal, g, MOF_S_RZ)[lane] = tot[0];
        scal_ptr(B.scal, g, MOF_S_RR)[lane] = tot[1];
        scal_ptr(B.scal, g, MOF_S_THR)[lane] = tol2;
        scal_ptr(B.scal, g, MOF_S_ALPHA)[lane] = 0.0;
        scal_ptr(B.scal, g, MOF_S_BETA)[lane] = 0.0;      // first p-update: p = z
        scal_ptr(B.scal, g, MOF_S_ZS)[lane] = 1.0;
        scal_ptr(B.scal, g, MOF_S_BETA_SAVED)[lane] = 0.0;
        if (mode == MODE_START_SSOR) {                       // level path: p'Ap accumulator, scaled by r'z for the first iteration
            fx_word(B.scal, g, MOF_S_FX_LO)[lane] = 0ull;
            fx_word(B.scal, g, MOF_S_FX_HI)[lane] = 0ull;
            fx_word(B.scal, g, MOF_S_FX_BAD)[lane] = 0ull;
            fx_word(B.scal, g, MOF_S_FX_K)[lane] = (unsigned long long)(long long)fx_scale_for(tot[0]);
            fx_word(B.scal, g, MOF_S_FX_K_RR)[lane] = (unsigned long long)(long long)fx_scale_for(tot[1]);
        }
        state_ptr(B.state, g, MOF_I_ITERS)[lane] = 0;
        if (!valid || tot[1] == 0.0) { act = 0; status[lane] = MOF_STATUS_ZERO_RHS; }
        else if (!isfinite(tot[1]) || !isfinite(tot[0])) { act = 0; status[lane] = MOF_STATUS_BREAKDOWN; }
        else { act = 1; status[lane] = MOF_STATUS_PENDING; }
    }
    active[lane] = act;
    const int any = __any_sync(kFull, act);
    const int gained = __popc(__ballot_sync(kFull, act && !was));
    if (lane == 0) {
        if (gained) atomicAdd(lanes_active_ptr(B.state, G), gained);
        int32_t* done = group_done_ptr(B.state, G) + g;
        if (mode != MODE_VERIFY) {
            *done = any ? 0 : 1;
            if (any) atomicAdd(groups_active_ptr(B.state, G), 1);
        } else if (any && *done) {
            *done = 0;
            atomicAdd(groups_active_ptr(B.state, G), 1);
        }
    }
}

// Level path, after a verification that makes frames resume: a resuming frame's counter did not move while
// it was frozen, so its stale w entries (and, had the back-transform not rewritten them, t) carry exactly the
// parity its next sweep will wait for.  Give every entry of the resuming frames the opposite one; their old
// values are never used as data again.
__global__ void __launch_bounds__(256) restamp_kernel(mof_batch_dev B, int64_t N) {
    const int64_t g = blockIdx.y;
    const int G = B.n_groups;
    if (group_done_ptr(B.state, G)[g]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!state_ptr(B.state, g, MOF_I_ACTIVE)[lane]) return;
    const int stale = (state_ptr(B.state, g, MOF_I_ITERS)[lane] + 1) & 1;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        B.t[i0] = with_parity(B.t[i0], stale);
        B.t[i1] = with_parity(B.t[i1], stale);
        B.ap[i0] = with_parity(B.ap[i0], stale);
        B.ap[i1] = with_parity(B.ap[i1], stale);
    }
}

// SSOR path, end of the solve: x = S xs with xs (solution of the scaled system) in B.t
__global__ void __launch_bounds__(256) unscale_kernel(mof_batch_dev B, int64_t N) {
    const int64_t g = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    for (int q = 0; q < kRowsPerWarp; ++q) {
        const int64_t v = row0 + q;
        if (v >= N) break;
        const size_t i0 = mof_ix_vec(N, g, v, 0) + lane, i1 = i0 + MOF_W;
        const size_t im = mof_ix_minv(N, g, v, 0) + lane;
        const double s0 = B.minv[im], s1 = B.minv[im + MOF_W], s2 = B.minv[im + 2 * MOF_W];
        const double t0 = B.t[i0], t1 = B.t[i1];
        // a frame with f = 0 has V = 0 exactly (spsolve(a, 0) = 0); its t may carry the validity bit of the level path
        const bool zero = state_ptr(B.state, g, MOF_I_STATUS)[lane] == MOF_STATUS_ZERO_RHS;
        B.x[i0] = zero ? 0.0 : s0 * t0 + s1 * t1;
        B.x[i1] = zero ? 0.0 : s1 * t0 + s2 * t1;
    }
}

// x [G][N][2][32] -> V[k][perm[v] + N*alpha]   (reference layout, cof:149)
__global__ void __launch_bounds__(256) unpack_kernel(int64_t N, int32_t n_frames, const int32_t* __restrict__ perm,
                                                     const double* __restrict__ x, double* __restrict__ V, int64_t ld) {
    __shared__ double s[2][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t g = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    for (int vv = ty; vv < 32; vv += 8) {
        int64_t v = v0 + vv;
        if (v < N) {
            s[0][vv][tx] = x[mof_ix_vec(N, g, v, 0) + tx];
            s[1][vv][tx] = x[mof_ix_vec(N, g, v, 1) + tx];
        }
    }
    __syncthreads();
    const int64_t v = v0 + tx;
    if (v >= N) return;
    const int64_t o = perm[v];
    for (int fr = ty; fr < 32; fr += 8) {
        int64_t k = g * 32 + fr;
        if (k < n_frames) {
            V[k * ld + o] = s[0][tx][fr];
            V[k * ld + N + o] = s[1][tx][fr];
        }
    }
}

int check_batch(const mof_mesh_dev* mesh, const mof_batch_dev* b) {
    if (!mesh || !b) return mof_set_error(-1, "NULL mesh/batch");
    if (b->n_groups <= 0 || b->n_groups > 65535) return mof_set_error(-1, "n_groups out of range (1..65535)");
    if (b->n_frames < 0 || b->n_frames > b->n_groups * MOF_GROUP) return mof_set_error(-1, "n_frames does not fit n_groups");
    if (!mesh->rowptr || !mesh->col) return mof_set_error(-1, "mesh pattern missing");
    return 0;
}

}  // namespace

namespace {
thread_local int32_t g_last_path[4] = {0, 0, 0, 0};   // path, grid of the iteration kernel, CTAs per SM, reason of a fallback
}
extern "C" int mof_pcg_last_path(int32_t* info) {
    if (info) for (int k = 0; k < 4; ++k) info[k] = g_last_path[k];
    return g_last_path[0];
}

extern "C" int64_t mof_state_ints(int32_t n_groups) {
    return (int64_t)n_groups * MOF_I_COUNT * MOF_W + 2 * (int64_t)n_groups + 2;
}

extern "C" int mof_level_desc_build(const mof_mesh_dev* mesh, int32_t* desc, void* stream) {
    MOF_REQUIRE(mesh && desc && mesh->rowptr && mesh->col && mesh->diag, "NULL argument");
    cudaStream_t st = mof_stream(stream);
    const int64_t N = mesh->n_vertices;
    int32_t* d_max = nullptr;
    MOF_CUDA_TRY(cudaMalloc(&d_max, sizeof(int32_t)));
    cudaError_t e = cudaMemsetAsync(d_max, 0, sizeof(int32_t), st);
    if (e == cudaSuccess) {
        level_desc_kernel<<<mof_cdiv(2 * N, 256), 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, N, desc, d_max);
        e = cudaGetLastError();
    }
    int32_t h_max = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_max, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_max);
    if (e != cudaSuccess) return mof_set_error(-100 - (int)e, "mof_level_desc_build failed: %s", cudaGetErrorString(e));
    if (h_max > 32) return mof_set_error(1, "mof_level_desc_build: a row has %d blocks on one side of its diagonal (limit 32)", h_max);
    return 0;
}

extern "C" int mof_spmv_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, const double* x, double* y,
                              void* stream) {
    if (int rc = check_batch(mesh, batch)) return rc;
    MOF_REQUIRE(x && y && batch->vals, "NULL argument");
    const int ntiles = (int)mof_num_tiles(mesh->n_vertices);
    dim3 grid(ntiles, batch->n_groups);
    spmv_kernel<false><<<grid, 256, 0, mof_stream(stream)>>>(mesh->rowptr, mesh->col, batch->vals, x, y,
                                                            mesh->n_vertices, mesh->n_blocks, ntiles, nullptr,
                                                            nullptr, nullptr, batch->n_groups);
    MOF_LAUNCH_CHECK("spmv_kernel<false>");
    return 0;
}

extern "C" int mof_unpack_solution(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double* V, int64_t ld,
                                   void* stream) {
    if (int rc = check_batch(mesh, batch)) return rc;
    MOF_REQUIRE(V && batch->x && ld >= 2 * mesh->n_vertices, "bad V / ld");
    dim3 grid(mof_cdiv(mesh->n_vertices, 32), batch->n_groups), block(32, 8);
    unpack_kernel<<<grid, block, 0, mof_stream(stream)>>>(mesh->n_vertices, batch->n_frames, mesh->perm, batch->x, V, ld);
    MOF_LAUNCH_CHECK("unpack_kernel");
    return 0;
}

extern "C" int mof_pcg_solve_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double tol, double omega,
                                   int32_t max_iter, int32_t check_every, int32_t max_restarts, int32_t* iters,
                                   double* relres, int32_t* status, mof_pcg_profile* prof, void* stream) {
    if (int rc = check_batch(mesh, batch)) return rc;
    const mof_batch_dev& B = *batch;
    MOF_REQUIRE(B.vals && B.rhs && B.minv && B.x && B.r && B.z && B.p && B.ap && B.partial && B.scal && B.state,
                "batch buffers missing");
    MOF_REQUIRE(tol > 0 && max_iter >= 0, "bad tol / max_iter");
    MOF_REQUIRE(omega >= 0.0 && omega < 2.0, "omega must be 0 (block Jacobi) or in (0,2) (SSOR)");
    const bool ssor = omega > 0.0;
    const int C = mesh->n_colors;
    const int L = mesh->n_levels;
    const bool levels = ssor && L > 0;
    const int32_t* lp = mesh->level_ptr;
    if (levels) {
        MOF_REQUIRE(lp && lp[0] == 0 && lp[L] == mesh->n_vertices, "level row ranges do not cover the mesh");
        MOF_REQUIRE(B.t && mesh->diag, "SSOR needs batch->t and mesh->diag");
    } else if (ssor) {
        MOF_REQUIRE(C > 0 && C <= MOF_MAX_COLORS,
                    "SSOR needs a mesh built with the block-multicolour (reorder = 2) or level-scheduled (reorder = 3) ordering");
        MOF_REQUIRE(B.t && mesh->diag, "SSOR needs batch->t and mesh->diag");
        MOF_REQUIRE(mesh->color_tile_ptr[0] == 0 && mesh->color_tile_ptr[C] == mof_num_tiles(mesh->n_vertices),
                    "colour tile ranges do not cover the mesh");
    }
    if (check_every <= 0) check_every = 32;
    if (max_restarts < 0) max_restarts = 0;
    cudaStream_t st = mof_stream(stream);
    const int64_t N = mesh->n_vertices, nb = mesh->n_blocks;
    const int G = B.n_groups;
    const int ntiles = (int)mof_num_tiles(N);
    const double tol2 = tol * tol;
    const double inv_omega = ssor ? 1.0 / omega : 0.0;
    dim3 grid(ntiles, G);
    int32_t* d_active_groups = B.state + (size_t)G * MOF_I_COUNT * MOF_W + 2 * (size_t)G;
    int64_t launches = 0;


    // Persistent level kernels (default of the level path): cooperative launches sized to the device.
    const char* persist_env = getenv("MOF_LEVEL_PERSIST");
    bool persist = levels && mesh->level_desc && B.ready && G <= kPersistMaxGroups &&
                   ((double)N + kRowBlock) * (double)G < 2147483648.0 - 65536.0 && !(persist_env && persist_env[0] == '0');
    g_last_path[3] = persist ? 0 : (!levels ? 0 : !mesh->level_desc ? 1 : !B.ready ? 2 : G > kPersistMaxGroups ? 3 : 4);
    int grid_iter = 0, grid_sweep = 0;
    const void* iter_fn = nullptr;
    using CfgWide = StageCfg<4, 2>;       // four blocks per row staged, two stages per warp
    using CfgDeep = StageCfg<3, 3>;       // three blocks staged, three stages per warp (regular meshes)
    const bool deep = mesh->level_stage_blocks == 3;
    const size_t persist_smem = (size_t)kWarps * (deep ? CfgDeep::kWarpBytes : CfgWide::kWarpBytes);
    const void* sweep_fn[2] = {deep ? (const void*)level_sweep_kernel<0, CfgDeep> : (const void*)level_sweep_kernel<0, CfgWide>,
                               deep ? (const void*)level_sweep_kernel<1, CfgDeep> : (const void*)level_sweep_kernel<1, CfgWide>};
    LevelArgs largs{mesh->level_desc, mesh->col, B, N, nb, ntiles, omega, inv_omega, nullptr};
    unsigned long long* probe_buf = nullptr;
    struct ProbeGuard {
        unsigned long long*& p;
        ~ProbeGuard() { if (p) cudaFree(p); }
    } probe_guard{probe_buf};
    int32_t stamp = 0;                   // last stamp written to B.ready by a sweep of this solve
    if (persist) {
        int dev = 0, coop = 0, sms = 0, occ_iter = 0, occ_f = 0, occ_b = 0;
        MOF_CUDA_TRY(cudaGetDevice(&dev));
        MOF_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        MOF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const char* probe_env = getenv("MOF_LEVEL_PROBE");           // development probe: per-item cycle breakdown on stderr
        const bool probe = probe_env && probe_env[0] == '1';
        iter_fn = deep ? (probe ? (const void*)level_iter_kernel<1, CfgDeep> : (const void*)level_iter_kernel<0, CfgDeep>)
                       : (probe ? (const void*)level_iter_kernel<1, CfgWide> : (const void*)level_iter_kernel<0, CfgWide>);
        if (probe) {
            const size_t pb = (16 + 2 * (size_t)N) * sizeof(unsigned long long);
            MOF_CUDA_TRY(cudaMalloc(&probe_buf, pb));
            MOF_CUDA_TRY(cudaMemsetAsync(probe_buf, 0, pb, st));
            largs.probe = probe_buf;
        }
        MOF_CUDA_TRY(cudaFuncSetAttribute(iter_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        MOF_CUDA_TRY(cudaFuncSetAttribute(sweep_fn[0], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        MOF_CUDA_TRY(cudaFuncSetAttribute(sweep_fn[1], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        // static + dynamic shared memory exceeds 48 KB: opt in
        MOF_CUDA_TRY(cudaFuncSetAttribute(iter_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)persist_smem));
        MOF_CUDA_TRY(cudaFuncSetAttribute(sweep_fn[0], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)persist_smem));
        MOF_CUDA_TRY(cudaFuncSetAttribute(sweep_fn[1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)persist_smem));
        MOF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_iter, iter_fn, 256, persist_smem));
        MOF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, sweep_fn[0], 256, persist_smem));
        MOF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, sweep_fn[1], 256, persist_smem));
        const int occ_s = occ_b < occ_f ? occ_b : occ_f;
        if (!coop || occ_iter < 1 || occ_s < 1) {
            persist = false;             // no co-residency guarantee: per-level launches
            g_last_path[3] = !coop ? 5 : 6;
        } else {
            g_last_path[2] = occ_iter;
            const char* cta_env = getenv("MOF_LEVEL_CTAS_PER_SM");   // tuning knob (1..occupancy)
            int want = cta_env ? atoi(cta_env) : 0;
            grid_iter = sms * ((want >= 1 && want < occ_iter) ? want : occ_iter);
            grid_sweep = sms * ((want >= 1 && want < occ_s) ? want : occ_s);
            // a grid that is a multiple of the group count keeps every warp on one group from item to item
            // (its per-frame scalars stay in registers) while all groups are active
            if (2 * G <= grid_iter) grid_iter = grid_iter / G * G;
            if (2 * G <= grid_sweep) grid_sweep = grid_sweep / G * G;
            MOF_CUDA_TRY(cudaMemsetAsync(B.ready, 0, (size_t)G * N * sizeof(int32_t), st));
        }
    }
    g_last_path[0] = !ssor ? MOF_PATH_JACOBI : !levels ? MOF_PATH_MULTICOLOUR : persist ? MOF_PATH_LEVEL_PERSISTENT : MOF_PATH_LEVEL_LAUNCHES;
    g_last_path[1] = grid_iter;
    auto persist_sweep = [&](int dir, const double* vin, double* vout, cudaStream_t st) -> cudaError_t {
        ++stamp;
        void* args[] = {(void*)&largs, (void*)&vin, (void*)&vout, (void*)&stamp};
        return cudaLaunchCooperativeKernel(sweep_fn[dir ? 1 : 0],
                                           dim3(grid_sweep), dim3(256), args, persist_smem, st);
    };
    cudaError_t persist_err = cudaSuccess;
    auto sweep_back = [&](int mode, double* tout, cudaStream_t st) {
        if (levels && persist && mode == 1) {
            const cudaError_t e = persist_sweep(0, nullptr, tout, st);
            if (e != cudaSuccess) persist_err = e;
            ++launches;
            return;
        }
        if (levels) {
            for (int l = L - 1; l >= 0; --l) {
                const int r0 = lp[l], r1 = lp[l + 1];
                if (r1 <= r0) continue;
                dim3 gs(mof_cdiv(r1 - r0, kWarps), G);
                if (mode == 0) level_back_kernel<0><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, tout, N, nb, r0, r1, omega);
                else           level_back_kernel<1><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, tout, N, nb, r0, r1, omega);
                ++launches;
            }
            return;
        }
        for (int c = C - 1; c >= 0; --c) {
            const int t0 = mesh->color_tile_ptr[c], t1 = mesh->color_tile_ptr[c + 1];
            if (t1 <= t0) continue;
            dim3 gs(mof_cdiv(t1 - t0, kWarps), G);
            if (mode == 0) sweep_back_kernel<0><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, tout, N, nb, t0, t1, omega);
            else           sweep_back_kernel<1><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, tout, N, nb, t0, t1, omega);
            ++launches;
        }
    };
    auto sweep_fwd = [&](int mode, const double* pin, double* wout, cudaStream_t st) {
        if (levels && persist && mode == 1) {
            const cudaError_t e = persist_sweep(1, pin, wout, st);
            if (e != cudaSuccess) persist_err = e;
            ++launches;
            return;
        }
        if (levels) {
            for (int l = 0; l < L; ++l) {
                const int r0 = lp[l], r1 = lp[l + 1];
                if (r1 <= r0) continue;
                dim3 gs(mof_cdiv(r1 - r0, kWarps), G);
                if (mode == 0) level_fwd_kernel<0><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, pin, wout, N, nb, r0, r1, omega);
                else           level_fwd_kernel<1><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, pin, wout, N, nb, r0, r1, omega);
                ++launches;
            }
            if (mode == 0) {
                level_alpha_kernel<<<G, 32, 0, st>>>(B);
                ++launches;
            }
            return;
        }
        for (int c = 0; c < C; ++c) {
            const int t0 = mesh->color_tile_ptr[c], t1 = mesh->color_tile_ptr[c + 1];
            if (t1 <= t0) continue;
            dim3 gs(mof_cdiv(t1 - t0, kWarps), G);
            if (mode == 0) sweep_fwd_kernel<0><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, pin, wout, N, nb, t0, t1, ntiles, omega);
            else           sweep_fwd_kernel<1><<<gs, 256, 0, st>>>(mesh->rowptr, mesh->col, mesh->diag, B, pin, wout, N, nb, t0, t1, ntiles, omega);
            ++launches;
        }
    };

    // group_done, tickets, groups_active, lanes_active <- 0
    MOF_CUDA_TRY(cudaMemsetAsync(B.state + (size_t)G * MOF_I_COUNT * MOF_W, 0, (2 * (size_t)G + 2) * sizeof(int32_t), st));
    if (!ssor) {
        init_kernel<<<grid, 256, 0, st>>>(B, N, ntiles, MODE_START_JACOBI, tol2, 0, 0, 0.0, 0, nullptr);
        launches += 1;
    } else {
        init_kernel<<<grid, 256, 0, st>>>(B, N, ntiles, MODE_NORM, tol2, 0, 1, inv_omega, 0, nullptr);
        sweep_fwd(1, B.rhs, B.r, st);                               // r = (Dt+L)^-1 b
        init_kernel<<<grid, 256, 0, st>>>(B, N, ntiles, MODE_START_SSOR, tol2, 0, 1, inv_omega, levels ? 1 : 0, nullptr);
        launches += 2;
    }
    MOF_LAUNCH_CHECK("pcg start kernels");

    // optional sampled per-kernel timing (one iteration per check interval) for the roofline report
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (prof)
        for (int q = 0; q < 6; ++q) MOF_CUDA_TRY(cudaEventCreate(&ev[q]));
    struct EventGuard {
        cudaEvent_t* e;
        ~EventGuard() { for (int q = 0; q < 6; ++q) if (e[q]) cudaEventDestroy(e[q]); }
    } guard{ev};
    if (prof && persist)                 // in-kernel phase clocks accumulate in scal[0][SPARE][0..3]
        MOF_CUDA_TRY(cudaMemsetAsync(scal_host_ptr(B.scal, 0, MOF_S_SPARE), 0, 4 * sizeof(double), st));

    int32_t h_act[2] = {0, 0};          // groups, frames still iterating
    int32_t& h_active = h_act[0];
    MOF_CUDA_TRY(cudaMemcpyAsync(h_act, d_active_groups, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MOF_CUDA_TRY(cudaStreamSynchronize(st));
    int it = 0, rounds = 0;
    const char* calib_env = getenv("MOF_CALIBRATE");
    const bool calibrate = !(calib_env && calib_env[0] == '0');       // MOF_CALIBRATE=0: round-1 behaviour (tests compare both)
    bool calibrated = false;
    // margin on the squared residual: the ratio measured after the first interval drifts by up to ~2x until convergence
    // (wrapped-phase input at 328k vertices); 0.35 kept every frame of C2 and C4 inside one verification round, 0.5 did not
    const double calib_margin = 0.35;
    double* xphys = ssor ? B.t : B.x;   // where the solution of A x = b lives at verification time

    // Level path: one iteration is ~2 L small launches with identical arguments every time (alpha, beta
    // live in device memory), so it is captured once into a CUDA graph and replayed.  Capture happens on
    // an internal stream (the caller's may be the legacy default stream, which cannot capture).
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaStream_t cap = nullptr;
    struct GraphGuard {
        cudaGraph_t& g;
        cudaGraphExec_t& e;
        cudaStream_t& s;
        ~GraphGuard() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); if (s) cudaStreamDestroy(s); }
    } graph_guard{graph, graph_exec, cap};
    const char* graph_env = getenv("MOF_LEVEL_GRAPH");
    const bool use_graph = levels && !persist && !(graph_env && graph_env[0] == '0');
    int64_t launches_per_graph = 0;
    if (use_graph && h_active > 0 && max_iter > 0) {
        MOF_CUDA_TRY(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        const int64_t before = launches;
        MOF_CUDA_TRY(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        sweep_back(0, B.t, cap);
        sweep_fwd(0, B.p, B.ap, cap);
        level_update_kernel<<<grid, 256, 0, cap>>>(B, N);
        level_beta_kernel<<<G, 32, 0, cap>>>(B, inv_omega);
        ++launches;
        const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
        launches_per_graph = launches - before + 1;
        launches = before;
        if (ce != cudaSuccess) return mof_set_error(-100, "mof_pcg_solve_batch: graph capture failed: %s", cudaGetErrorString(ce));
        MOF_CUDA_TRY(cudaGraphInstantiate(&graph_exec, graph, 0));
    }
    for (;;) {
        while (h_active > 0 && it < max_iter) {
            const int n = (max_iter - it) < check_every ? (max_iter - it) : check_every;
            if (persist) {               // one cooperative launch = n whole iterations
                int timing = prof ? 1 : 0;
                int n_arg = n;
                void* args[] = {(void*)&largs, (void*)&n_arg, (void*)&stamp, (void*)&timing};
                if (prof) cudaEventRecord(ev[4], st);
                MOF_CUDA_TRY(cudaLaunchCooperativeKernel(iter_fn, dim3(grid_iter), dim3(256), args, persist_smem, st));
                if (prof) cudaEventRecord(ev[5], st);
                stamp += 2 * n;
                launches += 1;
            } else
            for (int q = 0; q < n; ++q) {
                const bool sample = prof && q == 0;
                if (sample) cudaEventRecord(ev[0], st);
                if (!ssor) {
                    pupdate_kernel<<<grid, 256, 0, st>>>(B, N);
                    if (sample) cudaEventRecord(ev[1], st);
                    spmv_kernel<true><<<grid, 256, 0, st>>>(mesh->rowptr, mesh->col, B.vals, B.p, B.ap, N, nb, ntiles,
                                                           B.partial, B.scal, B.state, G);
                    if (sample) cudaEventRecord(ev[2], st);
                    update_kernel<false><<<grid, 256, 0, st>>>(B, N, ntiles, 0.0);
                    launches += 3;
                } else if (graph_exec && !sample) {
                    MOF_CUDA_TRY(cudaGraphLaunch(graph_exec, st));
                    launches += launches_per_graph;
                } else {
                    sweep_back(0, B.t, st);                         // x += alpha p ; p <- zs r/omega + beta p ; t = (Dt+U)^-1 p
                    if (sample) cudaEventRecord(ev[1], st);
                    sweep_fwd(0, B.p, B.ap, st);                    // w ; alpha
                    if (sample) cudaEventRecord(ev[2], st);
                    if (levels) {
                        level_update_kernel<<<grid, 256, 0, st>>>(B, N);
                        level_beta_kernel<<<G, 32, 0, st>>>(B, inv_omega);
                        launches += 2;
                    } else {
                        update_kernel<true><<<grid, 256, 0, st>>>(B, N, ntiles, inv_omega);
                        launches += 1;
                    }
                }
                if (sample) cudaEventRecord(ev[3], st);
            }
            MOF_LAUNCH_CHECK("pcg iteration kernels");
            it += n;
            const int32_t groups_before = h_act[0], lanes_before = h_act[1];
            MOF_CUDA_TRY(cudaMemcpyAsync(h_act, d_active_groups, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            MOF_CUDA_TRY(cudaStreamSynchronize(st));
            if (prof && persist) {
                float ms = 0;
                MOF_CUDA_TRY(cudaEventElapsedTime(&ms, ev[4], ev[5]));
                prof->ms_iter += ms;
                prof->iter_launches += 1;
                prof->iterations_total += n;
            } else if (prof) {
                float ms[3] = {0, 0, 0};
                for (int q = 0; q < 3; ++q) MOF_CUDA_TRY(cudaEventElapsedTime(&ms[q], ev[q], ev[q + 1]));
                // block Jacobi: [pupdate, spmv, update]; SSOR: [backward sweeps, forward sweeps, update]
                prof->ms_spmv += ssor ? ms[0] : ms[1];
                prof->ms_pupdate += ssor ? ms[1] : ms[0];
                prof->ms_update += ms[2];
                prof->samples += 1;
                prof->group_launches += groups_before;
                prof->frame_launches += lanes_before;
                prof->iterations_total += n;
            }
            if (ssor && !calibrated && calibrate && h_active > 0 && it < max_iter) {                 // once: thresholds from the measured true / recurrence ratio
                calibrated = true;
                double* ax = B.z;
                sweep_back(1, B.t, st);
                spmv_kernel<false><<<grid, 256, 0, st>>>(mesh->rowptr, mesh->col, B.vals, B.t, ax, N, nb, ntiles, nullptr,
                                                        nullptr, nullptr, G);
                init_kernel<<<grid, 256, 0, st>>>(B, N, ntiles, MODE_CALIBRATE, tol2, 0, 1, inv_omega, 0, ax, calib_margin);
                launches += 2;
                MOF_LAUNCH_CHECK("calibration kernels");
            }
        }
        // confirm on the true residual b - A x; frames that miss tol resume with a tighter threshold
        if (ssor) sweep_back(1, B.t, st);                            // xs = (Dt+U)^-1 (xhat + pending alpha p)
        // A x goes to z on the SSOR paths (free there; the level path's w must keep its validity bits for a resume)
        double* ax = ssor ? B.z : B.ap;
        spmv_kernel<false><<<grid, 256, 0, st>>>(mesh->rowptr, mesh->col, B.vals, xphys, ax, N, nb, ntiles, nullptr,
                                                nullptr, nullptr, G);
        const int last_round = (rounds >= max_restarts || it >= max_iter) ? 1 : 0;
        init_kernel<<<grid, 256, 0, st>>>(B, N, ntiles, MODE_VERIFY, tol2, last_round, ssor ? 1 : 0, inv_omega, 0, ax);
        launches += 2;
        MOF_LAUNCH_CHECK("verification kernels");
        if (persist_err != cudaSuccess)
            return mof_set_error(-100 - (int)persist_err, "mof_pcg_solve_batch: cooperative launch of a level sweep failed: %s",
                                 cudaGetErrorString(persist_err));
        MOF_CUDA_TRY(cudaMemcpyAsync(h_act, d_active_groups, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MOF_CUDA_TRY(cudaStreamSynchronize(st));
        if (h_active <= 0 || last_round) break;
        if (levels) {
            restamp_kernel<<<grid, 256, 0, st>>>(B, N);
            MOF_LAUNCH_CHECK("restamp_kernel");
            launches += 1;
        }
        ++rounds;
    }
    if (ssor) {  // hand the solution of the original system back in batch->x (mof_unpack_solution reads it)
        unscale_kernel<<<grid, 256, 0, st>>>(B, N);
        MOF_LAUNCH_CHECK("unscale_kernel");
        launches += 1;
    }
    if (probe_buf) {
        std::vector<unsigned long long> h(16 + 2 * (size_t)N);
        MOF_CUDA_TRY(cudaMemcpyAsync(h.data(), probe_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MOF_CUDA_TRY(cudaStreamSynchronize(st));
        for (int d = 0; d < 2; ++d) {
            const double n = (double)(h[d * 8 + 4] ? h[d * 8 + 4] : 1);
            fprintf(stderr, "[mof probe] %s sweep: items %llu; cycles per item: poll + gather %.0f, stage wait %.0f, "
                            "arithmetic + result stores %.0f, tail + head %.0f\n",
                    d ? "forward" : "backward", h[d * 8 + 4], h[d * 8 + 0] / n, h[d * 8 + 1] / n, h[d * 8 + 2] / n, h[d * 8 + 3] / n);
        }
        if (const char* path = getenv("MOF_LEVEL_PROBE_FILE")) {       // per-row finish times of the last iteration (group act[0])
            if (FILE* f = fopen(path, "wb")) {
                fwrite(h.data() + 16, sizeof(unsigned long long), 2 * (size_t)N, f);
                fclose(f);
            }
        }
    }
    if (prof) prof->launches_total += launches;
    if (prof && persist) {
        double ns[4] = {0, 0, 0, 0};
        MOF_CUDA_TRY(cudaMemcpyAsync(ns, scal_host_ptr(B.scal, 0, MOF_S_SPARE), sizeof(ns), cudaMemcpyDeviceToHost, st));
        MOF_CUDA_TRY(cudaStreamSynchronize(st));
        for (int k = 0; k < 4; ++k) prof->phase_ns[k] += ns[k];
    }

    // per-frame report
    const size_t nf = (size_t)G * MOF_W;
    int32_t* h_state = (int32_t*)malloc(nf * MOF_I_COUNT * sizeof(int32_t));
    double* h_scal = (double*)malloc(nf * MOF_S_COUNT * sizeof(double));
    if (!h_state || !h_scal) { free(h_state); free(h_scal); return mof_set_error(-3, "out of host memory"); }
    cudaError_t e1 = cudaMemcpyAsync(h_state, B.state, nf * MOF_I_COUNT * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaMemcpyAsync(h_scal, B.scal, nf * MOF_S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaError_t e3 = cudaStreamSynchronize(st);
    int worst = 0;
    if (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
        for (int g = 0; g < G; ++g)
            for (int l = 0; l < MOF_W; ++l) {
                const size_t k = (size_t)g * MOF_W + l;
                int32_t s = h_state[((size_t)g * MOF_I_COUNT + MOF_I_STATUS) * MOF_W + l];
                if (s == MOF_STATUS_PENDING) s = MOF_STATUS_MAXITER;
                const double bb = h_scal[((size_t)g * MOF_S_COUNT + MOF_S_BBT) * MOF_W + l];
                const double rt = h_scal[((size_t)g * MOF_S_COUNT + MOF_S_RRTRUE) * MOF_W + l];
                if (iters) iters[k] = h_state[((size_t)g * MOF_I_COUNT + MOF_I_ITERS) * MOF_W + l];
                if (relres) relres[k] = bb > 0 ? sqrt(rt / bb) : 0.0;
                if (status) status[k] = s;
                if ((int64_t)k < B.n_frames && s != MOF_STATUS_CONVERGED && s != MOF_STATUS_ZERO_RHS && s > worst) worst = s;
            }
    }
    free(h_state);
    free(h_scal);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return mof_set_error(-100, "mof_pcg_solve_batch: report copy failed: %s",
                             cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    if (worst) mof_set_error(worst, "mof_pcg_solve_batch: at least one frame did not converge (status %d)", worst);
    return worst;
}

// K1 -- per-frame P1 assembly for a batch of frames (worker face loop,
// utils/compute_optical_flow.py:113-141, and a = a1 + lambda_*a2, :144).
//
// Thread mapping: lane = frame inside a group of 32, warp = block row.  Every block of a
// row gathers its contributing faces in ascending face order (the reference's accumulation
// order) from the per-block contributor lists built once per mesh, so there are no
// atomics and the result is bit-reproducible.  Mesh data (e, grad_w, tri, integral, a2)
// is warp-uniform: one broadcast transaction serves 32 frames.  Per-frame data (It, dIt
// in, vals/rhs/minv out) is frame-minor: each warp access is one 256-byte line.
//
// HBM bytes per frame: write vals 32 nb + rhs 16 N + minv 24 N,
// read It/dIt 16 N (+ mesh data / 32).  At N = 163,842: 36.7 + 2.6 + 3.9 + 2.6 MB.
#include "mof_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRowsPerWarp = MOF_TILE_ROWS / kWarpsPerCta;

// (T,N) row-major frames in reference vertex order -> It/dIt [G][N][32], renumbered.
__global__ void __launch_bounds__(256) pack_kernel(int64_t N, int32_t n_frames, const int32_t* __restrict__ perm,
                                                   const double* __restrict__ I_now,
                                                   const double* __restrict__ I_next, int64_t ld,
                                                   const double* __restrict__ dt, double* __restrict__ It,
                                                   double* __restrict__ dIt) {
    __shared__ double sI[32][33];
    __shared__ double sD[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t g = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    const int64_t v = v0 + tx;
    const int64_t o = v < N ? perm[v] : 0;
    for (int fr = ty; fr < 32; fr += 8) {
        int64_t k = g * 32 + fr;
        double a = 0.0, d = 0.0;
        if (k < n_frames && v < N) {
            a = I_now[k * ld + o];
            d = (I_next[k * ld + o] - a) / dt[k];       // cof:307: (I_kplus1[i] - I_k[i]) / t
        }
        sI[fr][tx] = a;
        sD[fr][tx] = d;
    }
    __syncthreads();
    for (int vv = ty; vv < 32; vv += 8) {
        int64_t w = v0 + vv;
        if (w < N) {
            It[mof_ix_sca(N, g, w) + tx] = sI[tx][vv];
            dIt[mof_ix_sca(N, g, w) + tx] = sD[tx][vv];
        }
    }
}

// K1 in two launches.  First the diagonal blocks (the row's incident faces: a1 + lambda a2 and the rhs f, cof:113-146),
// from which everything per-vertex follows: block Jacobi -> D^-1; SSOR -> S = D^-1/2, the scaled rhs S f and the
// scaled diagonal block S D S.  Then every off-diagonal block is assembled and written ONCE, already scaled
// (Ah_ij = S_i A_ij S_j needs the neighbour's S_j, complete after the first launch).  Round 1 wrote the unscaled
// matrix and then read-modified-wrote all of it in a separate scaling kernel (2 x 32 nb bytes per frame more).
// Both launches are latency-bound on their index chains (contributor list -> face -> vertex values), so occupancy
// decides: capped at 64 registers (4 CTAs = 32 warps per SM, 40 bytes of spills) they take 35.4 ms per 999-frame batch,
// uncapped (74 registers, 3 CTAs) 41.6 ms, at 48 registers 43.7 ms.
__global__ void __launch_bounds__(256, 4) assemble_diag_kernel(mof_mesh_dev M, mof_batch_dev B, double lambda_, bool ssor) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = blockIdx.y;
    const int64_t N = M.n_vertices, nb = M.n_blocks;
    const double* It_l = B.It + mof_ix_sca(N, g, 0) + lane;
    const double* dIt_l = B.dIt + mof_ix_sca(N, g, 0) + lane;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int64_t v = row0 + rr;
        if (v >= N) break;
        const int32_t bd = M.diag[v];
        double a[4], f[2], mi[3];
        mof_assemble_block_body<true>(M, v, bd, It_l, dIt_l, lambda_, a, f);
        if (!ssor) {
            mof_inv2_body(a, mi);                              // block Jacobi: D^-1
        } else {
            mof_inv_sqrt2_body(a, mi);                         // SSOR: S = D^-1/2
            double o[4];
            mof_scale_block_body(mi, mi, a, o);                // S D S (identity up to rounding; the SpMV reads it)
            for (int c = 0; c < 4; ++c) a[c] = o[c];
            const double f0 = f[0], f1 = f[1];
            f[0] = mi[0] * f0 + mi[1] * f1;                    // bh = S b
            f[1] = mi[1] * f0 + mi[2] * f1;
        }
        B.rhs[mof_ix_vec(N, g, v, 0) + lane] = f[0];
        B.rhs[mof_ix_vec(N, g, v, 1) + lane] = f[1];
        B.minv[mof_ix_minv(N, g, v, 0) + lane] = mi[0];
        B.minv[mof_ix_minv(N, g, v, 1) + lane] = mi[1];
        B.minv[mof_ix_minv(N, g, v, 2) + lane] = mi[2];
        double* out = B.vals + mof_ix_val(nb, g, bd, 0) + lane;
        out[0] = a[0];
        out[MOF_W] = a[1];
        out[2 * MOF_W] = a[2];
        out[3 * MOF_W] = a[3];
    }
}

__global__ void __launch_bounds__(256, 4) assemble_offdiag_kernel(mof_mesh_dev M, mof_batch_dev B, double lambda_, bool ssor) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = blockIdx.y;
    const int64_t N = M.n_vertices, nb = M.n_blocks;
    const double* It_l = B.It + mof_ix_sca(N, g, 0) + lane;
    const double* dIt_l = B.dIt + mof_ix_sca(N, g, 0) + lane;
    const int64_t row0 = (int64_t)blockIdx.x * MOF_TILE_ROWS + warp * kRowsPerWarp;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int64_t v = row0 + rr;
        if (v >= N) break;
        const int32_t bs = M.rowptr[v], be = M.rowptr[v + 1], bd = M.diag[v];
        double si[3] = {0.0, 0.0, 0.0};
        if (ssor)
            for (int c = 0; c < 3; ++c) si[c] = B.minv[mof_ix_minv(N, g, v, c) + lane];
        for (int32_t b = bs; b < be; ++b) {
            if (b == bd) continue;
            double a[4], f[2];
            mof_assemble_block_body<false>(M, v, b, It_l, dIt_l, lambda_, a, f);
            if (ssor) {
                const int64_t j = M.col[b];
                double sj[3], o[4];
                for (int c = 0; c < 3; ++c) sj[c] = B.minv[mof_ix_minv(N, g, j, c) + lane];
                mof_scale_block_body(si, sj, a, o);
                for (int c = 0; c < 4; ++c) a[c] = o[c];
            }
            double* out = B.vals + mof_ix_val(nb, g, b, 0) + lane;
            __stcs(out, a[0]);
            __stcs(out + MOF_W, a[1]);
            __stcs(out + 2 * MOF_W, a[2]);
            __stcs(out + 3 * MOF_W, a[3]);
        }
    }
}

}  // namespace

extern "C" int64_t mof_num_tiles(int64_t n_vertices) { return (n_vertices + MOF_TILE_ROWS - 1) / MOF_TILE_ROWS; }

extern "C" int mof_pack_frames(const mof_mesh_dev* mesh, const mof_batch_dev* batch, const double* I_now,
                               const double* I_next, int64_t ld, const double* dt, void* stream) {
    MOF_REQUIRE(mesh && batch && I_now && I_next && dt, "NULL argument");
    MOF_REQUIRE(batch->n_groups > 0 && batch->n_frames >= 0 && batch->n_frames <= batch->n_groups * MOF_GROUP,
                "n_frames does not fit n_groups");
    MOF_REQUIRE(ld >= mesh->n_vertices, "ld < n_vertices");
    dim3 grid(mof_cdiv(mesh->n_vertices, 32), batch->n_groups), block(32, 8);
    pack_kernel<<<grid, block, 0, mof_stream(stream)>>>(mesh->n_vertices, batch->n_frames, mesh->perm, I_now,
                                                        I_next, ld, dt, batch->It, batch->dIt);
    MOF_LAUNCH_CHECK("pack_kernel");
    return 0;
}

extern "C" int mof_assemble_batch(const mof_mesh_dev* mesh, const mof_batch_dev* batch, double lambda_,
                                  double omega, void* stream) {
    MOF_REQUIRE(mesh && batch, "NULL argument");
    MOF_REQUIRE(omega >= 0.0 && omega < 2.0, "omega must be 0 (block Jacobi) or in (0,2) (SSOR)");
    MOF_REQUIRE(batch->It && batch->dIt && batch->vals && batch->rhs && batch->minv, "batch buffers missing");
    dim3 grid((unsigned)mof_num_tiles(mesh->n_vertices), batch->n_groups);
    assemble_diag_kernel<<<grid, 256, 0, mof_stream(stream)>>>(*mesh, *batch, lambda_, omega > 0.0);
    MOF_LAUNCH_CHECK("assemble_diag_kernel");
    assemble_offdiag_kernel<<<grid, 256, 0, mof_stream(stream)>>>(*mesh, *batch, lambda_, omega > 0.0);
    MOF_LAUNCH_CHECK("assemble_offdiag_kernel");
    return 0;
}

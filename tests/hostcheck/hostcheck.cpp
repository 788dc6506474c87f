// TEST INFRASTRUCTURE ONLY (built by tests/hostcheck/__init__.py with g++, no CUDA).
// Wraps the __host__ __device__ kernel bodies of csrc/mof_bodies.h in plain loops with
// the same frame-minor layout as the CUDA kernels, so that indexing and operation order
// can be checked against the oracle in the GPU-less build container.  The product
// library (libmof_b200.so) never runs these loops; its entry points launch CUDA kernels.
#include <cstdint>
#include <cstring>

#include "mof_b200.h"
#include "mof_bodies.h"

#include <cmath>
#include <vector>

extern "C" {

void hc_geom(const mof_mesh_dev* M, const double* coords, const double* normals, double* e, double* grad_w,
             double* integral) {
    for (int64_t v = 0; v < M->n_vertices; ++v) mof_basis_body(normals + 3 * v, e + 6 * v);
    for (int64_t f = 0; f < M->n_faces; ++f) mof_face_geom_body(coords, M->tri, M->areas, f, grad_w, integral);
}

void hc_a2(const mof_mesh_dev* M, double* a2v) {
    for (int64_t v = 0; v < M->n_vertices; ++v)
        for (int32_t b = M->rowptr[v]; b < M->rowptr[v + 1]; ++b) mof_a2_block_body(*M, v, b, a2v + 4 * b);
}

// It/dIt [G][N][32] from (n_frames, N) rows (pack_kernel)
void hc_pack(const mof_mesh_dev* M, int32_t G, int32_t n_frames, const double* I_now, const double* I_next, int64_t ld,
             const double* dt, double* It, double* dIt) {
    const int64_t N = M->n_vertices;
    for (int64_t g = 0; g < G; ++g)
        for (int64_t v = 0; v < N; ++v)
            for (int l = 0; l < MOF_W; ++l) {
                int64_t k = g * MOF_W + l;
                double a = 0, d = 0;
                if (k < n_frames) {
                    a = I_now[k * ld + M->perm[v]];
                    d = (I_next[k * ld + M->perm[v]] - a) / dt[k];
                }
                It[mof_ix_sca(N, g, v) + l] = a;
                dIt[mof_ix_sca(N, g, v) + l] = d;
            }
}

void hc_assemble(const mof_mesh_dev* M, int32_t G, const double* It, const double* dIt, double lambda_, double omega,
                 double* vals, double* rhs, double* minv) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l) {
            const double* It_l = It + mof_ix_sca(N, g, 0) + l;
            const double* dIt_l = dIt + mof_ix_sca(N, g, 0) + l;
            for (int64_t v = 0; v < N; ++v)
                for (int32_t b = M->rowptr[v]; b < M->rowptr[v + 1]; ++b) {
                    double a[4], f[2];
                    if (b == M->diag[v]) {
                        mof_assemble_block_body<true>(*M, v, b, It_l, dIt_l, lambda_, a, f);
                        double mi[3];
                        if (omega == 0.0) mof_inv2_body(a, mi);
                        else              mof_inv_sqrt2_body(a, mi);
                        for (int c = 0; c < 2; ++c) rhs[mof_ix_vec(N, g, v, c) + l] = f[c];
                        for (int c = 0; c < 3; ++c) minv[mof_ix_minv(N, g, v, c) + l] = mi[c];
                    } else {
                        mof_assemble_block_body<false>(*M, v, b, It_l, dIt_l, lambda_, a, f);
                    }
                    for (int c = 0; c < 4; ++c) vals[mof_ix_val(nb, g, b, c) + l] = a[c];
                }
            if (omega == 0.0) continue;
            // scale_kernel: Ah = S A S, bh = S b
            for (int64_t v = 0; v < N; ++v) {
                double si[3];
                for (int c = 0; c < 3; ++c) si[c] = minv[mof_ix_minv(N, g, v, c) + l];
                double f0 = rhs[mof_ix_vec(N, g, v, 0) + l], f1 = rhs[mof_ix_vec(N, g, v, 1) + l];
                rhs[mof_ix_vec(N, g, v, 0) + l] = si[0] * f0 + si[1] * f1;
                rhs[mof_ix_vec(N, g, v, 1) + l] = si[1] * f0 + si[2] * f1;
                for (int32_t b = M->rowptr[v]; b < M->rowptr[v + 1]; ++b) {
                    double sj[3], a[4], o[4];
                    for (int c = 0; c < 3; ++c) sj[c] = minv[mof_ix_minv(N, g, M->col[b], c) + l];
                    for (int c = 0; c < 4; ++c) a[c] = vals[mof_ix_val(nb, g, b, c) + l];
                    mof_scale_block_body(si, sj, a, o);
                    for (int c = 0; c < 4; ++c) vals[mof_ix_val(nb, g, b, c) + l] = o[c];
                }
            }
        }
}

void hc_spmv(const mof_mesh_dev* M, int32_t G, const double* vals, const double* x, double* y) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l)
            for (int64_t v = 0; v < N; ++v) {
                double y0 = 0, y1 = 0;
                for (int32_t b = M->rowptr[v]; b < M->rowptr[v + 1]; ++b) {
                    int64_t j = M->col[b];
                    double x0 = x[mof_ix_vec(N, g, j, 0) + l], x1 = x[mof_ix_vec(N, g, j, 1) + l];
                    y0 += vals[mof_ix_val(nb, g, b, 0) + l] * x0 + vals[mof_ix_val(nb, g, b, 1) + l] * x1;
                    y1 += vals[mof_ix_val(nb, g, b, 2) + l] * x0 + vals[mof_ix_val(nb, g, b, 3) + l] * x1;
                }
                y[mof_ix_vec(N, g, v, 0) + l] = y0;
                y[mof_ix_vec(N, g, v, 1) + l] = y1;
            }
}

// SSOR sweeps over tiles [tile0, tile1) of one colour, all groups and lanes (sweep_*_kernel)
void hc_sweep_back(const mof_mesh_dev* M, int32_t G, const double* vals, const double* r, double* p, double* t,
                   int32_t tile0, int32_t tile1, const double* beta, const double* zs, double omega, int32_t mode) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l)
            for (int32_t tile = tile0; tile < tile1; ++tile) {
                int64_t r0 = (int64_t)tile * MOF_TILE_ROWS, r1 = r0 + MOF_TILE_ROWS < N ? r0 + MOF_TILE_ROWS : N;
                mof_sweep_back_body(M->rowptr, M->col, M->diag, vals + (size_t)g * nb * 4 * MOF_W + l,
                                    r + (size_t)g * N * 2 * MOF_W + l, p + (size_t)g * N * 2 * MOF_W + l,
                                    t + (size_t)g * N * 2 * MOF_W + l, r0, r1, beta ? beta[g * MOF_W + l] : 0.0,
                                    (zs ? zs[g * MOF_W + l] : 1.0) / omega, omega, mode);
            }
}

// dot[g][lane] += p'(t+w) over the tiles (mode 0)
void hc_sweep_fwd(const mof_mesh_dev* M, int32_t G, const double* vals, const double* pin, const double* t, double* w,
                  int32_t tile0, int32_t tile1, double omega, int32_t mode, double* dot) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l)
            for (int32_t tile = tile0; tile < tile1; ++tile) {
                int64_t r0 = (int64_t)tile * MOF_TILE_ROWS, r1 = r0 + MOF_TILE_ROWS < N ? r0 + MOF_TILE_ROWS : N;
                double d = mof_sweep_fwd_body(M->rowptr, M->col, M->diag, vals + (size_t)g * nb * 4 * MOF_W + l,
                                              pin + (size_t)g * N * 2 * MOF_W + l, t + (size_t)g * N * 2 * MOF_W + l,
                                              w + (size_t)g * N * 2 * MOF_W + l, r0, r1, omega, mode);
                if (dot) dot[g * MOF_W + l] += d;
            }
}

// The same sweeps over an arbitrary row range [r0, r1) (level-scheduled path: one dependency level, or a single row)
void hc_sweep_rows_back(const mof_mesh_dev* M, int32_t G, const double* vals, const double* r, double* p, double* t,
                        int64_t r0, int64_t r1, const double* beta, const double* zs, double omega, int32_t mode) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l)
            mof_sweep_back_body(M->rowptr, M->col, M->diag, vals + (size_t)g * nb * 4 * MOF_W + l,
                                r + (size_t)g * N * 2 * MOF_W + l, p + (size_t)g * N * 2 * MOF_W + l,
                                t + (size_t)g * N * 2 * MOF_W + l, r0, r1, beta ? beta[g * MOF_W + l] : 0.0,
                                (zs ? zs[g * MOF_W + l] : 1.0) / omega, omega, mode);
}

void hc_sweep_rows_fwd(const mof_mesh_dev* M, int32_t G, const double* vals, const double* pin, const double* t, double* w,
                       int64_t r0, int64_t r1, double omega, int32_t mode, double* dot) {
    const int64_t N = M->n_vertices, nb = M->n_blocks;
    for (int64_t g = 0; g < G; ++g)
        for (int l = 0; l < MOF_W; ++l) {
            double d = mof_sweep_fwd_body(M->rowptr, M->col, M->diag, vals + (size_t)g * nb * 4 * MOF_W + l,
                                          pin + (size_t)g * N * 2 * MOF_W + l, t + (size_t)g * N * 2 * MOF_W + l,
                                          w + (size_t)g * N * 2 * MOF_W + l, r0, r1, omega, mode);
            if (dot) dot[g * MOF_W + l] += d;
        }
}

void hc_tangent(int64_t N, int64_t n_frames, const double* V, int64_t ldV, const double* e, double* Vxyz, double* speed,
                double* vmax) {
    for (int64_t k = 0; k < n_frames; ++k) {
        double mx = 0;
        for (int64_t i = 0; i < N; ++i) {
            double* o = Vxyz + (k * N + i) * 3;
            mof_tangent_body(V[k * ldV + i], V[k * ldV + N + i], e + 6 * i, o);
            double len = mof_len3_body(o);
            speed[k * N + i] = len;
            if (!(len <= mx)) mx = len;
        }
        vmax[k] = mx;
    }
}

// one frame; returns counts; lists ascending
void hc_detect(int64_t N, int64_t F, const double* coords, const int32_t* tri, const double* Vxyz, double vmax, double eps,
               int32_t* nv, int32_t* vidx, int32_t* nf, int32_t* fidx, double* lam_mu, int8_t* sign) {
    uint8_t* vf = new uint8_t[N];
    *nv = *nf = 0;
    for (int64_t i = 0; i < N; ++i) {
        vf[i] = mof_vertex_zero_body(Vxyz + 3 * i, vmax, eps);
        if (vf[i]) vidx[(*nv)++] = (int32_t)i;
    }
    for (int64_t t = 0; t < F; ++t) {
        int64_t a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        if (vf[a] | vf[b] | vf[c]) continue;
        double l, m;
        int s;
        if (mof_face_zero_body(coords + 3 * a, coords + 3 * b, coords + 3 * c, Vxyz + 3 * a, Vxyz + 3 * b, Vxyz + 3 * c, vmax,
                               &l, &m, &s)) {
            fidx[*nf] = (int32_t)t;
            lam_mu[2 * *nf] = l;
            lam_mu[2 * *nf + 1] = m;
            sign[*nf] = (int8_t)s;
            ++*nf;
        }
    }
    delete[] vf;
}
// Sequential walk through the steps of csrc/winding.cu for one point (closest vertex, breadth-first
// rings with a visited mask, rank by (polar key, vertex id), sequential angle sum, acceptance rule).
void hc_winding(int64_t N, const double* coords, const double* V, const double* e, const int32_t* ring_ptr,
                const int32_t* ring_idx, const double* P, int max_level, int32_t* closest, int32_t* count, int32_t* type,
                double* winding) {
    double best = INFINITY;
    int arg = 0;
    for (int64_t v = 0; v < N; ++v) {
        const double d = mof_dist3_body(coords + 3 * v, P);
        if (d < best) { best = d; arg = (int)v; }
    }
    std::vector<uint8_t> seen(N, 0);
    std::vector<int32_t> cur{arg}, nxt;
    seen[arg] = 1;
    const double* O = coords + 3 * (size_t)arg;
    const double* e1 = e + 6 * (size_t)arg;
    const double* e2 = e1 + 3;
    int flag = 0;
    *count = 0;
    for (int l = 0; l < max_level; ++l) winding[l] = NAN;
    for (int level = 0; level < max_level; ++level) {
        nxt.clear();
        for (int32_t v : cur)
            for (int j = ring_ptr[v]; j < ring_ptr[v + 1]; ++j)
                if (!seen[ring_idx[j]]) { seen[ring_idx[j]] = 1; nxt.push_back(ring_idx[j]); }
        const int n = (int)nxt.size();
        if (n == 0) break;
        std::vector<double> key(n), vx(n), vy(n), sx(n), sy(n);
        for (int t = 0; t < n; ++t)
            mof_winding_element_body(O, coords + 3 * (size_t)nxt[t], V + 3 * (size_t)nxt[t], e1, e2, &key[t], &vx[t], &vy[t]);
        for (int t = 0; t < n; ++t) {
            int r = 0;
            for (int j = 0; j < n; ++j) r += (key[j] < key[t] || (key[j] == key[t] && nxt[j] < nxt[t])) ? 1 : 0;
            sx[r] = vx[t];
            sy[r] = vy[t];
        }
        double sum = 0.0;
        for (int t = 0; t < n; ++t) sum += mof_signed_angle_body(sx[t], sy[t], sx[(t + 1) % n], sy[(t + 1) % n]);
        const double w = sum / (2 * 3.141592653589793);
        winding[level] = w;
        if (!mof_winding_accept_body(level, w, &flag)) break;
        ++*count;
        cur.swap(nxt);
    }
    *closest = arg;
    *type = flag;
}

// K6: coefficient rows (wave_coef_kernel) ...
void hc_wave_coef(const mof_mesh_dev* M, double* cw, double* cg) {
    for (int64_t v = 0; v < M->n_vertices; ++v) mof_wave_coef_row_body(*M, v, cw, cg);
}

// ... and the row products of wave_rows_kernel on the packed signal It[G][N][32] (internal order), written to the
// caller's (n_out, N[, 3]) array in reference order.  C = 2: coef = cw -> wave speed; C = 3: coef = cg -> grad_point.
void hc_wave_rows(const mof_mesh_dev* M, int C, int64_t n_rows, int64_t out0, int64_t n_out, int64_t t_first, int64_t T_trial,
                  const double* I, int64_t ld, double dt, int phase_mode, const double* coef, double* out) {
    const int64_t N = M->n_vertices, G = (n_rows + MOF_W - 1) / MOF_W;
    std::vector<double> It((size_t)G * N * MOF_W, 0.0);                      // wave_pack_kernel
    for (int64_t k = 0; k < n_rows; ++k)
        for (int64_t v = 0; v < N; ++v) It[mof_ix_sca(N, k / MOF_W, v) + k % MOF_W] = I[k * ld + M->perm[v]];
    auto at = [&](int64_t v, int64_t row) { return It[mof_ix_sca(N, row >> 5, v) + (row & 31)]; };
    const double inv_dt = 1.0 / dt;
    for (int64_t r = out0; r < out0 + n_out; ++r)
        for (int64_t v = 0; v < N; ++v) {
            double acc[3] = {0.0, 0.0, 0.0};
            for (int32_t j = M->rowptr[v]; j < M->rowptr[v + 1]; ++j)
                for (int c = 0; c < C; ++c) acc[c] = fma(coef[(size_t)j * C + c], at(M->col[j], r), acc[c]);
            const size_t o = (size_t)(r - out0) * N + M->perm[v];
            if (C == 3) {
                for (int c = 0; c < 3; ++c) out[3 * o + c] = acc[c];
                continue;
            }
            const int64_t t = t_first + r;
            const bool first = t == 0, last = t == T_trial - 1;
            const double cur = at(v, r), prev = r > 0 ? at(v, r - 1) : 0.0, next = r + 1 < n_rows ? at(v, r + 1) : 0.0;
            double far2 = 0.0;
            if (!phase_mode && first) far2 = at(v, r + 2);
            else if (!phase_mode && last && r >= 2) far2 = at(v, r - 2);
            out[o] = mof_wave_speed_body(mof_wave_td_body(phase_mode, first, last, T_trial, cur, prev, next, far2, inv_dt), acc[0], acc[1]);
        }
}
}

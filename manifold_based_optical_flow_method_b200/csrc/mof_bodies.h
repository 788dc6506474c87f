// Scalar "bodies" of the kernels: the arithmetic one (row|face|vertex, frame) pair
// performs, written once as __host__ __device__ inline functions.  The CUDA kernels in
// geom.cu / assemble.cu / pcg.cu / detect.cu wrap them in the thread mapping; the
// CPU-only test helper tests/hostcheck/hostcheck.cpp wraps the same bodies in plain
// loops so that indexing and operation order can be checked against the oracle in the
// build container, which has no GPU.  (That helper is test infrastructure: the product
// library never calls these bodies on the host.)
//
// Reference citations: cof = utils/compute_optical_flow.py, fsp =
// utils/find_singularity_point.py under /root/reference.
#ifndef MOF_BODIES_H
#define MOF_BODIES_H

#include <math.h>
#include <stdint.h>

#include "mof_b200.h"

#if defined(__CUDACC__)
#define MOF_HD __host__ __device__ __forceinline__
#else
#define MOF_HD inline
#endif

// Exactly-rounded multiply/add that the compiler may not contract into an FMA, for
// the few places where bit-equality with numpy's separate multiply and add is tested.
#if defined(__CUDA_ARCH__)
#define MOF_MUL(a, b) __dmul_rn((a), (b))
#define MOF_ADD(a, b) __dadd_rn((a), (b))
#else
#define MOF_MUL(a, b) ((a) * (b))
#define MOF_ADD(a, b) ((a) + (b))
#endif

#define MOF_W MOF_GROUP

// ---- frame-minor layout offsets (elements) --------------------------------------
MOF_HD size_t mof_ix_vec(int64_t N, int64_t g, int64_t v, int c) { return (size_t)((g * N + v) * 2 + c) * MOF_W; }
MOF_HD size_t mof_ix_val(int64_t nb, int64_t g, int64_t b, int c) { return (size_t)((g * nb + b) * 4 + c) * MOF_W; }
MOF_HD size_t mof_ix_sca(int64_t N, int64_t g, int64_t v) { return (size_t)(g * N + v) * MOF_W; }
MOF_HD size_t mof_ix_minv(int64_t N, int64_t g, int64_t v, int c) { return (size_t)((g * N + v) * 3 + c) * MOF_W; }

// ---- K0 ----------------------------------------------------------------------
// compute_orthonormal_basis, cof:210-235.  e[0..2] = e1, e[3..5] = e2.
MOF_HD void mof_basis_body(const double* n, double* e) {
    double a0, a1, a2;
    if (n[0] != 0 || n[1] != 0) { a0 = -n[1]; a1 = n[0]; a2 = 0.0; }     // cof:222-223
    else                        { a0 = 0.0;  a1 = -n[2]; a2 = n[1]; }    // cof:225
    // e2 = n x e1 (cof:228)
    double b0 = MOF_ADD(MOF_MUL(n[1], a2), -MOF_MUL(n[2], a1));
    double b1 = MOF_ADD(MOF_MUL(n[2], a0), -MOF_MUL(n[0], a2));
    double b2 = MOF_ADD(MOF_MUL(n[0], a1), -MOF_MUL(n[1], a0));
    double na = sqrt(MOF_ADD(MOF_ADD(MOF_MUL(a0, a0), MOF_MUL(a1, a1)), MOF_MUL(a2, a2)));   // cof:231
    double nb = sqrt(MOF_ADD(MOF_ADD(MOF_MUL(b0, b0), MOF_MUL(b1, b1)), MOF_MUL(b2, b2)));   // cof:232
    e[0] = a0 / na; e[1] = a1 / na; e[2] = a2 / na;
    e[3] = b0 / nb; e[4] = b1 / nb; e[5] = b2 / nb;
}

// compute_gradient_w, cof:238-255: foot H of the perpendicular from p_i on line (p_j,p_k),
// returns (H - p_i) / |H - p_i|^2  (minus the textbook gradient; used consistently).
MOF_HD void mof_gradw_body(const double* pi, const double* pj, const double* pk, double* out) {
    double jk0 = pk[0] - pj[0], jk1 = pk[1] - pj[1], jk2 = pk[2] - pj[2];      // cof:249
    double ji0 = pi[0] - pj[0], ji1 = pi[1] - pj[1], ji2 = pi[2] - pj[2];      // cof:250
    double d1 = ji0 * jk0 + ji1 * jk1 + ji2 * jk2;
    double d2 = jk0 * jk0 + jk1 * jk1 + jk2 * jk2;
    double q0 = d1 * jk0 / d2, q1 = d1 * jk1 / d2, q2 = d1 * jk2 / d2;          // cof:251-252
    double h0 = pj[0] - pi[0] + q0, h1 = pj[1] - pi[1] + q1, h2 = pj[2] - pi[2] + q2;   // cof:253
    double hh = h0 * h0 + h1 * h1 + h2 * h2;
    out[0] = h0 / hh; out[1] = h1 / hh; out[2] = h2 / hh;                         // cof:254
}

// One face: grad_w[f][0..2] with the argument orders of cof:63-68, integral (cof:73-75).
MOF_HD void mof_face_geom_body(const double* coords, const int32_t* tri, const double* areas,
                               int64_t f, double* grad_w, double* integral) {
    const double* A = coords + 3 * (int64_t)tri[3 * f];
    const double* B = coords + 3 * (int64_t)tri[3 * f + 1];
    const double* C = coords + 3 * (int64_t)tri[3 * f + 2];
    mof_gradw_body(A, B, C, grad_w + 9 * f);
    mof_gradw_body(B, A, C, grad_w + 9 * f + 3);
    mof_gradw_body(C, A, B, grad_w + 9 * f + 6);
    integral[2 * f] = areas[f] / 6;
    integral[2 * f + 1] = areas[f] / 12;
}

// a2 block b of row v: sum over contributing faces (ascending) of
// (e_v^al . e_j^be) * (g_m . g_n) * A_T   (compute_a2 cof:258-270, accumulation cof:78-93).
MOF_HD void mof_a2_block_body(const mof_mesh_dev& M, int64_t v, int64_t b, double* out) {
    const double* ei = M.e + 6 * v;
    const double* ej = M.e + 6 * (int64_t)M.col[b];
    double ee[4];
    for (int al = 0; al < 2; ++al)
        for (int be = 0; be < 2; ++be)
            ee[2 * al + be] = ei[3 * al] * ej[3 * be] + ei[3 * al + 1] * ej[3 * be + 1] + ei[3 * al + 2] * ej[3 * be + 2];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int32_t q = M.cptr[b]; q < M.cptr[b + 1]; ++q) {
        int32_t ent = M.centry[q];
        int64_t f = ent >> 4;
        int m = (ent >> 2) & 3, n = ent & 3;
        const double* gm = M.grad_w + 9 * f + 3 * m;
        const double* gn = M.grad_w + 9 * f + 3 * n;
        double gg = gm[0] * gn[0] + gm[1] * gn[1] + gm[2] * gn[2];
        double A = M.areas[f];
        for (int c = 0; c < 4; ++c) acc[c] += ee[c] * gg * A;
    }
    for (int c = 0; c < 4; ++c) out[c] = acc[c];
}

// ---- K1 ----------------------------------------------------------------------
// Block b (row v) of a = a1 + lambda*a2 for one frame (lane), plus -- on the diagonal
// block -- the rhs f and the inverse of the diagonal block.
//   It_l / dIt_l : base of this group's It / dIt already offset by the lane; vertex u
//                  lives at [u * MOF_W].
//   grad_M_I (cof:116-117), compute_a1 (cof:285), compute_f (cof:305-311).
template <bool DIAG>
MOF_HD void mof_assemble_block_body(const mof_mesh_dev& M, int64_t v, int64_t b, const double* It_l,
                                    const double* dIt_l, double lambda_, double* a, double* f) {
    const double* ei = M.e + 6 * v;
    const double* ej = M.e + 6 * (int64_t)M.col[b];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0, f0 = 0.0, f1 = 0.0;
    for (int32_t q = M.cptr[b]; q < M.cptr[b + 1]; ++q) {
        int32_t ent = M.centry[q];
        int64_t fc = ent >> 4;
        int m = (ent >> 2) & 3, n = ent & 3;
        int64_t t0 = M.tri[3 * fc], t1 = M.tri[3 * fc + 1], t2 = M.tri[3 * fc + 2];
        const double* g = M.grad_w + 9 * fc;
        double I0 = It_l[t0 * MOF_W], I1 = It_l[t1 * MOF_W], I2 = It_l[t2 * MOF_W];
        double Gx = I0 * g[0] + I1 * g[3] + I2 * g[6];
        double Gy = I0 * g[1] + I1 * g[4] + I2 * g[7];
        double Gz = I0 * g[2] + I1 * g[5] + I2 * g[8];
        double ci0 = Gx * ei[0] + Gy * ei[1] + Gz * ei[2];
        double ci1 = Gx * ei[3] + Gy * ei[4] + Gz * ei[5];
        double cj0 = Gx * ej[0] + Gy * ej[1] + Gz * ej[2];
        double cj1 = Gx * ej[3] + Gy * ej[4] + Gz * ej[5];
        double integ = M.integral[2 * fc + (m == n ? 0 : 1)];           // cof:131-132
        acc0 += ci0 * cj0 * integ;
        acc1 += ci0 * cj1 * integ;
        acc2 += ci1 * cj0 * integ;
        acc3 += ci1 * cj1 * integ;
        if (DIAG) {
            double d0 = dIt_l[t0 * MOF_W], d1 = dIt_l[t1 * MOF_W], d2 = dIt_l[t2 * MOF_W];
            double dm = m == 0 ? d0 : (m == 1 ? d1 : d2);
            double oth = m == 0 ? d1 + d2 : (m == 1 ? d0 + d2 : d0 + d1);
            double w = 2 * dm + oth;                                     // cof:311
            double A = M.areas[fc];
            f0 += ci0 * w * A / 12;
            f1 += ci1 * w * A / 12;
        }
    }
    const double* a2 = M.a2v + 4 * b;
    a[0] = acc0 + lambda_ * a2[0];                                      // cof:144
    a[1] = acc1 + lambda_ * a2[1];
    a[2] = acc2 + lambda_ * a2[2];
    a[3] = acc3 + lambda_ * a2[3];
    if (DIAG) { f[0] = f0; f[1] = f1; }
}

// inverse of the symmetric 2x2 diagonal block [[a0,a1],[a2,a3]] -> (m00, m01, m11)
MOF_HD void mof_inv2_body(const double* a, double* m) {
    double det = a[0] * a[3] - a[1] * a[2];
    double id = 1.0 / det;
    m[0] = a[3] * id;
    m[1] = -0.5 * (a[1] + a[2]) * id;
    m[2] = a[0] * id;
}

// ---- SSOR sweeps (Eisenstat form of the SSOR-preconditioned CG) ------------------
// The SSOR path first scales the system symmetrically with S = D^-1/2 (D = 2x2 diagonal blocks):
// Ah = S A S has identity diagonal blocks, bh = S b, x = S xh.  SSOR is invariant under this
// scaling (same iterates as SSOR on A), but the sweeps no longer need any diagonal data.
// With Ah = L + I + U in the block-multicolour numbering and Dt = I / omega the preconditioned
// operator is
//     At v = t + (Dt+L)^-1 (v - ((2-omega)/omega) t),   t = (Dt+U)^-1 v,
// i.e. one backward and one forward block-triangular solve, each reading half of the
// off-diagonal blocks.  A sweep walks the rows of one patch (tile) sequentially for one
// frame; patches of one colour are independent of each other.  Pointers carry the
// (group, lane) offset: vertex v component c at [(2v+c)*W], block b entry c at [(4b+c)*W].

// S = M^-1/2 of the SPD block M = [[a0,a1],[a2,a3]] (a1 ~ a2): sqrt(M) = (M + s I)/t with
// s = sqrt(det M), t = sqrt(tr M + 2 s), hence M^-1/2 = [[c+s, -b],[-b, a+s]] / (s t).
MOF_HD void mof_inv_sqrt2_body(const double* a, double* o) {
    const double b = 0.5 * (a[1] + a[2]);
    const double s = sqrt(a[0] * a[3] - b * b);
    const double t = sqrt(a[0] + a[3] + 2.0 * s);
    const double r = 1.0 / (s * t);
    o[0] = (a[3] + s) * r;
    o[1] = -b * r;
    o[2] = (a[0] + s) * r;
}

// Backward sweep over rows r1-1 .. r0 of a patch: t_i = omega (rhs_i - sum_{j>i} U_ij t_j).
//   mode 0 (iteration): rhs_i = zsw r_i + beta p_i and p_i <- rhs_i: the CG direction update
//                       p = z + beta p with z = Dt r = r / omega fused in (zsw = 1/omega; a frozen
//                       frame has zsw = 0, beta = 1 and keeps its p)
//   mode 1 (back-transform): rhs_i = p_i (p is the input vector and is left untouched)
MOF_HD void mof_sweep_back_body(const int32_t* rowptr, const int32_t* col, const int32_t* diag,
                                const double* vals_l, const double* r_l, double* p_l, double* t_l,
                                int64_t r0, int64_t r1, double beta, double zsw, double omega, int mode) {
    for (int64_t i = r1 - 1; i >= r0; --i) {
        double a0, a1;
        if (mode == 0) {
            a0 = zsw * r_l[(2 * i) * MOF_W] + beta * p_l[(2 * i) * MOF_W];
            a1 = zsw * r_l[(2 * i + 1) * MOF_W] + beta * p_l[(2 * i + 1) * MOF_W];
            p_l[(2 * i) * MOF_W] = a0;
            p_l[(2 * i + 1) * MOF_W] = a1;
        } else {
            a0 = p_l[(2 * i) * MOF_W];
            a1 = p_l[(2 * i + 1) * MOF_W];
        }
        for (int32_t b = diag[i] + 1; b < rowptr[i + 1]; ++b) {
            const int64_t j = col[b];
            const double* a = vals_l + (size_t)b * 4 * MOF_W;
            const double t0 = t_l[(2 * j) * MOF_W], t1 = t_l[(2 * j + 1) * MOF_W];
            a0 -= a[0] * t0 + a[MOF_W] * t1;
            a1 -= a[2 * MOF_W] * t0 + a[3 * MOF_W] * t1;
        }
        t_l[(2 * i) * MOF_W] = omega * a0;
        t_l[(2 * i + 1) * MOF_W] = omega * a1;
    }
}

// Forward sweep over rows r0 .. r1-1: w_i = omega (v_i - sum_{j<i} L_ij w_j).
//   mode 0 (iteration): v_i = p_i - ((2-omega)/omega) t_i ; returns sum_i p_i . (t_i + w_i)
//   mode 1 (transform a rhs): v_i = p_i ; returns 0
MOF_HD double mof_sweep_fwd_body(const int32_t* rowptr, const int32_t* col, const int32_t* diag,
                                 const double* vals_l, const double* p_l, const double* t_l, double* w_l,
                                 int64_t r0, int64_t r1, double omega, int mode) {
    double dot = 0.0;
    const double ks = (2.0 - omega) / omega;
    for (int64_t i = r0; i < r1; ++i) {
        const double p0 = p_l[(2 * i) * MOF_W], p1 = p_l[(2 * i + 1) * MOF_W];
        double a0 = p0, a1 = p1, t0 = 0.0, t1 = 0.0;
        if (mode == 0) {
            t0 = t_l[(2 * i) * MOF_W];
            t1 = t_l[(2 * i + 1) * MOF_W];
            a0 -= ks * t0;
            a1 -= ks * t1;
        }
        for (int32_t b = rowptr[i]; b < diag[i]; ++b) {
            const int64_t j = col[b];
            const double* a = vals_l + (size_t)b * 4 * MOF_W;
            const double w0 = w_l[(2 * j) * MOF_W], w1 = w_l[(2 * j + 1) * MOF_W];
            a0 -= a[0] * w0 + a[MOF_W] * w1;
            a1 -= a[2 * MOF_W] * w0 + a[3 * MOF_W] * w1;
        }
        const double o0 = omega * a0, o1 = omega * a1;
        w_l[(2 * i) * MOF_W] = o0;
        w_l[(2 * i + 1) * MOF_W] = o1;
        if (mode == 0) dot += p0 * (t0 + o0) + p1 * (t1 + o1);
    }
    return dot;
}

// Symmetric scaling of one off-diagonal/diagonal block: Ah = S_i A S_j (S symmetric, 3 values each)
MOF_HD void mof_scale_block_body(const double* si, const double* sj, const double* a, double* o) {
    // T = S_i A
    const double t00 = si[0] * a[0] + si[1] * a[2], t01 = si[0] * a[1] + si[1] * a[3];
    const double t10 = si[1] * a[0] + si[2] * a[2], t11 = si[1] * a[1] + si[2] * a[3];
    o[0] = t00 * sj[0] + t01 * sj[1];
    o[1] = t00 * sj[1] + t01 * sj[2];
    o[2] = t10 * sj[0] + t11 * sj[1];
    o[3] = t10 * sj[1] + t11 * sj[2];
}

// ---- K4 ----------------------------------------------------------------------
// process_V_k fsp:61-66: V1*e1 + V2*e2 (separate multiplies and add, as numpy does).
MOF_HD void mof_tangent_body(double v1, double v2, const double* e, double* out) {
    out[0] = MOF_ADD(MOF_MUL(v1, e[0]), MOF_MUL(v2, e[3]));
    out[1] = MOF_ADD(MOF_MUL(v1, e[1]), MOF_MUL(v2, e[4]));
    out[2] = MOF_ADD(MOF_MUL(v1, e[2]), MOF_MUL(v2, e[5]));
}
// fsp:161 / S3:132: sqrt(x^2 + y^2 + z^2)
MOF_HD double mof_len3_body(const double* v) {
    return sqrt(MOF_ADD(MOF_ADD(MOF_MUL(v[0], v[0]), MOF_MUL(v[1], v[1])), MOF_MUL(v[2], v[2])));
}

// ---- K5 ----------------------------------------------------------------------
// is_zero_velocity_vertex(V_i / vmax, eps), fsp:72-90,166.
MOF_HD bool mof_vertex_zero_body(const double* V, double vmax, double eps) {
    double t[3] = {V[0] / vmax, V[1] / vmax, V[2] / vmax};
    return mof_len3_body(t) <= eps;
}

// has_zero_velocity_interior, fsp:93-137, for one face; VA,VB,VC are the raw vertex
// velocities (divided by vmax here, fsp:178-179).  The reference solves the 3x2
// least-squares system M [lam mu]' = -VC_p with M = [VA_p-VC_p | VB_p-VC_p]; all three
// vectors lie in the face plane, so the system is solved exactly in an orthonormal basis
// (u, w) of that plane by Cramer's rule.  sign = per-face Poincare index (orientation of
// (VA_p-VC_p, VB_p-VC_p) relative to the face winding (B-A)x(C-A)).
MOF_HD bool mof_face_zero_body(const double* A, const double* B, const double* C, const double* VA,
                               const double* VB, const double* VC, double vmax, double* lam,
                               double* mu, int* sign) {
    double ab[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    double ac[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};   // fsp:113
    double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= nn; n[1] /= nn; n[2] /= nn;                                                                          // fsp:114
    double lu = sqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]);
    double u[3] = {ab[0] / lu, ab[1] / lu, ab[2] / lu};
    double w[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
    double P[3][2];
    const double* Vs[3] = {VA, VB, VC};
    for (int k = 0; k < 3; ++k) {
        double s[3] = {Vs[k][0] / vmax, Vs[k][1] / vmax, Vs[k][2] / vmax};                                       // fsp:178-179
        double dn = s[0] * n[0] + s[1] * n[1] + s[2] * n[2];
        double p[3] = {s[0] - dn * n[0], s[1] - dn * n[1], s[2] - dn * n[2]};                                    // fsp:117-119
        P[k][0] = p[0] * u[0] + p[1] * u[1] + p[2] * u[2];
        P[k][1] = p[0] * w[0] + p[1] * w[1] + p[2] * w[2];
    }
    double ax = P[0][0] - P[2][0], ay = P[0][1] - P[2][1];      // column 0 of M (fsp:122)
    double bx = P[1][0] - P[2][0], by = P[1][1] - P[2][1];      // column 1 of M
    double cx = -P[2][0], cy = -P[2][1];                        // rhs -VC_p (fsp:128)
    double det = ax * by - ay * bx;
    double l = (cx * by - cy * bx) / det;
    double m = (ax * cy - ay * cx) / det;
    *lam = l;
    *mu = m;
    *sign = det > 0 ? 1 : (det < 0 ? -1 : 0);
    return (l + m <= 1) && (l >= 0) && (m >= 0);               // fsp:130 (false for NaN)
}

// ---- Jacobian classification (fsp:355-498) -----------------------------------------------
// One near point's contribution to the 2x2 "Jacobian" the reference accumulates (fsp:383-399 /
// :442-458): the neighbour's velocity / vmax projected on the plane (e1, e2) and expressed in that
// basis (u, v), divided by the neighbour's offset from `origin` along e1 and e2.  Divisions by a
// zero offset give inf / nan exactly like the reference.
MOF_HD double mof_dot3_plain(const double* a, const double* b) {       // separate multiplies and adds, like numpy
    return MOF_ADD(MOF_ADD(MOF_MUL(a[0], b[0]), MOF_MUL(a[1], b[1])), MOF_MUL(a[2], b[2]));
}

MOF_HD void mof_jacobian_term_body(const double* origin, const double* X, const double* V, double vmax, const double* e1,
                                   const double* e2, double* J) {
    // no FMA contraction anywhere here: on symmetric meshes an offset along e1 / e2 can be exactly 0
    // in the reference's arithmetic, and the resulting inf / nan decide the class
    const double n[3] = {MOF_ADD(MOF_MUL(e1[1], e2[2]), -MOF_MUL(e1[2], e2[1])), MOF_ADD(MOF_MUL(e1[2], e2[0]), -MOF_MUL(e1[0], e2[2])),
                         MOF_ADD(MOF_MUL(e1[0], e2[1]), -MOF_MUL(e1[1], e2[0]))};
    const double nn = mof_dot3_plain(n, n);
    const double s[3] = {V[0] / vmax, V[1] / vmax, V[2] / vmax};
    const double sn = mof_dot3_plain(s, n);
    const double p[3] = {MOF_ADD(s[0], -(MOF_MUL(sn, n[0]) / nn)), MOF_ADD(s[1], -(MOF_MUL(sn, n[1]) / nn)),
                         MOF_ADD(s[2], -(MOF_MUL(sn, n[2]) / nn))};                                   // fsp:206-210
    const double u = mof_dot3_plain(p, e1) / mof_dot3_plain(e1, e1);                                     // fsp:266-267
    const double v = mof_dot3_plain(p, e2) / mof_dot3_plain(e2, e2);
    const double b[3] = {MOF_ADD(X[0], -origin[0]), MOF_ADD(X[1], -origin[1]), MOF_ADD(X[2], -origin[2])};   // fsp:231
    const double bn = mof_dot3_plain(b, n);
    const double q[3] = {MOF_ADD(b[0], -(MOF_MUL(bn, n[0]) / nn)), MOF_ADD(b[1], -(MOF_MUL(bn, n[1]) / nn)),
                         MOF_ADD(b[2], -(MOF_MUL(bn, n[2]) / nn))};                                   // fsp:235
    const double d1 = mof_dot3_plain(q, e1);                                                             // fsp:238-239
    const double d2 = mof_dot3_plain(q, e2);
    J[0] = MOF_ADD(J[0], u / d1); J[1] = MOF_ADD(J[1], u / d2);                                          // fsp:396-399
    J[2] = MOF_ADD(J[2], v / d1); J[3] = MOF_ADD(J[3], v / d2);
}

// classify_critical_point, fsp:463-498: 0 Node, 1 Focus, 2 Saddle, 3 Indeterminate
MOF_HD int mof_classify_body(const double* J) {
    const double trace = J[0] + J[3];
    const double det = J[0] * J[3] - J[1] * J[2];
    if (det > 0) return trace * trace > 4 * det ? 0 : 1;
    if (det < 0) return 2;
    return 3;
}

// find_nearest_edge_and_vertices, fsp:318-351, with its flat argmin over the three 3-vectors
// |cross(P - X, v)| / |v|: flat index 0 -> edge AB (0), 1 -> BC (1), anything else -> CA (2).
MOF_HD int mof_nearest_edge_body(const double* A, const double* B, const double* C, const double* P) {
    const double* X[3] = {A, B, C};
    const double* Y[3] = {B, C, A};
    double best = 0.0;
    int arg = -1;
    for (int k = 0; k < 3; ++k) {
        const double v[3] = {Y[k][0] - X[k][0], Y[k][1] - X[k][1], Y[k][2] - X[k][2]};
        const double w[3] = {P[0] - X[k][0], P[1] - X[k][1], P[2] - X[k][2]};
        const double c[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
        const double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int m = 0; m < 3; ++m) {
            const double d = fabs(c[m] / len);
            if (arg < 0 || d < best) { best = d; arg = 3 * k + m; }      // first minimum, like np.argmin
        }
    }
    return arg == 0 ? 0 : (arg == 1 ? 1 : 2);
}

// ---- multi-ring winding numbers (S7_winding_line.py:59-165) ----
// One ring vertex X with velocity V seen from the centre vertex O in its tangent basis (e1, e2):
// polar-angle sort key (S7:36-45, :97) and the tangent components of V (S7:26-33, :48-57).
MOF_HD void mof_winding_element_body(const double* O, const double* X, const double* V, const double* e1, const double* e2,
                                     double* key, double* vx, double* vy) {
    const double n[3] = {MOF_ADD(MOF_MUL(e1[1], e2[2]), -MOF_MUL(e1[2], e2[1])), MOF_ADD(MOF_MUL(e1[2], e2[0]), -MOF_MUL(e1[0], e2[2])),
                         MOF_ADD(MOF_MUL(e1[0], e2[1]), -MOF_MUL(e1[1], e2[0]))};
    const double nn = mof_dot3_plain(n, n);
    const double b[3] = {MOF_ADD(X[0], -O[0]), MOF_ADD(X[1], -O[1]), MOF_ADD(X[2], -O[2])};
    const double bn = mof_dot3_plain(b, n);
    const double q[3] = {MOF_ADD(b[0], -(MOF_MUL(bn, n[0]) / nn)), MOF_ADD(b[1], -(MOF_MUL(bn, n[1]) / nn)),
                         MOF_ADD(b[2], -(MOF_MUL(bn, n[2]) / nn))};
    *key = atan2(mof_dot3_plain(q, e2), mof_dot3_plain(q, e1));
    const double vn = mof_dot3_plain(V, n);
    const double t[3] = {MOF_ADD(V[0], -(MOF_MUL(vn, n[0]) / nn)), MOF_ADD(V[1], -(MOF_MUL(vn, n[1]) / nn)),
                         MOF_ADD(V[2], -(MOF_MUL(vn, n[2]) / nn))};
    *vx = mof_dot3_plain(t, e1) / mof_dot3_plain(e1, e1);
    *vy = mof_dot3_plain(t, e2) / mof_dot3_plain(e2, e2);
}

// angle_between_vectors, S7:59-74: signed angle from a to b, counter-clockwise positive; NaN for a zero vector.
MOF_HD double mof_signed_angle_body(double ax, double ay, double bx, double by) {
    const double la = sqrt(MOF_ADD(MOF_MUL(ax, ax), MOF_MUL(ay, ay)));
    const double lb = sqrt(MOF_ADD(MOF_MUL(bx, bx), MOF_MUL(by, by)));
    const double ux = ax / la, uy = ay / la, wx = bx / lb, wy = by / lb;
    double d = MOF_ADD(MOF_MUL(ux, wx), MOF_MUL(uy, wy));
    if (d > 1.0) d = 1.0;
    else if (d < -1.0) d = -1.0;
    double ang = acos(d);
    if (MOF_ADD(MOF_MUL(ux, wy), -MOF_MUL(uy, wx)) < 0.0) ang = -ang;
    return ang;
}

// The acceptance rule of S7:150-163 for the winding number w of ring `level`; *flag is the type
// fixed by the first ring (+1 / -1, 0 = neither).  Returns true when the ring counts.
MOF_HD bool mof_winding_accept_body(int level, double w, int* flag) {
    if (level == 0) {
        if (w >= -1.01 && w <= -0.99) { *flag = -1; return true; }
        if (w >= 0.99 && w <= 1.01) { *flag = 1; return true; }
        return false;
    }
    if (*flag == 1) return w >= 0.999 && w <= 1.001;
    if (*flag == -1) return w >= -1.001 && w <= -0.999;
    return false;
}

// distance of S7:130's find_closest_point (Euclidean, like np.linalg.norm of the difference)
MOF_HD double mof_dist3_body(const double* a, const double* b) {
    const double d[3] = {MOF_ADD(a[0], -b[0]), MOF_ADD(a[1], -b[1]), MOF_ADD(a[2], -b[2])};
    return sqrt(mof_dot3_plain(d, d));
}

// scipy.interpolate.Rbf multiquadric (scipy/interpolate/_rbf.py _h_multiquadric with cdist's Euclidean
// norm): phi = sqrt((1/eps * |x - c|)^2 + 1); used by S2_interpolate.py:41-42.
MOF_HD double mof_rbf_phi_body(const double* x, const double* c, double inv_eps) {
    const double d[3] = {MOF_ADD(x[0], -c[0]), MOF_ADD(x[1], -c[1]), MOF_ADD(x[2], -c[2])};
    const double s = MOF_MUL(inv_eps, sqrt(mof_dot3_plain(d, d)));
    return sqrt(MOF_ADD(MOF_MUL(s, s), 1.0));
}

// ---- K6 (S5 wave speed; S5 = S5_compute_wave_v.py under /root/reference) -------------
// angle_subtract, S5:224-233: np.mod(f1 - f2 + pi, 2 pi) - pi, result in [-pi, pi).  fmod is exact, so the three
// range branches ARE fmod plus numpy's sign fix for |d| < 4 pi (d - 2 pi is exact for d in [2 pi, 4 pi) by Sterbenz);
// inputs in [-pi, pi] never leave them.
MOF_HD double mof_angle_subtract_body(double a, double b) {
    const double kPi = 3.141592653589793, kTwoPi = 2.0 * 3.141592653589793;
    const double d = MOF_ADD(MOF_ADD(a, -b), kPi);
    double m;
    if (d > -kTwoPi && d < 2.0 * kTwoPi) {
        m = d;
        if (d < 0.0) m = MOF_ADD(d, kTwoPi);
        if (d >= kTwoPi) m = MOF_ADD(d, -kTwoPi);
    } else {
        m = fmod(d, kTwoPi);
        if (m != 0.0 && m < 0.0) m = MOF_ADD(m, kTwoPi);      // numpy's mod takes the sign of the divisor
    }
    return MOF_ADD(m, -kPi);
}

// Coefficient rows of vertex v, aligned with its block row j = rowptr[v] .. rowptr[v+1]-1 (column u = col[j]):
//   cg[j][3] = (1 / sum of A_f over the faces of v) * sum over the faces f holding v and u of A_f grad_w[f][position of u]
//              -> grad_point[v] = sum_j cg[j] I[u_j]                      (compute_grad_M_I, S5:136-171, faces ascending)
//   cw[j][2] = (alpha, beta) of cg[j] projected into the plane of (e1, e2) (project_vector_to_plane S5:173-180,
//              express_vector_on_basis S5:182-191) -> (alpha, beta)[v] = sum_j cw[j] I[u_j]
// Either output may be NULL.
MOF_HD void mof_wave_coef_row_body(const mof_mesh_dev& M, int64_t v, double* cw, double* cg) {
    const int32_t bd = M.diag[v];
    double asum = 0.0;                                                       // S5:169: every face of v is on its diagonal block
    for (int32_t q = M.cptr[bd]; q < M.cptr[bd + 1]; ++q) asum += M.areas[M.centry[q] >> 4];
    const double inv_asum = 1.0 / asum;
    const double* e1 = M.e + 6 * v;
    const double* e2 = e1 + 3;
    const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];   // S5:177
    const double inv_nn = 1.0 / (nx * nx + ny * ny + nz * nz);
    const double inv_e1 = 1.0 / (e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    const double inv_e2 = 1.0 / (e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
    for (int32_t j = M.rowptr[v]; j < M.rowptr[v + 1]; ++j) {
        double x = 0.0, y = 0.0, z = 0.0;
        for (int32_t q = M.cptr[j]; q < M.cptr[j + 1]; ++q) {
            const int32_t ce = M.centry[q];
            const int64_t f = ce >> 4;
            const double* gw = M.grad_w + 9 * f + 3 * (ce & 3);             // hat gradient of the column vertex in f (S5:154-158)
            const double A = M.areas[f];
            x += gw[0] * A; y += gw[1] * A; z += gw[2] * A;                  // S5:165
        }
        x *= inv_asum; y *= inv_asum; z *= inv_asum;
        if (cg) { cg[3 * (size_t)j] = x; cg[3 * (size_t)j + 1] = y; cg[3 * (size_t)j + 2] = z; }
        if (cw) {
            const double s = (x * nx + y * ny + z * nz) * inv_nn;
            const double px = x - s * nx, py = y - s * ny, pz = z - s * nz;
            cw[2 * (size_t)j] = (px * e1[0] + py * e1[1] + pz * e1[2]) * inv_e1;
            cw[2 * (size_t)j + 1] = (px * e2[0] + py * e2[1] + pz * e2[2]) * inv_e2;
        }
    }
}

// Time derivative of one (vertex, frame): wrapped differences for phases (compute_temporal_gradient_phase, S5:60-77)
// or np.gradient(axis=0, edge_order=2) / dt for amplitudes (S5:24).  first / last: the frame is the trial's first / last;
// far2: the value two frames inside the trial from that end (amplitude mode only).  x inv_dt is one rounding away from
// the reference's / dt.
MOF_HD double mof_wave_td_body(int phase_mode, bool first, bool last, int64_t T_trial, double cur, double prev, double next,
                               double far2, double inv_dt) {
    if (phase_mode) {
        if (T_trial <= 1) return 0.0;
        return mof_angle_subtract_body(last ? cur : next, first ? cur : prev) * (first || last ? inv_dt : 0.5 * inv_dt);
    }
    if (first) return (-1.5 * cur + 2.0 * next - 0.5 * far2) * inv_dt;
    if (last) return (1.5 * cur - 2.0 * prev + 0.5 * far2) * inv_dt;
    return ((next - prev) / 2.0) * inv_dt;
}

// S5:117,121 (S5:52,56): td / sqrt(alpha^2 + beta^2); 0 / 0 -> nan and x / 0 -> +-inf come out of 0 * inf and x * inf alike.
MOF_HD double mof_wave_speed_body(double td, double al, double be) {
#if defined(__CUDA_ARCH__)
    return td * rsqrt(fma(be, be, al * al));
#else
    return td * (1.0 / sqrt(fma(be, be, al * al)));
#endif
}

#endif  // MOF_BODIES_H

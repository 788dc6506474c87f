// K4 -- tangent coefficients -> xyz, speed and per-frame maximum speed
//       (process_V_k, utils/find_singularity_point.py:28-69; V_c,
//       S3_compute_v_and_detection_singularity.py:130-132; v_length_max, fsp:161-162).
// K5 -- singular vertices and faces holding an interior zero of the velocity field
//       (find_singularity_points, fsp:140-189), with an ordered stream compaction so the
//       output lists are in ascending vertex / face index like the reference's loops.
//
// All kernels here are one-pass, bandwidth-bound and run once per frame (the PCG runs
// ~10^3 passes per frame), so they use the plain one-thread-per-element mapping.
#include "mof_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;

// Maxima of non-negative doubles are taken on their bit patterns; +NaN orders above +inf,
// so a NaN speed makes the maximum NaN exactly like np.max (fsp:162).
__device__ __forceinline__ unsigned long long block_max_bits(unsigned long long v) {
    __shared__ unsigned long long s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(kFull, v, o);
        v = w > v ? w : v;
    }
    if (lane == 0) s[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x + 31) / 32 ? s[lane] : 0ull;
        for (int o = 16; o; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(kFull, v, o);
            v = w > v ? w : v;
        }
    }
    return v;   // valid in thread 0
}

__global__ void __launch_bounds__(256) tangent_kernel(int64_t N, const double* __restrict__ V, int64_t ldV,
                                                      const double* __restrict__ e, double* __restrict__ Vxyz,
                                                      double* __restrict__ speed, double* __restrict__ vmax) {
    const int64_t k = blockIdx.y;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long bits = 0ull;
    if (i < N) {
        double out[3];
        mof_tangent_body(V[k * ldV + i], V[k * ldV + N + i], e + 6 * i, out);
        double* dst = Vxyz + ((size_t)k * N + i) * 3;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2];
        const double len = mof_len3_body(out);
        if (speed) speed[(size_t)k * N + i] = len;
        bits = (unsigned long long)__double_as_longlong(len);
    }
    if (vmax) {
        bits = block_max_bits(bits);
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(vmax + k), bits);
    }
}

__global__ void __launch_bounds__(256) vmax_kernel(int64_t N, const double* __restrict__ Vxyz, double* __restrict__ vmax) {
    const int64_t k = blockIdx.y;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long bits = 0ull;
    if (i < N) bits = (unsigned long long)__double_as_longlong(mof_len3_body(Vxyz + ((size_t)k * N + i) * 3));
    bits = block_max_bits(bits);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(vmax + k), bits);
}

// ---- pass 1 -----------------------------------------------------------------------
__global__ void __launch_bounds__(MOF_DETECT_CHUNK) vflag_kernel(int64_t N, const double* __restrict__ Vxyz,
                                                                const double* __restrict__ vmax, double eps,
                                                                uint8_t* __restrict__ vflag, int32_t* __restrict__ vcnt) {
    const int64_t k = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * MOF_DETECT_CHUNK + threadIdx.x;
    int flag = 0;
    if (i < N) {
        flag = mof_vertex_zero_body(Vxyz + ((size_t)k * N + i) * 3, vmax[k], eps) ? 1 : 0;
        vflag[(size_t)k * N + i] = (uint8_t)flag;
    }
    const int c = __syncthreads_count(flag);
    if (threadIdx.x == 0) vcnt[(size_t)k * gridDim.x + blockIdx.x] = c;
}

__device__ __forceinline__ bool face_test(int64_t N, int64_t k, int64_t t, const double* __restrict__ coords,
                                          const int32_t* __restrict__ tri, const double* __restrict__ Vxyz,
                                          double vmax, double* lam, double* mu, int* sign, int64_t* abc) {
    const int64_t a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    abc[0] = a; abc[1] = b; abc[2] = c;
    const double* Vk = Vxyz + (size_t)k * N * 3;
    return mof_face_zero_body(coords + 3 * a, coords + 3 * b, coords + 3 * c, Vk + 3 * a, Vk + 3 * b, Vk + 3 * c,
                              vmax, lam, mu, sign);
}

__global__ void __launch_bounds__(MOF_DETECT_CHUNK) fflag_kernel(int64_t N, int64_t F, const double* __restrict__ coords,
                                                                const int32_t* __restrict__ tri,
                                                                const double* __restrict__ Vxyz,
                                                                const double* __restrict__ vmax,
                                                                const uint8_t* __restrict__ vflag,
                                                                uint8_t* __restrict__ fflag, int32_t* __restrict__ fcnt) {
    const int64_t k = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * MOF_DETECT_CHUNK + threadIdx.x;
    int flag = 0;
    if (t < F) {
        const uint8_t* vf = vflag + (size_t)k * N;
        const bool skip = vf[tri[3 * t]] | vf[tri[3 * t + 1]] | vf[tri[3 * t + 2]];      // fsp:171-172
        if (!skip) {
            double lam, mu; int sign; int64_t abc[3];
            flag = face_test(N, k, t, coords, tri, Vxyz, vmax[k], &lam, &mu, &sign, abc) ? 1 : 0;
        }
        fflag[(size_t)k * F + t] = (uint8_t)flag;
    }
    const int c = __syncthreads_count(flag);
    if (threadIdx.x == 0) fcnt[(size_t)k * gridDim.x + blockIdx.x] = c;
}

// per frame: exclusive scan of the chunk counts in place, totals out
__global__ void scan_kernel(int64_t n_frames, int nvc, int nfc, int32_t* __restrict__ vcnt, int32_t* __restrict__ fcnt,
                            int32_t* __restrict__ totals) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_frames) return;
    int32_t run = 0;
    for (int c = 0; c < nvc; ++c) { int32_t t = vcnt[k * nvc + c]; vcnt[k * nvc + c] = run; run += t; }
    totals[2 * k] = run;
    run = 0;
    for (int c = 0; c < nfc; ++c) { int32_t t = fcnt[k * nfc + c]; fcnt[k * nfc + c] = run; run += t; }
    totals[2 * k + 1] = run;
}

// ---- pass 2 -----------------------------------------------------------------------
// rank of this thread among the flagged threads of the CTA (ascending thread index)
__device__ __forceinline__ int block_rank(int flag) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(kFull, flag);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    return base + __popc(bal & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(MOF_DETECT_CHUNK) vcompact_kernel(int64_t N, const uint8_t* __restrict__ vflag,
                                                                   const int32_t* __restrict__ vcnt,
                                                                   const int64_t* __restrict__ voff,
                                                                   int32_t* __restrict__ vertex_idx) {
    const int64_t k = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * MOF_DETECT_CHUNK + threadIdx.x;
    const int flag = i < N ? vflag[(size_t)k * N + i] : 0;
    const int rank = block_rank(flag);
    if (flag) vertex_idx[voff[k] + vcnt[(size_t)k * gridDim.x + blockIdx.x] + rank] = (int32_t)i;
}

__global__ void __launch_bounds__(MOF_DETECT_CHUNK) fcompact_kernel(int64_t N, int64_t F, const double* __restrict__ coords,
                                                                   const int32_t* __restrict__ tri,
                                                                   const double* __restrict__ Vxyz,
                                                                   const double* __restrict__ vmax,
                                                                   const uint8_t* __restrict__ fflag,
                                                                   const int32_t* __restrict__ fcnt,
                                                                   const int64_t* __restrict__ foff,
                                                                   int32_t* __restrict__ face_idx, double* __restrict__ lam_mu,
                                                                   double* __restrict__ P, int8_t* __restrict__ index) {
    const int64_t k = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * MOF_DETECT_CHUNK + threadIdx.x;
    const int flag = t < F ? fflag[(size_t)k * F + t] : 0;
    const int rank = block_rank(flag);
    if (!flag) return;
    double lam, mu; int sign; int64_t abc[3];
    face_test(N, k, t, coords, tri, Vxyz, vmax[k], &lam, &mu, &sign, abc);
    const int64_t o = foff[k] + fcnt[(size_t)k * gridDim.x + blockIdx.x] + rank;
    face_idx[o] = (int32_t)t;
    lam_mu[2 * o] = lam;
    lam_mu[2 * o + 1] = mu;
    const double nu = MOF_ADD(MOF_ADD(1.0, -lam), -mu);                       // fsp:182: (1 - lam - mu)
    for (int c = 0; c < 3; ++c)                                               // fsp:181-182
        P[3 * o + c] = MOF_ADD(MOF_ADD(MOF_MUL(lam, coords[3 * abc[0] + c]), MOF_MUL(mu, coords[3 * abc[1] + c])),
                               MOF_MUL(nu, coords[3 * abc[2] + c]));
    if (index) index[o] = (int8_t)sign;
}

// ---- Jacobian classification of the detected critical points (fsp:355-498, 561-605) ----------
// One thread per point (tens of points per frame).  Vertices use the 1-ring as near points and the
// vertex's own tangent basis; interior points use the vertices of their face plus those of the face
// across the nearest edge, in ascending vertex order, and a basis built from the face normal.
__device__ __forceinline__ int64_t frame_of(const int64_t* __restrict__ off, int64_t n_frames, int64_t q) {
    int64_t lo = 0, hi = n_frames;                 // off[lo] <= q < off[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void classify_vertex_kernel(int64_t N, int64_t n_frames, const double* __restrict__ coords,
                                       const double* __restrict__ Vxyz, const double* __restrict__ vmax,
                                       const double* __restrict__ e, const int32_t* __restrict__ ring_ptr,
                                       const int32_t* __restrict__ ring_idx, const int64_t* __restrict__ voff,
                                       const int32_t* __restrict__ vertex_idx, double* __restrict__ jac,
                                       int8_t* __restrict__ cls) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= voff[n_frames]) return;
    const int64_t k = frame_of(voff, n_frames, q);
    const int64_t i = vertex_idx[q];
    const double* Vk = Vxyz + (size_t)k * N * 3;
    double J[4] = {0.0, 0.0, 0.0, 0.0};
    for (int32_t r = ring_ptr[i]; r < ring_ptr[i + 1]; ++r) {
        const int64_t nb = ring_idx[r];
        mof_jacobian_term_body(coords + 3 * i, coords + 3 * nb, Vk + 3 * nb, vmax[k], e + 6 * i, e + 6 * i + 3, J);
    }
    for (int c = 0; c < 4; ++c) jac[4 * q + c] = J[c];
    cls[q] = (int8_t)mof_classify_body(J);
}

__global__ void classify_face_kernel(int64_t N, int64_t n_frames, const double* __restrict__ coords,
                                     const int32_t* __restrict__ tri, const double* __restrict__ Vxyz,
                                     const double* __restrict__ vmax, const int32_t* __restrict__ face_nbr,
                                     const int64_t* __restrict__ foff, const int32_t* __restrict__ face_idx,
                                     const double* __restrict__ P, double* __restrict__ jac, int8_t* __restrict__ cls) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= foff[n_frames]) return;
    const int64_t k = frame_of(foff, n_frames, q);
    const int64_t t = face_idx[q];
    const int64_t a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    const double *A = coords + 3 * a, *B = coords + 3 * b, *C = coords + 3 * c;
    const double ab[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, ac[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
    const double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= nl; n[1] /= nl; n[2] /= nl;                                           // calculate_normal, fsp:285-289
    double eb[6];
    mof_basis_body(n, eb);                                                        // compute_orthonormal_basis, fsp:304-314
    const double* Pq = P + 3 * q;
    const int edge = mof_nearest_edge_body(A, B, C, Pq);
    const int64_t other = face_nbr[3 * t + edge];
    int64_t near[6] = {a, b, c, -1, -1, -1};
    int cnt = 3;
    if (other >= 0)
        for (int m = 0; m < 3; ++m) {
            const int64_t v = tri[3 * other + m];
            bool dup = false;
            for (int z = 0; z < cnt; ++z) dup |= near[z] == v;
            if (!dup) near[cnt++] = v;
        }
    for (int x = 1; x < cnt; ++x)                                                  // ascending vertex order
        for (int y = x; y > 0 && near[y - 1] > near[y]; --y) { int64_t w = near[y]; near[y] = near[y - 1]; near[y - 1] = w; }
    const double* Vk = Vxyz + (size_t)k * N * 3;
    double J[4] = {0.0, 0.0, 0.0, 0.0};
    for (int z = 0; z < cnt; ++z)
        mof_jacobian_term_body(Pq, coords + 3 * near[z], Vk + 3 * near[z], vmax[k], eb, eb + 3, J);
    for (int m = 0; m < 4; ++m) jac[4 * q + m] = J[m];
    cls[q] = (int8_t)mof_classify_body(J);
}

}  // namespace

extern "C" int mof_classify_singularities(int64_t N, int64_t F, int64_t n_frames, const double* coords, const int32_t* tri,
                                          const double* Vxyz, const double* vmax, const double* e,
                                          const int32_t* ring_ptr, const int32_t* ring_idx, const int32_t* face_nbr,
                                          const int64_t* voff, const int64_t* foff, int64_t n_vertex_points,
                                          int64_t n_face_points, const int32_t* vertex_idx, const int32_t* face_idx,
                                          const double* P, double* jac_v, int8_t* cls_v, double* jac_f, int8_t* cls_f,
                                          void* stream) {
    MOF_REQUIRE(N > 0 && F >= 0 && n_frames > 0 && coords && tri && Vxyz && vmax && e && voff && foff, "bad arguments");
    cudaStream_t st = mof_stream(stream);
    if (n_vertex_points > 0) {
        MOF_REQUIRE(ring_ptr && ring_idx && vertex_idx && jac_v && cls_v, "vertex arguments missing");
        classify_vertex_kernel<<<mof_cdiv(n_vertex_points, 128), 128, 0, st>>>(N, n_frames, coords, Vxyz, vmax, e, ring_ptr,
                                                                             ring_idx, voff, vertex_idx, jac_v, cls_v);
        MOF_LAUNCH_CHECK("classify_vertex_kernel");
    }
    if (n_face_points > 0) {
        MOF_REQUIRE(face_nbr && face_idx && P && jac_f && cls_f, "face arguments missing");
        classify_face_kernel<<<mof_cdiv(n_face_points, 128), 128, 0, st>>>(N, n_frames, coords, tri, Vxyz, vmax, face_nbr, foff,
                                                                         face_idx, P, jac_f, cls_f);
        MOF_LAUNCH_CHECK("classify_face_kernel");
    }
    return 0;
}

extern "C" int mof_tangent_to_xyz(int64_t N, int64_t n_frames, const double* V, int64_t ldV, const double* e,
                                  double* Vxyz, double* speed, double* vmax, void* stream) {
    MOF_REQUIRE(N > 0 && n_frames >= 0 && V && e && Vxyz && ldV >= 2 * N, "bad arguments");
    MOF_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
    if (n_frames == 0) return 0;
    if (vmax) MOF_CUDA_TRY(cudaMemsetAsync(vmax, 0, n_frames * sizeof(double), mof_stream(stream)));
    dim3 grid(mof_cdiv(N, 256), (unsigned)n_frames);
    tangent_kernel<<<grid, 256, 0, mof_stream(stream)>>>(N, V, ldV, e, Vxyz, speed, vmax);
    MOF_LAUNCH_CHECK("tangent_kernel");
    return 0;
}

extern "C" int mof_vmax(int64_t N, int64_t n_frames, const double* Vxyz, double* vmax, void* stream) {
    MOF_REQUIRE(N > 0 && n_frames >= 0 && Vxyz && vmax, "bad arguments");
    MOF_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
    if (n_frames == 0) return 0;
    MOF_CUDA_TRY(cudaMemsetAsync(vmax, 0, n_frames * sizeof(double), mof_stream(stream)));
    dim3 grid(mof_cdiv(N, 256), (unsigned)n_frames);
    vmax_kernel<<<grid, 256, 0, mof_stream(stream)>>>(N, Vxyz, vmax);
    MOF_LAUNCH_CHECK("vmax_kernel");
    return 0;
}

extern "C" int mof_singularity_flags(int64_t N, int64_t F, int64_t n_frames, const double* coords, const int32_t* tri,
                                     const double* Vxyz, const double* vmax, double eps, uint8_t* vflag,
                                     uint8_t* fflag, int32_t* vcnt, int32_t* fcnt, int32_t* totals, void* stream) {
    MOF_REQUIRE(N > 0 && F >= 0 && n_frames >= 0 && coords && tri && Vxyz && vmax && vflag && fflag && vcnt && fcnt && totals,
                "bad arguments");
    MOF_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
    if (n_frames == 0) return 0;
    cudaStream_t st = mof_stream(stream);
    const unsigned nvc = mof_cdiv(N, MOF_DETECT_CHUNK), nfc = mof_cdiv(F, MOF_DETECT_CHUNK);
    vflag_kernel<<<dim3(nvc, (unsigned)n_frames), MOF_DETECT_CHUNK, 0, st>>>(N, Vxyz, vmax, eps, vflag, vcnt);
    MOF_LAUNCH_CHECK("vflag_kernel");
    if (F > 0) {
        fflag_kernel<<<dim3(nfc, (unsigned)n_frames), MOF_DETECT_CHUNK, 0, st>>>(N, F, coords, tri, Vxyz, vmax, vflag, fflag, fcnt);
        MOF_LAUNCH_CHECK("fflag_kernel");
    }
    scan_kernel<<<mof_cdiv(n_frames, 128), 128, 0, st>>>(n_frames, (int)nvc, (int)nfc, vcnt, fcnt, totals);
    MOF_LAUNCH_CHECK("scan_kernel");
    return 0;
}

extern "C" int mof_singularity_compact(int64_t N, int64_t F, int64_t n_frames, const double* coords, const int32_t* tri,
                                       const double* Vxyz, const double* vmax, const uint8_t* vflag,
                                       const uint8_t* fflag, const int32_t* vcnt, const int32_t* fcnt,
                                       const int64_t* voff, const int64_t* foff, int32_t* vertex_idx,
                                       int32_t* face_idx, double* lam_mu, double* P, int8_t* index, void* stream) {
    MOF_REQUIRE(N > 0 && F >= 0 && n_frames >= 0 && coords && tri && Vxyz && vmax && vflag && fflag && vcnt && fcnt && voff && foff,
                "bad arguments");
    MOF_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
    if (n_frames == 0) return 0;
    cudaStream_t st = mof_stream(stream);
    const unsigned nvc = mof_cdiv(N, MOF_DETECT_CHUNK), nfc = mof_cdiv(F, MOF_DETECT_CHUNK);
    if (vertex_idx) {
        vcompact_kernel<<<dim3(nvc, (unsigned)n_frames), MOF_DETECT_CHUNK, 0, st>>>(N, vflag, vcnt, voff, vertex_idx);
        MOF_LAUNCH_CHECK("vcompact_kernel");
    }
    if (F > 0 && face_idx) {
        MOF_REQUIRE(lam_mu && P, "lam_mu / P missing");
        fcompact_kernel<<<dim3(nfc, (unsigned)n_frames), MOF_DETECT_CHUNK, 0, st>>>(N, F, coords, tri, Vxyz, vmax, fflag, fcnt,
                                                                                    foff, face_idx, lam_mu, P, index);
        MOF_LAUNCH_CHECK("fcompact_kernel");
    }
    return 0;
}

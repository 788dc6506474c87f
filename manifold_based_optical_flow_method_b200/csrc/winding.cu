// K7 -- multi-ring winding numbers of the detected singular points ("next" row 3 of SURVEY.md 8f):
//       calculate_winding_numbers, S7_winding_line.py:120-165, with angle_between_vectors /
//       winding_number (:59-87) and the polar-angle ordering (:93-102).
//
// One CTA per singular point.  The CTA
//   1. finds the closest mesh vertex (S7:130; block argmin, lowest index on ties),
//   2. grows breadth-first rings around it (S7:131, pyvista point_neighbors_levels) with a visited
//      bitmask of N bits in shared memory (atomicOr = test-and-set, so a ring needs no dedupe pass),
//   3. per ring: tangent-plane polar key and velocity components of every ring vertex, a rank
//      sort by (key, vertex id) -- the reference's stable lexsort of an ascending ring -- the
//      signed turning angles between cyclic neighbours, their sequential sum / 2 pi,
//   4. applies the reference's acceptance rule and stops at the first ring that fails it.
// Rings hold at most kRingCap vertices (25 rings of a cortical mesh hold ~150-300); a larger
// ring sets status 2 for the point and the Python layer raises.  This runs once per frame on a
// few dozen points: it is latency-, not bandwidth-bound, and is kept simple.
#include "mof_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRingCap = 1024;
constexpr size_t kFixedSmem = (size_t)kRingCap * (2 * sizeof(int32_t) + 5 * sizeof(double));

__global__ void __launch_bounds__(kThreads) winding_kernel(
    int64_t N, const double* __restrict__ coords, const double* __restrict__ Vxyz, const double* __restrict__ e,
    const int32_t* __restrict__ ring_ptr, const int32_t* __restrict__ ring_idx, const double* __restrict__ points,
    const int32_t* __restrict__ frame_of_point, int max_level, int mask_words, int32_t* __restrict__ closest,
    int32_t* __restrict__ counts, int8_t* __restrict__ types, double* __restrict__ winding, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);          // polar keys, later the turning angles
    double* vx = key + kRingCap;
    double* vy = vx + kRingCap;
    double* svx = vy + kRingCap;                                 // components in sorted order
    double* svy = svx + kRingCap;
    int32_t* ring_a = reinterpret_cast<int32_t*>(svy + kRingCap);
    int32_t* ring_b = ring_a + kRingCap;
    unsigned* mask = reinterpret_cast<unsigned*>(ring_b + kRingCap);

    __shared__ double s_best[kThreads / 32];
    __shared__ int s_arg[kThreads / 32];
    __shared__ int s_index, s_next, s_go;

    const int tid = threadIdx.x;
    const int64_t p = blockIdx.x;
    const double P[3] = {points[3 * p], points[3 * p + 1], points[3 * p + 2]};
    const double* V = Vxyz + (size_t)frame_of_point[p] * (size_t)N * 3;

    for (int w = tid; w < mask_words; w += kThreads) mask[w] = 0u;
    if (winding) for (int l = tid; l < max_level; l += kThreads) winding[p * max_level + l] = nan("");

    // 1. closest vertex: (distance, index) lexicographic minimum; NaN distances never win
    double best = INFINITY;
    int arg = 0x7fffffff;
    for (int64_t v = tid; v < N; v += kThreads) {
        const double d = mof_dist3_body(coords + 3 * v, P);
        if (d < best) { best = d; arg = (int)v; }
    }
    for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_arg[tid >> 5] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w)
            if (s_best[w] < best || (s_best[w] == best && s_arg[w] < arg)) { best = s_best[w]; arg = s_arg[w]; }
        if (arg == 0x7fffffff) arg = 0;                          // every distance NaN: np.argmin returns the first NaN
        s_index = arg;
        s_next = 0;
        mask[arg >> 5] |= 1u << (arg & 31);
        ring_a[0] = arg;
    }
    __syncthreads();
    const int index = s_index;
    const double O[3] = {coords[3 * (size_t)index], coords[3 * (size_t)index + 1], coords[3 * (size_t)index + 2]};
    double e1[3], e2[3];
    for (int c = 0; c < 3; ++c) { e1[c] = e[6 * (size_t)index + c]; e2[c] = e[6 * (size_t)index + 3 + c]; }

    int32_t* cur = ring_a;
    int32_t* nxt = ring_b;
    int n_cur = 1, count = 0, flag = 0, stat = 0;
    for (int level = 0; level < max_level; ++level) {
        // 2. next ring = unvisited neighbours of the current ring
        for (int t = tid; t < n_cur; t += kThreads) {
            const int v = cur[t];
            for (int j = ring_ptr[v]; j < ring_ptr[v + 1]; ++j) {
                const int w = ring_idx[j];
                const unsigned bit = 1u << (w & 31);
                if (!(atomicOr(&mask[w >> 5], bit) & bit)) {
                    const int pos = atomicAdd(&s_next, 1);
                    if (pos < kRingCap) nxt[pos] = w;
                }
            }
        }
        __syncthreads();
        const int n = s_next;
        if (n == 0) break;                                       // mesh exhausted (the reference raises here)
        if (n > kRingCap) { stat = 2; break; }
        // 3a. per ring vertex: polar key and tangent components of its velocity
        for (int t = tid; t < n; t += kThreads) {
            const size_t w = (size_t)nxt[t];
            mof_winding_element_body(O, coords + 3 * w, V + 3 * w, e1, e2, &key[t], &vx[t], &vy[t]);
        }
        __syncthreads();
        // 3b. rank sort by (key, vertex id)
        for (int t = tid; t < n; t += kThreads) {
            const double kt = key[t];
            const int wt = nxt[t];
            int r = 0;
            for (int j = 0; j < n; ++j) r += (key[j] < kt || (key[j] == kt && nxt[j] < wt)) ? 1 : 0;
            svx[r] = vx[t];
            svy[r] = vy[t];
        }
        __syncthreads();
        // 3c. turning angle from sorted position t to its cyclic successor
        for (int t = tid; t < n; t += kThreads) {
            const int u = t + 1 == n ? 0 : t + 1;
            key[t] = mof_signed_angle_body(svx[t], svy[t], svx[u], svy[u]);
        }
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int t = 0; t < n; ++t) sum = MOF_ADD(sum, key[t]);          // same order as S7:82-86
            const double w = sum / (2 * 3.141592653589793);
            if (winding) winding[p * max_level + level] = w;
            int f = flag;
            const bool ok = mof_winding_accept_body(level, w, &f);
            s_go = (ok ? 1 : 0) | ((f & 3) << 1);
            s_next = 0;
        }
        __syncthreads();
        const int go = s_go;
        if (level == 0) flag = ((go >> 1) & 3) == 3 ? -1 : ((go >> 1) & 3);
        if (!(go & 1)) break;
        ++count;
        int32_t* tmp = cur; cur = nxt; nxt = tmp;
        n_cur = n;
        __syncthreads();                                         // s_go / s_next are rewritten in the next level
    }
    if (tid == 0) {
        closest[p] = index;
        counts[p] = count;
        types[p] = (int8_t)flag;
        status[p] = stat;
    }
}

}  // namespace

extern "C" int mof_winding_numbers(int64_t N, int64_t n_frames, const double* coords, const double* Vxyz, const double* e,
                                   const int32_t* ring_ptr, const int32_t* ring_idx, int64_t n_points, const double* points,
                                   const int32_t* frame_of_point, int max_level, int32_t* closest, int32_t* counts,
                                   int8_t* types, double* winding, int32_t* status, void* stream) {
    MOF_REQUIRE(N > 0 && N < 0x7fffffff && n_frames > 0 && n_points >= 0 && max_level > 0, "bad sizes");
    if (n_points == 0) return 0;
    MOF_REQUIRE(coords && Vxyz && e && ring_ptr && ring_idx && points && frame_of_point && closest && counts && types && status,
                "null argument");
    MOF_REQUIRE(n_points <= 0x7fffffff, "too many points for one call");
    const int mask_words = (int)((N + 31) / 32);
    const size_t smem = kFixedSmem + (size_t)mask_words * sizeof(unsigned);
    MOF_REQUIRE(smem <= 227 * 1024, "mesh too large for the shared-memory visited mask (about 1.4 million vertices)");
    MOF_CUDA_TRY(cudaFuncSetAttribute(winding_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    winding_kernel<<<(unsigned)n_points, kThreads, smem, mof_stream(stream)>>>(
        N, coords, Vxyz, e, ring_ptr, ring_idx, points, frame_of_point, max_level, mask_words, closest, counts, types, winding,
        status);
    MOF_LAUNCH_CHECK("winding_kernel");
    return 0;
}

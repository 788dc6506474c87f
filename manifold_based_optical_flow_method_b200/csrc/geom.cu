// K0 -- once-per-mesh geometry: tangent bases, hat-function gradients, mass integrals and
// the frame-independent a2 block values (compute_geometrical_quantities,
// utils/compute_optical_flow.py:27-97).  One thread per vertex / face / block; the work is
// a few hundred KB..MB once per mesh, so these kernels are written for clarity.
#include "mof_common.cuh"

namespace {

__global__ void basis_kernel(int64_t N, const double* __restrict__ normals, double* __restrict__ e) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v < N) mof_basis_body(normals + 3 * v, e + 6 * v);
}

__global__ void gradw_kernel(int64_t F, const double* __restrict__ coords, const int32_t* __restrict__ tri,
                             const double* __restrict__ areas, double* __restrict__ grad_w,
                             double* __restrict__ integral) {
    int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f < F) mof_face_geom_body(coords, tri, areas, f, grad_w, integral);
}

// thread per block; the row of a block is found by binary search in rowptr
__global__ void a2_kernel(mof_mesh_dev M, double* __restrict__ a2v) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= M.n_blocks) return;
    int64_t lo = 0, hi = M.n_vertices;            // rowptr[lo] <= b < rowptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (M.rowptr[mid] <= b) lo = mid; else hi = mid;
    }
    double out[4];
    mof_a2_block_body(M, lo, b, out);
    for (int c = 0; c < 4; ++c) a2v[4 * b + c] = out[c];
}

}  // namespace

extern "C" int mof_geom_basis(int64_t N, const double* normals, double* e, void* stream) {
    MOF_REQUIRE(N > 0 && normals && e, "bad arguments");
    basis_kernel<<<mof_cdiv(N, 256), 256, 0, mof_stream(stream)>>>(N, normals, e);
    MOF_LAUNCH_CHECK("basis_kernel");
    return 0;
}

extern "C" int mof_geom_gradw(int64_t F, const double* coords, const int32_t* tri, const double* areas,
                              double* grad_w, double* integral, void* stream) {
    MOF_REQUIRE(F >= 0 && coords && tri && areas && grad_w && integral, "bad arguments");
    if (F == 0) return 0;
    gradw_kernel<<<mof_cdiv(F, 256), 256, 0, mof_stream(stream)>>>(F, coords, tri, areas, grad_w, integral);
    MOF_LAUNCH_CHECK("gradw_kernel");
    return 0;
}

extern "C" int mof_geom_a2(const mof_mesh_dev* mesh, double* a2v, void* stream) {
    MOF_REQUIRE(mesh && a2v && mesh->e && mesh->grad_w && mesh->areas, "bad arguments");
    a2_kernel<<<mof_cdiv(mesh->n_blocks, 256), 256, 0, mof_stream(stream)>>>(*mesh, a2v);
    MOF_LAUNCH_CHECK("a2_kernel");
    return 0;
}

// K6 -- wave speed from phase / amplitude maps (S5_compute_wave_v.py): per vertex and frame
//   grad_point = area-weighted mean of the per-face surface gradients of the incident faces
//                (compute_grad_M_I, S5:136-171, same hat-function "gradients" as cof:238-255),
//   its tangent-plane projection expressed in the basis (e1, e2) and the norm of that
//   (project_vector_to_plane S5:173-180, express_vector_on_basis S5:182-191, S5:96-117),
//   the time derivative: wrapped differences for phases (compute_temporal_gradient_phase,
//   S5:60-77, angle_subtract S5:224-233) or np.gradient(edge_order=2) for amplitudes (S5:24),
//   wave_velocity = time derivative / norm (S5:121 / S5:56).
// One thread per (vertex, frame); vertex-contiguous threads read a frame row of the (T,N)
// signal through L2.  One pass, bandwidth-bound: read 8 N (+ gathers), write 8 N (+ 24 N) per frame.
#include "mof_common.cuh"

namespace {

__device__ __forceinline__ double angle_subtract(double a, double b) {
    // np.mod(f1 - f2 + pi, 2 pi) - pi : result in [-pi, pi)            (S5:230)
    const double kPi = 3.141592653589793, kTwoPi = 2.0 * 3.141592653589793;
    double d = a - b + kPi;
    double m = fmod(d, kTwoPi);
    if (m != 0.0 && m < 0.0) m += kTwoPi;      // numpy's mod takes the sign of the divisor
    return m - kPi;
}

__global__ void __launch_bounds__(256) wave_speed_kernel(mof_mesh_dev M, int64_t T, const double* __restrict__ I, int64_t ld,
                                                         double dt, int phase_mode, double* __restrict__ grad_point,
                                                         double* __restrict__ wave) {
    const int64_t N = M.n_vertices;
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // internal vertex
    const int64_t t = blockIdx.y;
    if (v >= N) return;
    const int64_t o = M.perm[v];                                            // reference vertex
    const double* It = I + t * ld;
    // area-weighted mean of the face gradients, faces ascending (S5:161-169)
    double gx = 0.0, gy = 0.0, gz = 0.0, area_sum = 0.0;
    const int32_t bd = M.diag[v];
    for (int32_t q = M.cptr[bd]; q < M.cptr[bd + 1]; ++q) {
        const int64_t f = M.centry[q] >> 4;
        const double* g = M.grad_w + 9 * f;
        const double I0 = It[M.perm[M.tri[3 * f]]], I1 = It[M.perm[M.tri[3 * f + 1]]], I2 = It[M.perm[M.tri[3 * f + 2]]];
        const double A = M.areas[f];
        gx += (I0 * g[0] + I1 * g[3] + I2 * g[6]) * A;                      // S5:154-158,165
        gy += (I0 * g[1] + I1 * g[4] + I2 * g[7]) * A;
        gz += (I0 * g[2] + I1 * g[5] + I2 * g[8]) * A;
        area_sum += A;
    }
    gx /= area_sum; gy /= area_sum; gz /= area_sum;                         // S5:169
    if (grad_point) {
        double* gp = grad_point + ((size_t)t * N + o) * 3;
        gp[0] = gx; gp[1] = gy; gp[2] = gz;
    }
    if (!wave) return;
    const double* e1 = M.e + 6 * v;
    const double* e2 = e1 + 3;
    // project_vector_to_plane (S5:173-180)
    const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    const double s = (gx * nx + gy * ny + gz * nz) / (nx * nx + ny * ny + nz * nz);
    const double px = gx - s * nx, py = gy - s * ny, pz = gz - s * nz;
    // express_vector_on_basis (S5:182-191) and its norm (S5:117)
    const double al = (px * e1[0] + py * e1[1] + pz * e1[2]) / (e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    const double be = (px * e2[0] + py * e2[1] + pz * e2[2]) / (e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
    const double dis = sqrt(al * al + be * be);
    // time derivative
    double td;
    const double c = I[t * ld + o];
    if (phase_mode) {                                                       // S5:60-77
        if (T == 1) td = 0.0;
        else if (t == 0) td = angle_subtract(I[ld + o], c) / dt;
        else if (t == T - 1) td = angle_subtract(c, I[(t - 1) * ld + o]) / dt;
        else td = angle_subtract(I[(t + 1) * ld + o], I[(t - 1) * ld + o]) / (2 * dt);
    } else {                                                                // np.gradient(axis=0, edge_order=2) / dt, S5:24
        if (t == 0) td = (-1.5 * c + 2.0 * I[ld + o] - 0.5 * I[2 * ld + o]) / dt;
        else if (t == T - 1) td = (1.5 * c - 2.0 * I[(t - 1) * ld + o] + 0.5 * I[(t - 2) * ld + o]) / dt;
        else td = ((I[(t + 1) * ld + o] - I[(t - 1) * ld + o]) / 2.0) / dt;
    }
    wave[(size_t)t * N + o] = td / dis;                                     // S5:121
}

}  // namespace

extern "C" int mof_wave_speed(const mof_mesh_dev* mesh, int64_t n_frames, const double* I, int64_t ld, double dt,
                              int phase_mode, double* grad_point, double* wave, void* stream) {
    MOF_REQUIRE(mesh && I && n_frames >= 0 && ld >= mesh->n_vertices && dt != 0.0, "bad arguments");
    MOF_REQUIRE(grad_point || wave, "nothing to compute");
    MOF_REQUIRE(phase_mode || !wave || n_frames >= 3, "np.gradient(edge_order=2) needs at least 3 frames");
    MOF_REQUIRE(n_frames <= 65535, "at most 65535 frames per call");
    if (n_frames == 0) return 0;
    dim3 grid(mof_cdiv(mesh->n_vertices, 256), (unsigned)n_frames);
    wave_speed_kernel<<<grid, 256, 0, mof_stream(stream)>>>(*mesh, n_frames, I, ld, dt, phase_mode, grad_point, wave);
    MOF_LAUNCH_CHECK("wave_speed_kernel");
    return 0;
}

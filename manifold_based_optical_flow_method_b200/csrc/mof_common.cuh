// Shared device/host helpers for the sm_100a kernels of libmof_b200.
#ifndef MOF_COMMON_CUH
#define MOF_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "mof_b200.h"
#include "mof_bodies.h"
#include "mof_error.h"

#define MOF_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return mof_set_error(-100 - (int)_e, "%s:%d: %s failed: %s", __FILE__, __LINE__,    \
                                 #expr, cudaGetErrorString(_e));                                \
    } while (0)

#define MOF_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return mof_set_error(-100 - (int)_e, "launch of %s failed: %s", name,               \
                                 cudaGetErrorString(_e));                                       \
    } while (0)

#define MOF_REQUIRE(cond, msg)                                                                  \
    do {                                                                                        \
        if (!(cond)) return mof_set_error(-1, "%s: %s", __func__, msg);                         \
    } while (0)

static inline cudaStream_t mof_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned mof_cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// scal[g][MOF_S_*][32]
//   RZ r'z, PAP p'Ap, RR r'r (recurrence), BB ||r0||^2 of the iterated (possibly transformed) system,
//   ALPHA/BETA/ZS step scalars (frozen frame: 0 / 1 / 0), RRTRUE ||b - A x||^2, BBT ||b||^2,
//   THR per-frame threshold on RR/BB, BETA_SAVED beta of a frozen frame for an exact resume
enum { MOF_S_RZ = 0, MOF_S_PAP = 1, MOF_S_RR = 2, MOF_S_BB = 3, MOF_S_ALPHA = 4, MOF_S_BETA = 5,
       MOF_S_RRTRUE = 6, MOF_S_BBT = 7, MOF_S_THR = 8, MOF_S_ZS = 9, MOF_S_BETA_SAVED = 10, MOF_S_SPARE = 11,
       // level path: exact fixed-point accumulator of p'Ap (128-bit two's complement in two 64-bit words), its
       // binary scale (int64) and an overflow / non-finite flag (int64); the slots hold bit patterns, not doubles
       MOF_S_FX_LO = 12, MOF_S_FX_HI = 13, MOF_S_FX_K = 14, MOF_S_FX_BAD = 15,
       // the same accumulator serves r'r in the update phase (the phases alternate); its own binary scale:
       MOF_S_FX_K_RR = 16, MOF_S_SPARE2 = 17, MOF_S_SPARE3 = 18, MOF_S_SPARE4 = 19,
       MOF_S_COUNT = 20 };
static_assert(MOF_S_COUNT == MOF_SCAL_SLOTS, "scalar slots of the header and the kernels differ");
// state[g][MOF_I_*][32], then group_done[G], ticket[G], groups_active[1]
enum { MOF_I_ACTIVE = 0, MOF_I_ITERS = 1, MOF_I_STATUS = 2, MOF_I_SPARE = 3, MOF_I_COUNT = 4 };
#define MOF_STATUS_PENDING (-1)

#endif  // MOF_COMMON_CUH

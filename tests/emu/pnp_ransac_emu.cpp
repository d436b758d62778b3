// TEST INFRASTRUCTURE ONLY (built by tests/test_pnp_ransac_cpu.py into a temporary directory, never into libgdrn_b200.so).
//
// Compiles gdr_net_b200/csrc/pnp_ransac.cu for the HOST with ONE thread per CTA so the kernel logic (gather order, RNG replay,
// EPnP, RANSAC bookkeeping) can be single-stepped against cv2 in the CPU test suite, before any GPU minute is spent.  The
// CUDA built-ins the file uses are given their one-thread meaning here; nothing of this is visible to the product build.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#define __launch_bounds__(...)
struct EmuIdx {
    unsigned x, y, z;
};
static EmuIdx threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};

template <class T>
static inline T __shfl_xor_sync(unsigned, T, int) {  // a warp of one lane: the partner contributes nothing
    return T(0);
}
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline void __syncthreads() {}
static inline int atomicMin(int* a, int v) {
    const int old = *a;
    if (v < old) *a = v;
    return old;
}
// fp32 round-to-nearest single operations (compile with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

#define GDRN_PNP_THREADS 1
#define GDRN_PNP_HYP_THREADS 1
#define GDRN_PNP_GATHER_THREADS 1
#define GDRN_LAUNCH(kernel, grid, block, stream, ...)              \
    do {                                                           \
        const dim3 g_ = (grid);                                    \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                  \
            for (unsigned by_ = 0; by_ < g_.y; ++by_)              \
                for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {        \
                    blockIdx.x = bx_;                              \
                    blockIdx.y = by_;                              \
                    blockIdx.z = bz_;                              \
                    kernel(__VA_ARGS__);                           \
                }                                                  \
    } while (0)
#define cudaGetLastError() cudaSuccess
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)

#include "../../gdr_net_b200/csrc/gdrn_internal.h"
namespace gdrn {
static char g_msg[512];
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_msg, sizeof(g_msg), fmt, ap);
    va_end(ap);
    fprintf(stderr, "emu: %s\n", g_msg);
    return code;
}
int cuda_error(cudaError_t, const char*, int) { return GDRN_ERR_CUDA; }
void count_launch() {}
}  // namespace gdrn

#include "../../gdr_net_b200/csrc/pnp_ransac.cu"

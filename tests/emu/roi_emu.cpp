// TEST INFRASTRUCTURE ONLY (built by tests/test_roi_targets_cpu.py into a temporary directory, never into libgdrn_b200.so).
//
// Compiles gdr_net_b200/csrc/roi.cu for the HOST with ONE thread per CTA, so the CPU suite can check the restated
// cv2.getAffineTransform / cv2.warpAffine fixed-point arithmetic of the crop / target kernels against cv2 itself.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define __launch_bounds__(...)
struct EmuIdx {
    unsigned x, y, z;
};
static EmuIdx threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline void __syncthreads() {}
// explicitly rounded single operations (compile with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __double2int_rn(double a) { return (int)nearbyint(a); }  // round half to even, like cvt.rni

#define GDRN_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)   \
    do {                                                           \
        const dim3 g_ = (grid);                                    \
        gridDim.x = g_.x;                                          \
        gridDim.y = g_.y;                                          \
        gridDim.z = g_.z;                                          \
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

#include "../../gdr_net_b200/csrc/gdrn_internal.h"
namespace gdrn {
float s_fps[3 * 1024];  // the kernel's dynamic shared memory (`extern __shared__ float s_fps[]`)
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    return code;
}
int cuda_error(cudaError_t, const char*, int) { return GDRN_ERR_CUDA; }
void count_launch() {}
}  // namespace gdrn

#include "../../gdr_net_b200/csrc/roi.cu"

// Internal helpers shared by the .cu translation units of libgdrn_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define GDRN_OK 0
#define GDRN_ERR_ARG (-1)
#define GDRN_ERR_CUDA (-2)

namespace gdrn {

// thread-local last-error message (never throw across the C ABI)
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* file, int line);
int num_sms();
void count_launch();
bool pdl_enabled();     // programmatic dependent launch for the GEMM kernels (GDRN_PDL >= 1 / gdrn_set_pdl)
bool pdl_ew_enabled(bool forward_kernel);  // ... and for the elementwise / pack kernels (GDRN_PDL = 2: all, 3: forward-pass kernels only)
void set_pdl(int on);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// bf16 elements, 128-byte swizzle, zero fill for out-of-bounds box elements.
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box);

}  // namespace gdrn

#ifdef __CUDACC__
namespace gdrn {
// <<<>>> with the programmatic-dependent-launch attribute (when enabled).  ONLY for kernels whose first statements are
// pdl_launch_dependents(); pdl_wait();  (ptx.cuh) -- a kernel launched this way may start before its predecessor has finished.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool forward_kernel, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl_ew_enabled(forward_kernel)) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        na = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace gdrn
#define GDRN_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
    (void)gdrn::launch_pdl(false, kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)
#define GDRN_LAUNCH_PDL_FWD(kernel, grid, block, smem, stream, ...) \
    (void)gdrn::launch_pdl(true, kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)
#endif

// Plain launch, spelled as a macro so that tests/emu can rebuild a translation unit for the host (one thread per CTA) and
// single-step its kernels in the CPU test suite; in the library this IS kernel<<<grid, block, smem, stream>>>(args...).
#ifndef GDRN_LAUNCH_SMEM
#define GDRN_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#define GDRN_CUDA_OK(expr)                                                      \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return gdrn::cuda_error(_e, __FILE__, __LINE__); \
    } while (0)

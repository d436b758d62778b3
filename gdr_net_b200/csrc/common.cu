// Error state, device queries, TMA descriptor creation and the launch counter of libgdrn_b200.so.
#include <stdarg.h>

#include <atomic>

#include <stdlib.h>

#include "gdrn_internal.h"
#include "ptx.cuh"

namespace gdrn {

static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char* file, int line) {
    return set_error(GDRN_ERR_CUDA, "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Programmatic dependent launch for the GEMM kernels (GDRN_PDL=0 disables; A/B).  See ptx.cuh pdl_wait().
// mode 0: off; 1: GEMM kernels only; 2: GEMM + elementwise / pack kernels; 3: GEMM + the forward pass's elementwise kernels
static int g_pdl = -1;
static int pdl_mode() {
    if (g_pdl < 0) {
        const char* e = getenv("GDRN_PDL");
        g_pdl = e ? atoi(e) : 1;
    }
    return g_pdl;
}
bool pdl_enabled() { return pdl_mode() >= 1; }
bool pdl_ew_enabled(bool forward_kernel) { return pdl_mode() == 2 || (pdl_mode() == 3 && forward_kernel); }
void set_pdl(int mode) { g_pdl = mode < 0 ? 0 : mode; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (fn == nullptr) return set_error(GDRN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t d[5];
    cuuint64_t s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(out, GDRN_STORE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return set_error(GDRN_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (%d): rank=%d base=%p dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                         (int)r, rank, base, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                         (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
                         rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
    }
    return 0;
}

}  // namespace gdrn

extern "C" const char* gdrn_last_error() { return gdrn::g_err; }
extern "C" long gdrn_launch_count() { return gdrn::g_launches.load(); }
extern "C" int gdrn_set_pdl(int on) {
    gdrn::set_pdl(on);
    return 0;
}
extern "C" int gdrn_abi_version() { return 1; }

// 0 = bf16 planes, 1 = fp16 planes (compile-time GDRN_STORE_F16); the host side mirrors dtype and lo-plane scale
extern "C" int gdrn_storage_format() { return GDRN_STORE_F16; }
extern "C" float gdrn_lo_scale() { return gdrn::kLoScale; }

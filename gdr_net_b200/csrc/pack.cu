// Layout conversion kernels (HBM-bound, coalesced on the destination side):
//   * fp32 OIHW / IOHW / [N][K] parameters  ->  K-major bf16 (hi, lo) GEMM operands
//   * fp32 split-K weight-gradient workspace ->  fp32 gradients in the reference's parameter layout
//   * NCHW fp32 input crops                 ->  im2col matrix of the 7x7/2 stem (bf16 hi, lo)
// Parameter layouts follow the reference state_dict: Conv2d OIHW, ConvTranspose2d
// IOHW (cdpn_rot_head_region.py:82-91), Linear [out][in] with fc1's input flattened from NCHW
// (conv_pnp_net.py:145).
#include <cuda_fp16.h>

#include "gdrn_internal.h"
#include "ptx.cuh"

namespace gdrn {

// dst[o][tap * ipad + i] = src[o*so + i*si + r'*sr + s'*ss],  zero for i >= I, o >= O or columns >= taps*ipad
__global__ void pack_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst_hi,
                                   __nv_bfloat16* __restrict__ dst_lo, int O, int I, int KH, int KW, int opad, int ipad,
                                   int krow, long so, long si, long sr, long ss, int flip) {
    const long total = (long)opad * krow;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / krow);
        const int col = (int)(idx - (long)o * krow);
        const int tap = col / ipad;
        const int i = col - tap * ipad;
        float v = 0.f;
        if (o < O && i < I && tap < KH * KW) {
            int r = tap / KW, s = tap - (tap / KW) * KW;
            if (flip) {
                r = KH - 1 - r;
                s = KW - 1 - s;
            }
            v = __ldg(src + o * so + i * si + r * sr + s * ss);
        }
        uint16_t h, l;
        split1(v, h, l);
        reinterpret_cast<uint16_t*>(dst_hi)[idx] = h;
        if (dst_lo != nullptr) reinterpret_cast<uint16_t*>(dst_lo)[idx] = l;
    }
}

// All per-step weight re-packs of the network in ONE launch: a device-resident job table (built once by the host,
// pointers are stable) replaces ~100 tiny launches.  Jobs are padded to multiples of 256 elements so that each block
// belongs to exactly one job (one binary search per block).
struct PackJob {
    const float* src;
    __nv_bfloat16* dst_hi;
    __nv_bfloat16* dst_lo;
    long so, si, sr, ss;
    long elem_begin;  // prefix sum of ceil256(opad * ipad): (dst row, dst channel) work items
    int O, I, KH, KW, opad, ipad, krow, flip;
};

__global__ void __launch_bounds__(256) pack_weight_batched_kernel(const PackJob* __restrict__ jobs, int njobs, long total_blocks) {
    // One thread = one (dst row o, dst channel i) pair, looping over the taps: a warp reads 32 short contiguous runs of
    // the fp32 source (the taps of OIHW / IOHW are innermost: >= 50 % sector efficiency) and writes 64 contiguous bytes
    // per tap and plane.  (The first version mapped threads to dst elements: 2-byte gathers with a 36-byte stride.)
    __shared__ PackJob job;
    for (long blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const long e = blk * 256;
            int lo = 0, hi = njobs - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (jobs[mid].elem_begin <= e) lo = mid; else hi = mid - 1;
            }
            job = jobs[lo];
        }
        __syncthreads();
        const long item = blk * 256 + threadIdx.x - job.elem_begin;  // elem_begin counts (o, i) ITEMS here
        if (item >= (long)job.opad * job.ipad) continue;
        const int o = (int)(item / job.ipad);
        const int i = (int)(item - (long)o * job.ipad);
        const int taps = job.KH * job.KW;
        const bool valid = (o < job.O) && (i < job.I);
        const float* sp = job.src + o * job.so + i * job.si;
        __nv_bfloat16* dh = job.dst_hi + (long)o * job.krow + i;
        __nv_bfloat16* dl = job.dst_lo != nullptr ? job.dst_lo + (long)o * job.krow + i : nullptr;
        for (int tap = 0; tap < taps; ++tap) {
            float v = 0.f;
            if (valid) {
                int r = tap / job.KW, s2 = tap - (tap / job.KW) * job.KW;
                if (job.flip) {
                    r = job.KH - 1 - r;
                    s2 = job.KW - 1 - s2;
                }
                v = __ldg(sp + r * job.sr + s2 * job.ss);
            }
            uint16_t h, l;
            split1(v, h, l);
            reinterpret_cast<uint16_t*>(dh)[(long)tap * job.ipad] = h;
            if (dl != nullptr) reinterpret_cast<uint16_t*>(dl)[(long)tap * job.ipad] = l;
        }
    }
}

// grad[o*so + i*si + r'*sr + s'*ss] (+)= sum_ks ws[ks][o*krow + tap*ipad + i]
__global__ void unpack_wgrad_kernel(const float* __restrict__ ws, float* __restrict__ grad, int O, int I, int KH, int KW,
                                    int ipad, int krow_, int ksplit, long ks_stride, long so, long si, long sr, long ss, int flip,
                                    int accumulate) {
    // one thread per (o, tap, i): reads of every split are coalesced over i.  (A thread-per-(o, i) variant with contiguous
    // per-thread tap writes was measured 2.4x SLOWER: too few threads for 9 x ksplit dependent loads each.)
    const int taps = KH * KW;
    const long krow = krow_;
    const long total = (long)O * taps * I;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % I);
        const long t2 = idx / I;
        const int tap = (int)(t2 % taps);
        const int o = (int)(t2 / taps);
        const float* src = ws + (long)o * krow + (long)tap * ipad + i;
        float acc = 0.f;
        for (int ks = 0; ks < ksplit; ++ks) acc += src[ks * ks_stride];
        int r = tap / KW, s = tap - (tap / KW) * KW;
        if (flip) {
            r = KH - 1 - r;
            s = KW - 1 - s;
        }
        float* d = grad + o * so + i * si + r * sr + s * ss;
        *d = accumulate ? (*d + acc) : acc;
    }
}

// Stem im2col: x NCHW fp32 [B,3,H,W] -> A[(b,oh,ow)][192] with k = (r*7+s)*3 + c (147 valid), conv 7x7 s2 p3.
// One block per (image, output row): the 7 x 3 input rows it needs are staged in shared memory with a zero halo
// (coalesced fp32 loads), then every thread emits 16-byte groups of 8 consecutive k for consecutive positions
// (coalesced stores; the matrix is 403 MB per plane at B=64, so the stores are what matters).
template <int W>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a_hi,
                                                          __nv_bfloat16* __restrict__ a_lo, int B, int H) {
    constexpr int WP = W + 6;  // 3-pixel zero halo left and right
    __shared__ float tile[3][7][WP];
    const int Ho = H / 2, Wo = W / 2;
    const int b = blockIdx.x / Ho, oh = blockIdx.x - b * Ho;
    for (int e = threadIdx.x; e < 3 * 7 * WP; e += 256) {
        const int col = e % WP;
        const int r = (e / WP) % 7;
        const int c = e / (WP * 7);
        const int ih = oh * 2 + r - 3, iw = col - 3;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + (((long)b * 3 + c) * H + ih) * W + iw);
        tile[c][r][col] = v;
    }
    __syncthreads();
    const long row0 = ((long)b * Ho + oh) * Wo;
    for (int item = threadIdx.x; item < Wo * 24; item += 256) {
        const int ow = item / 24, g = item - ow * 24;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;
            float val = 0.f;
            if (k < 147) {
                const int tap = k / 3, c = k - tap * 3;
                const int r = tap / 7, s2 = tap - r * 7;
                val = tile[c][r][2 * ow + s2];
            }
            v[j] = val;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        }
        const long idx = (row0 + ow) * 24 + g;
        reinterpret_cast<uint4*>(a_hi)[idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (a_lo != nullptr) reinterpret_cast<uint4*>(a_lo)[idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

static inline int grid_for(long total, int block) {
    long g = (total + block - 1) / block;
    const long cap = (long)num_sms() * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace gdrn

using namespace gdrn;

extern "C" int gdrn_pack_weight(const float* src, void* dst_hi, void* dst_lo, int O, int I, int KH, int KW, int opad,
                                int ipad, int krow, long so, long si, long sr, long ss, int flip, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (krow < KH * KW * ipad || opad < O || ipad < I) return set_error(GDRN_ERR_ARG, "pack_weight: bad padding");
    const long total = (long)opad * krow;
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, (__nv_bfloat16*)dst_hi, (__nv_bfloat16*)dst_lo, O, I,
                                                                KH, KW, opad, ipad, krow, so, si, sr, ss, flip);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int gdrn_unpack_wgrad(const float* ws, float* grad, int O, int I, int KH, int KW, int ipad, int krow, int ksplit,
                                 long ks_stride, long so, long si, long sr, long ss, int flip, int accumulate,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const long total = (long)O * KH * KW * I;
    unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(ws, grad, O, I, KH, KW, ipad, krow, ksplit, ks_stride, so, si,
                                                                 sr, ss, flip, accumulate);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int gdrn_stem_im2col(const float* x, void* a_hi, void* a_lo, int B, int H, int W, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (H % 2 || W != 256) return set_error(GDRN_ERR_ARG, "stem_im2col: expects 256-wide crops (INPUT_RES=256), even height");
    stem_im2col_kernel<256><<<B * (H / 2), 256, 0, stream>>>(x, (__nv_bfloat16*)a_hi, (__nv_bfloat16*)a_lo, B, H);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int gdrn_pack_weight_batched(const void* jobs_dev, int njobs, long total_blocks, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    static_assert(sizeof(PackJob) == 96, "PackJob layout is mirrored by the host (gdr_net_b200/engine.py)");
    if (njobs <= 0) return 0;
    long grid = total_blocks;
    const long cap = (long)num_sms() * 16;
    if (grid > cap) grid = cap;
    pack_weight_batched_kernel<<<(int)grid, 256, 0, stream>>>(reinterpret_cast<const PackJob*>(jobs_dev), njobs, total_blocks);
    GDRN_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

// Shared between gemm_fwd.cu (1-CTA tiles) and gemm_fwd2.cu (2-CTA cta_group::2 tiles).
#pragma once
#include <cuda.h>

namespace gdrn {

struct GemmParams {
    CUtensorMap tmA[2][4];  // [plane hi/lo][stride-2 phase]
    CUtensorMap tmB[2];     // [plane hi/lo]
    int mode;               // 0 = plain 2-D GEMM, 1 = conv (4-D boxes)
    int M, N;               // valid rows / cols
    int num_m_tiles, num_n_tiles, num_kb;
    int cchunks, KW, pad, stride;
    int tiles_per_img, TH, TN;
    void* out_hi;
    void* out_lo;
    float* out_f32;
    int ldc;
    const float* bias;
    int act;       // 0 none, 1 LeakyReLU(0.1)
    float* stats;  // [2][N]: sum, sum of squares (accumulated with atomics) or nullptr
    int cluster;   // 1, or 2: CTA pairs share one n_tile and each TMA-multicasts half of the weight tile to both
};


int launch_gemm_2cta(const GemmParams& p, int nsplit, cudaStream_t stream);  // gemm_fwd2.cu

}  // namespace gdrn

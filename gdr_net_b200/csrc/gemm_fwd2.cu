// 2-CTA (tcgen05 cta_group::2) variant of the implicit-GEMM forward kernel: one 256 x 256 output tile per CTA PAIR.
//
// Why: with 128 x 256 tiles a single CTA must pull 16 KB (A) + 32 KB (B) per 64-deep k-block while the tensor core
// needs only 512 cycles for it = 96 B/cycle/SM, more than an SM's ~64 B/cycle L2->smem ingress: the 1-CTA kernel tops
// out at ~2/3 of the MMA rate (measured 57 % tensor-pipe, profiles/r1_ncu_full_kernel_metrics.txt).  In cta_group::2
// mode the pair issues ONE 256x256x16 MMA whose A rows and B rows are split across the two CTAs' shared memories:
// each CTA loads its own 128 pixels of A (16 KB) and only HALF of the weight tile (128 of 256 output channels, 16 KB)
// = 64 B/cycle/SM.  Each CTA's TMEM receives the accumulator rows of its own 128 pixels, so the epilogue is unchanged.
//
//   * both CTAs: TMA producer (own A box, own half of B) -- complete_tx lands on the LEADER's full barrier
//   * leader CTA only: MMA issuer (tcgen05.mma.cta_group::2), commits are multicast to both CTAs' barriers
//   * both CTAs: epilogue warpgroups; the peer's warps arrive remotely on the leader's tmem-empty barrier
#include <stdlib.h>

#include "gdrn_internal.h"
#include "gemm_params.h"
#include "ptx.cuh"

namespace gdrn {

namespace {

constexpr int kBM = 128;       // rows per CTA (256 per pair)
constexpr int kBN = 256;       // tile width (128 weight rows per CTA)
constexpr int kBK = 64;
constexpr int kEpi = 8;        // epilogue warps per CTA
constexpr int kTr = kEpi * 32 * 17 * 4;
constexpr int kAux = 4096 + kTr;

template <int NSPLIT>
struct Cfg2 {
    static constexpr int NPL = (NSPLIT == 1) ? 1 : 2;
    static constexpr int A_BYTES = kBM * kBK * 2;        // 16 KB
    static constexpr int B_BYTES = (kBN / 2) * kBK * 2;  // 16 KB: this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
    static constexpr int STAGES_RAW = (227 * 1024 - kAux - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kAux + 1024;
    static constexpr int TMEM_COLS = 512;  // two 256-column accumulator buffers
    static_assert(NSPLIT == 1, "2-CTA tiles are used by the single-plane mode only (TMEM: 2 x 256 columns)");
};

__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1) {
    // executed by both CTAs of the pair; clearing the peer bit of the barrier address makes the bytes count on CTA 0's barrier
    const uint32_t bar = smem_u32(leader_bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1, int c2,
                                             int c3) {
    const uint32_t bar = smem_u32(leader_bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {  // arrives on this barrier offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {  // arrive on the same barrier offset in cluster CTA 0
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

template <int NSPLIT>
__global__ void __launch_bounds__(128 + 32 * kEpi, 1) gemm_fwd2_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = Cfg2<NSPLIT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);  // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;                // per CTA: "this stage may be refilled"
    uint64_t* tfull_bar = empty_bar + STAGES;               // per CTA: accumulator buffer complete
    uint64_t* tempty_bar = tfull_bar + 2;                   // leader only: both CTAs drained the buffer (8 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_stats = reinterpret_cast<float*>(aux + 512);
    float* s_tr = reinterpret_cast<float*>(aux + 4096);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int my_pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_groups = (p.num_m_tiles / 2) * p.num_n_tiles;  // (pair of m_tiles, n_tile)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmB[0]);
        tma_prefetch_desc(&p.tmA[0][0]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * kBN; i += blockDim.x) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int grp = my_pair; grp < num_groups; grp += num_pairs) {
                const int n_tile = grp % p.num_n_tiles;
                const int m_tile = (grp / p.num_n_tiles) * 2 + (int)crank;
                int n0 = 0, h0 = 0;
                if (p.mode == 1) {
                    if (p.TN == 1) {
                        n0 = m_tile / p.tiles_per_img;
                        h0 = (m_tile % p.tiles_per_img) * p.TH;
                    } else {
                        n0 = m_tile * p.TN;
                    }
                }
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
                    uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
                    if (p.mode == 1) {
                        const int tap = kb / p.cchunks;
                        const int cc = kb - tap * p.cchunks;
                        const int r = tap / p.KW;
                        const int s = tap - r * p.KW;
                        int dh = r - p.pad, dw = s - p.pad, map = 0;
                        if (p.stride == 2) {
                            map = ((dh & 1) << 1) | (dw & 1);
                            dh >>= 1;
                            dw >>= 1;
                        }
                        tma2_load_4d(st, &p.tmA[0][map], &full_bar[stage], cc * kBK, dw, h0 + dh, n0);
                    } else {
                        tma2_load_2d(st, &p.tmA[0][0], &full_bar[stage], kb * kBK, m_tile * kBM);
                    }
                    tma2_load_2d(st + Cfg::A_BYTES, &p.tmB[0], &full_bar[stage], kb * kBK, n_tile * kBN + (int)crank * (kBN / 2));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(256, kBN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int grp = my_pair; grp < num_groups; grp += num_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t b_s = a_s + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t da = make_smem_desc(a_s + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc(b_s + k * 32, 16, 1024);
                        umma2(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma2_commit(&empty_bar[stage]);  // stage free in both CTAs once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma2_commit(&tfull_bar[acc]);  // accumulator complete -> both CTAs' epilogues
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int grp2 = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        float* tr = s_tr + (warp - 4) * (32 * 17);
        __nv_bfloat16* out_hi = reinterpret_cast<__nv_bfloat16*>(p.out_hi);
        for (int tg = my_pair + grp2 * num_pairs, it = grp2; tg < num_groups; tg += 2 * num_pairs, it += 2) {
            const int n_tile = tg % p.num_n_tiles;
            const int m_tile = (tg / p.num_n_tiles) * 2 + (int)crank;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const long grow = (long)m_tile * kBM + row;
            const bool row_ok = grow < p.M;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                const int col0 = n_tile * kBN + c * 32;
                if (col0 >= p.N) continue;
                float f[32];
                {
                    uint32_t raw[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + c * 32, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(raw[j]);
                }
                const bool full_chunk = (col0 + 32 <= p.N);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (full_chunk || col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
                }
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : 0.1f * f[j];
                }
                if (row_ok) {
                    if (p.out_f32 != nullptr) {
                        float* dst = p.out_f32 + grow * p.ldc + col0;
                        if (full_chunk) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.ldc) dst[j] = (col0 + j < p.N) ? f[j] : 0.f;
                        }
                    }
                    if (out_hi != nullptr) {
                        uint32_t hi[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float a = f[2 * j], b = f[2 * j + 1];
                            if (!full_chunk) {
                                if (col0 + 2 * j >= p.N) a = 0.f;
                                if (col0 + 2 * j + 1 >= p.N) b = 0.f;
                            }
                            hi[j] = pack_hi2(a, b);
                        }
                        const int ncopy = full_chunk ? 4 : ((min(p.ldc, col0 + 32) - col0) / 8);
                        uint4* dh = reinterpret_cast<uint4*>(out_hi + grow * p.ldc + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < ncopy) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    }
                }
                if (p.stats != nullptr) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) tr[lane * 17 + j] = row_ok ? f[h * 16 + j] : 0.f;
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
                        const int col = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float v = tr[(r0 + i) * 17 + col];
                            s1 += v;
                            s2 = fmaf(v, v, s2);
                        }
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                        __syncwarp();
                        if (lane < 16) {
                            atomicAdd(&s_stats[c * 32 + h * 16 + lane], s1);
                            atomicAdd(&s_stats[kBN + c * 32 + h * 16 + lane], s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cta0(&tempty_bar[acc]);
            }
        }
        if (p.stats != nullptr) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int n_tile = my_pair % p.num_n_tiles;  // fixed per pair (num_pairs % num_n_tiles == 0)
            for (int i = threadIdx.x - 128; i < kBN; i += 256) {
                const int col = n_tile * kBN + i;
                if (col < p.N) {
                    atomicAdd(p.stats + col, s_stats[i]);
                    atomicAdd(p.stats + p.N + col, s_stats[kBN + i]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may not exit (or free TMEM) while the leader's MMAs still read its smem / write its TMEM
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

int launch_gemm_2cta(const GemmParams& p, int nsplit, cudaStream_t stream) {
    if (nsplit != 1) return set_error(GDRN_ERR_ARG, "2-CTA tiles: single-plane mode only");
    using Cfg = Cfg2<1>;
    auto kern = gemm_fwd2_kernel<1>;
    static bool attr_set = false;
    if (!attr_set) {
        GDRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int groups = (p.num_m_tiles / 2) * p.num_n_tiles;
    int pairs = groups < num_sms() / 2 ? groups : num_sms() / 2;
    pairs -= pairs % p.num_n_tiles;
    if (pairs <= 0) pairs = p.num_n_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(pairs * 2);
    cfg.blockDim = dim3(128 + 32 * kEpi);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GDRN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return 0;
}

}  // namespace gdrn

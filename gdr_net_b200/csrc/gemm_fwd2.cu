// 2-CTA (tcgen05 cta_group::2) variant of the implicit-GEMM forward kernel: one 256-row x BLOCK_N output tile per CTA PAIR.
//
// Why: the 1-CTA kernel is SHARED-MEMORY-BANDWIDTH bound, not tensor-bound.  Per 64-deep k-block an SM's shared memory
// serves the TMA writes of the stage AND the operand reads of every MMA (A 4 KB + B BLOCK_N*32 B per 16-wide k-step):
//   1 pass, 128x256 tile : 48 KB written + 4 x 12 KB read  =  96 KB per 512 MMA cycles = 187 B/clk  (limit 128 B/clk)
//   3 pass, 128x128 tile : 64 KB written + 4 x 24 KB read  = 160 KB per 768 MMA cycles = 208 B/clk
// which is exactly what round 1 measured (tensor pipe 70 % / the 3-pass kernel at 0.66 of peak).  In cta_group::2 mode the
// pair issues ONE 256 x BLOCK_N x 16 MMA whose A rows AND B rows are split across the two CTAs' shared memories: each CTA
// stages (and its tensor core reads) its own 128 pixels of A but only HALF of the weight tile:
//   1 pass, pair 256x256 : 32 KB written + 4 x  8 KB read  =  64 KB per 512 MMA cycles = 125 B/clk
//   3 pass, pair 256x128 : 48 KB written + 4 x 18 KB read  = 120 KB per 768 MMA cycles = 156 B/clk
// (Round 1's first 2-CTA experiment showed no gain because it still had the slow single-lane producer / issuer loops; this
// version has the warp-uniform lean loops of gemm_fwd.cu.)
//
//   * both CTAs : TMA producer warp (own A box, own half of B; complete_tx lands on the LEADER's full barrier)
//   * leader CTA: MMA issuer warp (tcgen05.mma.cta_group::2); commits are multicast to both CTAs' barriers
//   * both CTAs : two epilogue warpgroups (shared code: gemm_epilogue.cuh); peers arrive remotely on the leader's
//                 tmem-empty barrier.  Each CTA's TMEM receives the accumulator rows of its own 128 pixels.
#include <stdlib.h>

#include "gdrn_internal.h"
#include "gemm_epilogue.cuh"
#include "gemm_params.h"
#include "ptx.cuh"

namespace gdrn {

namespace {

constexpr int kBM = 128;  // rows per CTA (256 per pair)
constexpr int kBK = 64;
constexpr int kEpi = 8;  // epilogue warps per CTA
constexpr int kTr = kEpi * 32 * 17 * 4;
constexpr int kAux = 4096 + kTr;

template <int BLOCK_N, int NSPLIT>
struct Cfg2 {
    static constexpr int NPL = (NSPLIT == 1) ? 1 : 2;
    static constexpr int A_BYTES = kBM * kBK * 2;            // 16 KB per plane
    static constexpr int B_BYTES = (BLOCK_N / 2) * kBK * 2;  // this CTA's half of the weight tile, per plane
    static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
    static constexpr int STAGES_RAW = (227 * 1024 - kAux - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kAux + 1024;
    static constexpr int NACC = (NSPLIT == 3) ? 2 : 1;  // main (+ cross-term) accumulator
    static constexpr int TMEM_COLS = 2 * NACC * BLOCK_N;  // two accumulator sets (the allocation is always the full 512 columns)
    static_assert(TMEM_COLS <= 512, "TMEM overflow");
    static_assert(STAGES >= 3, "pipeline too shallow");
};

template <int BLOCK_N, int NSPLIT>
__global__ void __launch_bounds__(128 + 32 * kEpi, 1) gemm_fwd2_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = Cfg2<BLOCK_N, NSPLIT>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);  // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;                // per CTA: "this stage may be refilled"
    uint64_t* tfull_bar = empty_bar + STAGES;               // per CTA: accumulator buffer complete
    uint64_t* tempty_bar = tfull_bar + 2;                   // leader only: both CTAs drained the buffer (2 x 4 warps)
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
        for (int pl = 0; pl < NPL; ++pl) {
            tma_prefetch_desc(&p.tmB[pl]);
            tma_prefetch_desc(&p.tmA[pl][0]);
        }
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
    for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue under our tail
    pdl_wait();               // nothing above touched global memory; from here on we read what the predecessor wrote

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs, whole warp, one lane issues)
        const uint32_t smem_base = smem_u32(smem), empty0 = smem_u32(empty_bar);
        const uint32_t full0_local = smem_u32(full_bar);
        const uint32_t full0 = mapa_cta0(full0_local);  // the LEADER's full barriers, as shared::cluster addresses
        const uint64_t tmB0 = reinterpret_cast<uint64_t>(&p.tmB[0]), tmB1 = reinterpret_cast<uint64_t>(&p.tmB[1]);
        const int cch = p.cchunks, mode = p.mode, KWv = p.KW, padv = p.pad, stride2 = (p.stride == 2);
        const int ntap = mode == 1 ? p.num_kb / (cch > 0 ? cch : 1) : 0;
        const int brow = (int)crank * (BLOCK_N / 2);  // this CTA's half of the weight tile
        int stage = 0;
        uint32_t phase = 0;
        for (int grp = my_pair; grp < num_groups; grp += num_pairs) {
            const int n_tile = grp % p.num_n_tiles;
            const int m_tile = (grp / p.num_n_tiles) * 2 + (int)crank;
            const int bn0 = n_tile * BLOCK_N + brow;
            if (mode == 1) {
                int n0 = 0, h0 = 0;
                if (p.TN == 1) {
                    n0 = m_tile / p.tiles_per_img;
                    h0 = (m_tile - n0 * p.tiles_per_img) * p.TH;
                } else {
                    n0 = m_tile * p.TN;
                }
                int r = 0, s2 = 0;
                for (int t = 0; t < ntap; ++t) {
                    int dh = r - padv, dw = s2 - padv, map = 0;
                    if (stride2) {
                        map = ((dh & 1) << 1) | (dw & 1);
                        dh >>= 1;  // arithmetic shift == floor division
                        dw >>= 1;
                    }
                    if (++s2 == KWv) {
                        s2 = 0;
                        ++r;
                    }
                    const uint64_t tmA0 = reinterpret_cast<uint64_t>(&p.tmA[0][map]);
                    const uint64_t tmA1 = reinterpret_cast<uint64_t>(&p.tmA[1][map]);
                    const int hh = h0 + dh;
                    int kw = t * cch * kBK;
                    for (int cc = 0; cc < cch; ++cc, kw += kBK) {
                        const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                        if (elect_one()) {
                            if (leader) mbar_arrive_expect_tx_u32(full0_local + stage * 8, 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
                            tma2_load_4d_u32(dst, tmA0, fb, cc * kBK, dw, hh, n0);
                            if (NPL == 2) tma2_load_4d_u32(dst + Cfg::A_BYTES, tmA1, fb, cc * kBK, dw, hh, n0);
                            tma2_load_2d_u32(dst + NPL * Cfg::A_BYTES, tmB0, fb, kw, bn0);
                            if (NPL == 2) tma2_load_2d_u32(dst + NPL * Cfg::A_BYTES + Cfg::B_BYTES, tmB1, fb, kw, bn0);
                        }
                        __syncwarp();
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            } else {
                const uint64_t tmA0 = reinterpret_cast<uint64_t>(&p.tmA[0][0]);
                const uint64_t tmA1 = reinterpret_cast<uint64_t>(&p.tmA[1][0]);
                const int row0 = m_tile * kBM;
                for (int kb = 0, kc = 0; kb < p.num_kb; ++kb, kc += kBK) {
                    const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                    mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx_u32(full0_local + stage * 8, 2 * Cfg::STAGE_BYTES);
                        tma2_load_2d_u32(dst, tmA0, fb, kc, row0);
                        if (NPL == 2) tma2_load_2d_u32(dst + Cfg::A_BYTES, tmA1, fb, kc, row0);
                        tma2_load_2d_u32(dst + NPL * Cfg::A_BYTES, tmB0, fb, kc, bn0);
                        if (NPL == 2) tma2_load_2d_u32(dst + NPL * Cfg::A_BYTES + Cfg::B_BYTES, tmB1, fb, kc, bn0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, whole warp, one lane issues)
        if (leader) {
            constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 0, 0);
            const uint64_t desc_const = make_smem_desc(0, 16, 1024);
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
            const int nkb = p.num_kb;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int grp = my_pair; grp < num_groups; grp += num_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait_u32(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_main = tmem_base + acc * Cfg::NACC * BLOCK_N;
                const uint32_t d_cross = d_main + BLOCK_N;
                uint32_t accum = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t da = desc_const | (uint64_t)(a_addr >> 4);
                    const uint64_t db = da + ((NPL * Cfg::A_BYTES) >> 4);
                    mbar_wait_u32(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            const uint32_t ac = (k == 0) ? accum : 1u;
                            if (NSPLIT == 3) {
                                umma2(d_main, da + 2 * k, db + 2 * k, idesc, ac);                                // hi x hi
                                umma2(d_cross, da + 2 * k, db + (Cfg::B_BYTES >> 4) + 2 * k, idesc, ac);          // hi x lo
                                umma2(d_cross, da + (Cfg::A_BYTES >> 4) + 2 * k, db + 2 * k, idesc, 1u);          // lo x hi
                            } else {
                                umma2(d_main, da + 2 * k, db + 2 * k, idesc, ac);
                            }
                        }
                        umma2_commit_u32(empty0 + stage * 8);  // stage free in BOTH CTAs once these MMAs retire
                    }
                    __syncwarp();
                    accum = 1u;
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma2_commit_u32(tfull0 + acc * 8);  // accumulator complete -> both CTAs' epilogues
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int grp2 = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        float* tr = s_tr + (warp - 4) * (32 * 17);
        for (int tg = my_pair + grp2 * num_pairs, it = grp2; tg < num_groups; tg += 2 * num_pairs, it += 2) {
            const int n_tile = tg % p.num_n_tiles;
            const int m_tile = (tg / p.num_n_tiles) * 2 + (int)crank;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            gemm_epilogue_tile<BLOCK_N, NSPLIT, 1>(p, tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::NACC * BLOCK_N, lane, row, m_tile,
                                                   n_tile, 0, 0, BLOCK_N / 32, tr, s_stats);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cta0(&tempty_bar[acc]);
            }
        }
        if (p.stats != nullptr) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int n_tile = my_pair % p.num_n_tiles;  // fixed per pair (num_pairs % num_n_tiles == 0)
            for (int i = threadIdx.x - 128; i < BLOCK_N; i += 256) {
                const int col = n_tile * BLOCK_N + i;
                if (col < p.N) {
                    atomicAdd(p.stats + col, s_stats[i]);
                    atomicAdd(p.stats + p.N + col, s_stats[BLOCK_N + i]);
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

template <int BLOCK_N, int NSPLIT>
int launch2(const GemmParams& p, cudaStream_t stream) {
    using Cfg = Cfg2<BLOCK_N, NSPLIT>;
    auto kern = gemm_fwd2_kernel<BLOCK_N, NSPLIT>;
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
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int na = 1;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    GDRN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return 0;
}

}  // namespace

// block_n is the PAIR tile width: 256 / 128 (1 pass) or 128 (3 pass); p.num_n_tiles = Cout_pad / block_n, p.num_m_tiles even,
// weight tensor maps with boxes of block_n / 2 rows (each CTA loads its half), p.nphase == 0.
int launch_gemm_2cta(const GemmParams& p, int block_n, int nsplit, cudaStream_t stream) {
    if (p.nphase != 0 || (p.num_m_tiles & 1)) return set_error(GDRN_ERR_ARG, "2-CTA tiles: even tile count, no phase decomposition");
    if (nsplit == 1 && block_n == 256) return launch2<256, 1>(p, stream);
    if (nsplit == 1 && block_n == 128) return launch2<128, 1>(p, stream);  // 128-channel layers: 48 instead of 64 KB of smem traffic per k-block
    if (nsplit == 3 && block_n == 128) return launch2<128, 3>(p, stream);
    return set_error(GDRN_ERR_ARG, "2-CTA tiles: unsupported block_n=%d nsplit=%d", block_n, nsplit);
}

}  // namespace gdrn

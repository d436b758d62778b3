// tcgen05 / TMEM / TMA implicit-GEMM forward kernel (sm_100a).
//
// One persistent, warp-specialised kernel serves every GEMM-shaped forward op of the GDR-Net hot
// path (reference call sites: resnet_backbone.py:69-76 convs, cdpn_rot_head_region.py:82-135,
// conv_pnp_net.py:76-80 + fc1/fc2/fc_r/fc_t :111-157) and, with flipped weights, every dgrad:
//
//   D[M = pixels][N = Cout] = sum_k A[M][K] * W[N][K]          (16-bit operands, fp32 accumulate in TMEM)
//
//   * A is never materialised: for a conv, each 64-wide k-block is one (tap, channel-chunk) and its
//     128x64 A tile is ONE 4-D TMA box {64 ch, Wo, TH, TN} of the NHWC activation tensor at the
//     tap-shifted coordinate; out-of-image rows/cols are zero-filled by TMA (= conv padding).
//     Stride-2 convs read from four phase sub-lattices (one tensor map per (row, col) parity);
//     stride-2 DGRADS run as four output-parity phases over the un-dilated dY (gdrn_conv_dgrad_s2).
//   * W is a K-major [Cout][KH*KW*Cin] 16-bit matrix (prepared by pack kernels), 2-D TMA boxes.
//   * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes.
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = two epilogue
//     warpgroups (tcgen05.ld -> bias/activation -> 16-bit hi[/lo] or fp32 stores, per-channel sum /
//     sum-of-squares for BatchNorm batch statistics).  Accumulators are double-buffered in TMEM and
//     each warpgroup drains one buffer, so two tile epilogues overlap the MMAs of the next tiles.
//     Producer and issuer loops run warp-uniformly (operands in uniform registers), one elected lane
//     issues the asynchronous instructions: ~33 instructions per k-block.
//   * NSPLIT = 3 is the fp32-faithful mode: operands are (hi, lo) 16-bit planes and each k-step issues
//     hi*hi + hi*lo + lo*hi (the lo*lo term, 2^-22 relative with fp16 planes, is dropped).
#include <stdlib.h>

#include <atomic>

#include "gdrn_internal.h"
#include "gemm_epilogue.cuh"
#include "gemm_params.h"
#include "ptx.cuh"

namespace gdrn {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kEpiWarps = 8;                    // two epilogue warpgroups, one per TMEM accumulator buffer
constexpr int kTrBytes = kEpiWarps * 32 * 17 * 4;  // per-warp 32x16 (+1 pad) fp32 transpose buffers for the BN statistics
constexpr int kAuxBytes = 4096 + kTrBytes;

template <int BLOCK_N, int NSPLIT>
struct GemmCfg {
    static constexpr int NPL = (NSPLIT == 1) ? 1 : 2;
    static constexpr int A_BYTES = kBlockM * kBlockK * 2;
    static constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
    static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
    static constexpr int STAGES_RAW = (kSmemBudget - kAuxBytes - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kAuxBytes + 1024;
    // tcgen05.mma accumulates into TMEM with round-toward-zero (measured: tools/exp_accum.py), so the error of a
    // K-long chain grows ~ (K/16) * 2^-24.  The fp32-faithful mode therefore keeps the small cross terms
    // (hi*lo + lo*hi) in their own accumulator and spreads the hi*hi k-blocks round-robin over three more; the
    // epilogue adds the four partials with round-to-nearest fp32 adds.
    // (Splitting the hi*hi chain over several accumulators was measured to change nothing at these K: the 16-bit
    // operand planes dominate the error, so one main + one cross accumulator is used and BLOCK_N can be 128.)
    static constexpr int NMAIN = 1;
    static constexpr int NACC = (NSPLIT == 3) ? 2 : 1;
    static constexpr int TMEM_COLS = 2 * NACC * BLOCK_N;  // double-buffered accumulator set
    static_assert(STAGES >= 2, "pipeline too shallow");
    static_assert(TMEM_COLS <= 512, "TMEM overflow");
};

// Sum the 32 lanes' values column-wise: on return lane j holds sum over lanes of v[j] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float send = upper ? v[i] : v[i + off];
            float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int BLOCK_N, int NSPLIT>
__global__ void __launch_bounds__(128 + 32 * kEpiWarps, 1) gemm_fwd_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, NSPLIT>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;

    // 1024-byte alignment by POINTER arithmetic on the __shared__ array (not through uintptr_t): the compiler keeps the
    // shared address space, so the epilogue's transposes / statistics compile to STS / LDS / ATOMS instead of generic
    // ST / LD / ATOM (+ an address-space check each) -- the BN-statistics epilogue was the bottleneck of the 64-channel layers
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_stats = reinterpret_cast<float*>(aux + 512);  // [2][BLOCK_N]
    float* s_tr = reinterpret_cast<float*>(aux + 4096);    // [kEpiWarps][32][17]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

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
            mbar_init(&tempty_bar[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue under our tail
    pdl_wait();               // nothing above touched global memory; from here on we read what the predecessor wrote
    // persistent CTAs stride over the tile list (phase-major for the stride-2 dgrad)
    const int my_cluster = blockIdx.x, num_clusters = gridDim.x;
    const int num_groups = (p.nphase ? p.nphase : 1) * num_tiles;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        {
            // Executed by the whole warp with warp-uniform control flow; one elected lane issues.  A single thread used to
            // spend ~1100 cycles per k-block here on index arithmetic (two integer divisions, generic->shared conversions, parameter reloads) -- more than the 128..512 cycles of MMA
            // work a k-block carries, so EVERY layer ran at the producer's pace (ncu: this warp ~100 % busy, tensor pipe
            // 11 % / 50 %).  Taps and channel chunks are now nested loops with incremental state; per k-block only the
            // barrier wait, the expect_tx and the TMA issues remain.
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            const uint64_t tmB0 = reinterpret_cast<uint64_t>(&p.tmB[0]), tmB1 = reinterpret_cast<uint64_t>(&p.tmB[1]);
            const int cch = p.cchunks, mode = p.mode, KWv = p.KW, padv = p.pad, stride2 = (p.stride == 2), nph = p.nphase;
            const int ntap_all = mode == 1 ? p.num_kb / (cch > 0 ? cch : 1) : 0;
            int stage = 0;
            uint32_t phase = 0;
            for (int grp = my_cluster; grp < num_groups; grp += num_clusters) {
                int rem = grp, tap_base = 0, ntap = ntap_all;
                if (nph) {
                    const int ph = grp / num_tiles;
                    rem = grp - ph * num_tiles;
                    tap_base = p.ph_tap0[ph];
                    ntap = p.ph_tap0[ph + 1] - tap_base;
                }
                const int n_tile = rem % p.num_n_tiles;
                const int m_tile = rem / p.num_n_tiles;
                const int bn0 = n_tile * BLOCK_N;
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
                        int dh, dw, map = 0, wt = t;
                        if (nph) {
                            dh = p.tap_dh[tap_base + t];
                            dw = p.tap_dw[tap_base + t];
                            wt = p.tap_w[tap_base + t];
                        } else {
                            dh = r - padv;
                            dw = s2 - padv;
                            if (stride2) {
                                map = ((dh & 1) << 1) | (dw & 1);
                                dh >>= 1;  // arithmetic shift == floor division
                                dw >>= 1;
                            }
                            if (++s2 == KWv) {
                                s2 = 0;
                                ++r;
                            }
                        }
                        const uint64_t tmA0 = reinterpret_cast<uint64_t>(&p.tmA[0][map]);
                        const uint64_t tmA1 = reinterpret_cast<uint64_t>(&p.tmA[1][map]);
                        const int hh = h0 + dh;
                        int kw = wt * cch * kBlockK;
                        for (int cc = 0; cc < cch; ++cc, kw += kBlockK) {
                            const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                            mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                            if (elect_one()) {
                                mbar_arrive_expect_tx_u32(fb, Cfg::STAGE_BYTES);
                                tma_load_4d_u32(dst, tmA0, fb, cc * kBlockK, dw, hh, n0);
                                if (NPL == 2) tma_load_4d_u32(dst + Cfg::A_BYTES, tmA1, fb, cc * kBlockK, dw, hh, n0);
                                tma_load_2d_u32(dst + NPL * Cfg::A_BYTES, tmB0, fb, kw, bn0);
                                if (NPL == 2) tma_load_2d_u32(dst + NPL * Cfg::A_BYTES + Cfg::B_BYTES, tmB1, fb, kw, bn0);
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
                    const int row0 = m_tile * kBlockM;
                    for (int kb = 0, kc = 0; kb < p.num_kb; ++kb, kc += kBlockK) {
                        const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx_u32(fb, Cfg::STAGE_BYTES);
                            tma_load_2d_u32(dst, tmA0, fb, kc, row0);
                            if (NPL == 2) tma_load_2d_u32(dst + Cfg::A_BYTES, tmA1, fb, kc, row0);
                            tma_load_2d_u32(dst + NPL * Cfg::A_BYTES, tmB0, fb, kc, bn0);
                            if (NPL == 2) tma_load_2d_u32(dst + NPL * Cfg::A_BYTES + Cfg::B_BYTES, tmB1, fb, kc, bn0);
                        }
                        __syncwarp();
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        {
            // Executed by the WHOLE warp with warp-uniform control flow; only the tcgen05 instructions are issued by
            // one elected lane.  (Inside an `if (lane == 0)` region the compiler wraps every UTCHMMA in an ELECT + 5x R2UR
            // "waterfall" loop, ~20 instructions per MMA.)
            // The shared-memory descriptors of a stage differ only in their 14-bit address field, so they are
            // one constant OR-ed with (address >> 4) and advanced by +2 (32 bytes) per 16-wide k-step -- the generic loop
            // rebuilt eight descriptors per k-block (126 instructions on this single thread).
            constexpr uint32_t idesc = make_idesc(kBlockM, BLOCK_N, 0, 0);
            // 3-pass mode: the w_hi and w_lo tiles of a stage are CONTIGUOUS in shared memory (BLOCK_N rows each, same
            // K-major swizzled layout) and the main / cross accumulators are ADJACENT in TMEM, so a_hi x [w_hi; w_lo] is ONE
            // MMA of width 2*BLOCK_N writing main | cross.  The kernel is shared-memory-bandwidth bound (TMA writes + MMA
            // operand reads exceed 128 B/clk): this reads a_hi once instead of twice per k-step (24 -> 20 KB per k-step).
            constexpr uint32_t idesc_wide = make_idesc(kBlockM, (NSPLIT == 3 ? 2 : 1) * BLOCK_N, 0, 0);
            static_assert(NSPLIT != 3 || (Cfg::NMAIN == 1 && 2 * BLOCK_N <= 256), "stacked hi|lo MMA needs adjacent accumulators");
            const uint64_t desc_const = make_smem_desc(0, 16, 1024);
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
            const int cch = p.cchunks, nph = p.nphase, nkb_all = p.num_kb;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int grp = my_cluster; grp < num_groups; grp += num_clusters, ++it) {
                const int acc = it & 1;
                mbar_wait_u32(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_main = tmem_base + acc * Cfg::NACC * BLOCK_N;
                const uint32_t d_cross = d_main + Cfg::NMAIN * BLOCK_N;
                int nkb = nkb_all;
                if (nph) {
                    const int ph = grp / num_tiles;
                    nkb = (p.ph_tap0[ph + 1] - p.ph_tap0[ph]) * cch;
                }
                uint32_t accum = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t da = desc_const | (uint64_t)(a_addr >> 4);
                    const uint64_t db = da + ((NPL * Cfg::A_BYTES) >> 4);
                    mbar_wait_u32(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint32_t ac = (k == 0) ? accum : 1u;
                            if (NSPLIT == 3) {
                                umma_bf16(d_main, da + 2 * k, db + 2 * k, idesc_wide, ac);                    // a_hi x [w_hi; w_lo] -> main | cross
                                umma_bf16(d_cross, da + (Cfg::A_BYTES >> 4) + 2 * k, db + 2 * k, idesc, 1u);  // a_lo x w_hi -> cross
                            } else {
                                umma_bf16(d_main, da + 2 * k, db + 2 * k, idesc, ac);
                            }
                        }
                        umma_commit_u32(empty0 + stage * 8);
                    }
                    __syncwarp();
                    accum = 1u;
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit_u32(tfull0 + acc * 8);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue: 2 warpgroups x 128 threads.
        // Group g drains TMEM accumulator buffer g, i.e. every other tile of this CTA: two tile epilogues are in flight
        // (needed for small-K layers such as 64->64 3x3, whose 9 k-blocks of MMA are shorter than one epilogue).
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int grp = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        float* tr = s_tr + (warp - 4) * (32 * 17);
        // One tile per CTA (small-M layers: 16x16 / 8x8 maps, FC): there is no second tile to overlap with, so both
        // warpgroups share the single tile's epilogue, each draining half of its column chunks.
        constexpr int NCH = BLOCK_N / 32;
        const bool split_epi = num_groups <= num_clusters && !p.no_split_epi;
        const int c_begin = split_epi ? grp * (NCH / 2) : 0;
        const int c_end = split_epi ? (grp + 1) * (NCH / 2) : NCH;
        for (int tg = split_epi ? my_cluster : my_cluster + grp * num_clusters, it = split_epi ? 0 : grp; tg < num_groups;
             tg += 2 * num_clusters, it += 2) {
            int rem = tg, ph = 0;
            if (p.nphase) {
                ph = tg / num_tiles;
                rem = tg - ph * num_tiles;
            }
            const int n_tile = rem % p.num_n_tiles;
            const int m_tile = rem / p.num_n_tiles;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            gemm_epilogue_tile<BLOCK_N, NSPLIT, Cfg::NMAIN>(p, tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::NACC * BLOCK_N, lane, row,
                                                            m_tile, n_tile, ph, c_begin, c_end, tr, s_stats);
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !(split_epi && grp == 1)) mbar_arrive(&tempty_bar[acc]);
        }
        if (p.stats != nullptr) {
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
            const int n_tile = my_cluster % p.num_n_tiles;   // fixed per CTA (num_clusters % num_n_tiles == 0)
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
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int NSPLIT>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, NSPLIT>;
    auto kern = gemm_fwd_kernel<BLOCK_N, NSPLIT>;
    static bool attr_set = false;
    if (!attr_set) {
        GDRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int groups = (p.nphase ? p.nphase : 1) * p.num_m_tiles * p.num_n_tiles;
    int clusters = groups < num_sms() ? groups : num_sms();
    clusters -= clusters % p.num_n_tiles;  // every CTA keeps one n_tile => per-CTA BatchNorm partial sums
    if (clusters <= 0) clusters = p.num_n_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(clusters);
    cfg.blockDim = dim3(128 + 32 * kEpiWarps);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
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

static thread_local int g_last_variant = 0;

static int dispatch_gemm(const GemmParams& p_in, int block_n, int nsplit, cudaStream_t stream) {
    static int no_split = -1;
    if (no_split < 0) {
        const char* e = getenv("GDRN_NO_SPLIT_EPI");
        no_split = e ? atoi(e) : 0;
    }
    GemmParams p = p_in;
    p.no_split_epi = no_split;
    g_last_variant = block_n * 10 + nsplit;
    if (nsplit == 1) {
        if (block_n == 256) return launch_gemm<256, 1>(p, stream);
        if (block_n == 128) return launch_gemm<128, 1>(p, stream);
        if (block_n == 64) return launch_gemm<64, 1>(p, stream);
    } else if (nsplit == 3) {
        if (block_n == 128) return launch_gemm<128, 3>(p, stream);
        if (block_n == 64) return launch_gemm<64, 3>(p, stream);
    }
    return set_error(GDRN_ERR_ARG, "gemm: unsupported block_n=%d nsplit=%d", block_n, nsplit);
}

// Tile width: the widest BLOCK_N (fewest A re-reads, best smem-bandwidth ratio) that still yields >= ~0.75 waves of
// CTAs on the 148 SMs; small-M layers (layer4, FC) fall back to narrower tiles instead of leaving SMs idle.
static int pick_block_n(int n_pad, int nsplit, int num_m_tiles) {
    const int want = (num_sms() * 3) / 4;
    // fp32x3: (main + cross) x 2 buffers x 128 columns = all 512 TMEM columns, so 128 is the widest tile
    if (nsplit == 1 && n_pad % 256 == 0 && num_m_tiles * (n_pad / 256) >= want) return 256;
    if (n_pad % 128 == 0 && num_m_tiles * (n_pad / 128) >= want) return 128;
    if (n_pad % 64 == 0 && num_m_tiles * (n_pad / 128) < num_sms() / 2) return 64;
    if (n_pad % 128 == 0) return 128;
    if (n_pad % 64 == 0) return 64;
    return -1;
}

// 2-CTA (cta_group::2) pair tiles (gemm_fwd2.cu): 256 x 256 (1 pass) / 256 x 128 (3 pass).  Each CTA stages only half of the
// weight tile, which takes the kernel off its shared-memory-bandwidth limit (see the header of gemm_fwd2.cu).
// GDRN_2CTA=0 disables (A/B); layers with an odd tile count or less than half a wave of pair tiles keep the 1-CTA kernel.
static int g_2cta_mode = -1;
static std::atomic<long> g_2cta_launches{0};
static bool want_2cta(int nsplit, int block_n, int num_m_tiles, int num_n_tiles) {
    if (g_2cta_mode < 0) {
        const char* e = getenv("GDRN_2CTA");
        g_2cta_mode = e ? atoi(e) : 1;
    }
    if (g_2cta_mode != 1 || (num_m_tiles & 1)) return false;
    if (!((nsplit == 1 && (block_n == 256 || block_n == 128)) || (nsplit == 3 && block_n == 128))) return false;
    return (num_m_tiles / 2) * num_n_tiles * 4 >= num_sms();
}

}  // namespace gdrn

using namespace gdrn;

// A/B switch: 1 = cta_group::2 pair tiles (gemm_fwd2.cu) for eligible layers (default), 0 = 1-CTA kernel everywhere
extern "C" int gdrn_set_2cta(int on) {
    g_2cta_mode = on ? 1 : 0;
    return 0;
}
extern "C" long gdrn_2cta_launch_count() { return gdrn::g_2cta_launches.load(); }

extern "C" int gdrn_conv_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, void* y_hi,
                             void* y_lo, float* y_f32, const float* bias, const void* res_hi, const void* res_lo, float* stats,
                             int N, int H, int W, int Cin, int Cout, int Cout_pad, int KH, int KW, int stride, int pad, int ldc,
                             int act, int nsplit, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (nsplit != 1 && nsplit != 3) return set_error(GDRN_ERR_ARG, "conv_fwd: nsplit must be 1 or 3");
    if (nsplit == 3 && (x_lo == nullptr || w_lo == nullptr)) return set_error(GDRN_ERR_ARG, "conv_fwd: lo planes missing");
    if (Cin % 64 != 0) return set_error(GDRN_ERR_ARG, "conv_fwd: Cin=%d must be a multiple of 64", Cin);
    if (stride != 1 && stride != 2) return set_error(GDRN_ERR_ARG, "conv_fwd: stride must be 1 or 2");
    if (H % stride || W % stride) return set_error(GDRN_ERR_ARG, "conv_fwd: H, W must be divisible by stride");
    if (ldc % 8 != 0 || ldc < Cout) return set_error(GDRN_ERR_ARG, "conv_fwd: bad ldc=%d", ldc);
    if (res_hi != nullptr && (Cout % 32 != 0 || y_hi == nullptr))
        return set_error(GDRN_ERR_ARG, "conv_fwd: a residual needs Cout %% 32 == 0 and a 16-bit output");
    if (act < 0 || act > 2) return set_error(GDRN_ERR_ARG, "conv_fwd: act must be 0 (none), 1 (LeakyReLU 0.1) or 2 (ReLU)");
    const int Ho = H / stride, Wo = W / stride;
    if (Wo > 128 || 128 % Wo != 0) return set_error(GDRN_ERR_ARG, "conv_fwd: unsupported output width %d", Wo);
    int TH = 128 / Wo;
    if (TH > Ho) TH = Ho;
    const int TN = 128 / (Wo * TH);
    if (Ho % TH != 0) return set_error(GDRN_ERR_ARG, "conv_fwd: Ho=%d not divisible by tile rows %d", Ho, TH);
    const int block_n = pick_block_n(Cout_pad, nsplit, (N * Ho * Wo + 127) / 128);
    if (block_n < 0) return set_error(GDRN_ERR_ARG, "conv_fwd: Cout_pad=%d must be a multiple of 64", Cout_pad);

    GemmParams p;
    memset(&p, 0, sizeof(p));
    const bool two_cta = want_2cta(nsplit, block_n, (N * Ho * Wo + 127) / 128, Cout_pad / block_n);
    const int cluster = two_cta ? 2 : 1;  // 2-CTA pairs: each CTA loads half of the weight tile
    const int npl = nsplit == 1 ? 1 : 2;
    const void* xs[2] = {x_hi, x_lo};
    const void* ws[2] = {w_hi, w_lo};
    const int K = KH * KW * Cin;
    for (int pl = 0; pl < npl; ++pl) {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xs[pl]);
        for (int ph = 0; ph < (stride == 2 ? 4 : 1); ++ph) {
            const int phh = ph >> 1, phw = ph & 1;
            uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
            uint64_t strides[3] = {(uint64_t)stride * Cin * 2, (uint64_t)stride * W * Cin * 2, (uint64_t)H * W * Cin * 2};
            uint32_t box[4] = {64, (uint32_t)Wo, (uint32_t)TH, (uint32_t)TN};
            const void* base = xb + ((size_t)phh * W + phw) * Cin;
            if (make_tmap(&p.tmA[pl][ph], base, 4, dims, strides, box)) return GDRN_ERR_CUDA;
        }
        uint64_t wdims[2] = {(uint64_t)K, (uint64_t)Cout_pad};
        uint64_t wstr[1] = {(uint64_t)K * 2};
        uint32_t wbox[2] = {64, (uint32_t)(block_n / cluster)};
        if (make_tmap(&p.tmB[pl], ws[pl], 2, wdims, wstr, wbox)) return GDRN_ERR_CUDA;
    }
    p.cluster = cluster;
    p.mode = 1;
    p.M = N * Ho * Wo;
    p.N = Cout;
    p.num_m_tiles = (p.M + 127) / 128;
    p.num_n_tiles = Cout_pad / block_n;
    p.cchunks = Cin / 64;
    p.num_kb = KH * KW * p.cchunks;
    p.KW = KW;
    p.pad = pad;
    p.stride = stride;
    p.TH = TH;
    p.TN = TN;
    p.tiles_per_img = (TN == 1) ? Ho / TH : 0;
    p.out_hi = y_hi;
    p.out_lo = y_lo;
    p.out_f32 = y_f32;
    p.ldc = ldc;
    p.bias = bias;
    p.res_hi = res_hi;
    p.res_lo = res_lo;
    p.act = act;
    p.stats = stats;
    if (two_cta) {
        g_2cta_launches.fetch_add(1, std::memory_order_relaxed);
        g_last_variant = 10000 + block_n * 10 + nsplit;  // +10000: the cta_group::2 kernel
        return launch_gemm_2cta(p, block_n, nsplit, stream);
    }
    return dispatch_gemm(p, block_n, nsplit, stream);
}

// Data gradient of a stride-2 convolution (k = 3, pad 1 or k = 1, pad 0), decomposed by OUTPUT PARITY: the pixels
// (2i + a, 2j + b) of dX only receive the filter taps r = a + pad (mod 2), so each of the four phases is a small stride-1
// conv over the UN-dilated dY with 1 / 2 / 2 / 4 taps -- 9 taps per 4 output pixels instead of the 36 of the
// "conv over zero-inserted dY" formulation, and no zero-inserted tensor in HBM.  One launch covers all phases.
//   du [N][Ho][Wo][Cy] (Cy % 64 == 0), w = the dgrad-packed weights [Cx_pad][K*K*Cy] with flipped taps (pack_conv_dgrad),
//   dx [N][2Ho][2Wo][ldc].  k = 1: only phase (0, 0) is computed, the caller passes a zero-filled dx.
extern "C" int gdrn_conv_dgrad_s2(const void* du_hi, const void* du_lo, const void* w_hi, const void* w_lo, void* dx_hi,
                                  void* dx_lo, const float* bias, float* stats, int N, int Ho, int Wo, int Cy, int Cx, int Cx_pad,
                                  int K, int pad, int ldc, int act, int nsplit, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (nsplit != 1 && nsplit != 3) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: nsplit must be 1 or 3");
    if (nsplit == 3 && (du_lo == nullptr || w_lo == nullptr)) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: lo planes missing");
    if (!((K == 3 && pad == 1) || (K == 1 && pad == 0))) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: k3 p1 or k1 p0 only");
    if (Cy % 64 != 0) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: Cy=%d must be a multiple of 64", Cy);
    if (ldc % 8 != 0 || ldc < Cx) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: bad ldc=%d", ldc);
    if (Wo > 128 || 128 % Wo != 0 || (Ho & (Ho - 1)) != 0) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: unsupported dY size %dx%d", Ho, Wo);
    int TH = 128 / Wo;
    if (TH > Ho) TH = Ho;
    const int TN = 128 / (Wo * TH);
    if (Ho % TH != 0) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: Ho=%d not divisible by tile rows %d", Ho, TH);
    const int num_m_tiles = (N * Ho * Wo + 127) / 128;
    const int nphase = K == 3 ? 4 : 1;
    const int block_n = pick_block_n(Cx_pad, nsplit, num_m_tiles * nphase);
    if (block_n < 0) return set_error(GDRN_ERR_ARG, "conv_dgrad_s2: Cx_pad=%d must be a multiple of 64", Cx_pad);

    GemmParams p;
    memset(&p, 0, sizeof(p));
    const int npl = nsplit == 1 ? 1 : 2;
    const void* xs[2] = {du_hi, du_lo};
    const void* ws[2] = {w_hi, w_lo};
    const int Kw = K * K * Cy;
    for (int pl = 0; pl < npl; ++pl) {
        uint64_t dims[4] = {(uint64_t)Cy, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
        uint64_t strides[3] = {(uint64_t)Cy * 2, (uint64_t)Wo * Cy * 2, (uint64_t)Ho * Wo * Cy * 2};
        uint32_t box[4] = {64, (uint32_t)Wo, (uint32_t)TH, (uint32_t)TN};
        if (make_tmap(&p.tmA[pl][0], xs[pl], 4, dims, strides, box)) return GDRN_ERR_CUDA;
        uint64_t wdims[2] = {(uint64_t)Kw, (uint64_t)Cx_pad};
        uint64_t wstr[1] = {(uint64_t)Kw * 2};
        uint32_t wbox[2] = {64, (uint32_t)block_n};
        if (make_tmap(&p.tmB[pl], ws[pl], 2, wdims, wstr, wbox)) return GDRN_ERR_CUDA;
    }
    p.cluster = 1;
    p.mode = 1;
    p.M = N * Ho * Wo;
    p.N = Cx;
    p.num_m_tiles = num_m_tiles;
    p.num_n_tiles = Cx_pad / block_n;
    p.cchunks = Cy / 64;
    p.KW = K;
    p.pad = pad;
    p.stride = 1;
    p.TH = TH;
    p.TN = TN;
    p.tiles_per_img = (TN == 1) ? Ho / TH : 0;
    p.out_hi = dx_hi;
    p.out_lo = dx_lo;
    p.ldc = ldc;
    p.bias = bias;    // the same kernel is the FORWARD of ConvTranspose2d(k3, s2, p1, op1) (cdpn_rot_head_region.py:82-91): then
    p.act = act;      // bias / activation (folded eval BatchNorm) and the BatchNorm batch statistics apply to its output
    p.stats = stats;
    // phases in order of decreasing tap count (longest tiles first on the persistent CTAs)
    p.nphase = nphase;
    int nt = 0;
    const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};
    for (int q = 0; q < nphase; ++q) {
        const int a = K == 3 ? order[q][0] : 0, b = K == 3 ? order[q][1] : 0;
        p.ph_a[q] = a;
        p.ph_b[q] = b;
        p.ph_tap0[q] = nt;
        for (int r = 0; r < K; ++r) {
            if ((a + pad - r) & 1) continue;
            for (int s2 = 0; s2 < K; ++s2) {
                if ((b + pad - s2) & 1) continue;
                p.tap_dh[nt] = (signed char)((a + pad - r) / 2);
                p.tap_dw[nt] = (signed char)((b + pad - s2) / 2);
                p.tap_w[nt] = (signed char)((K - 1 - r) * K + (K - 1 - s2));  // pack_conv_dgrad stores flipped taps
                ++nt;
            }
        }
    }
    p.ph_tap0[nphase] = nt;
    p.num_kb = nt * p.cchunks;  // bookkeeping only (per-tile counts come from the phase table)
    for (p.pw_log2 = 0; (1 << p.pw_log2) < Wo; ++p.pw_log2) {}
    for (p.ph_log2 = 0; (1 << p.ph_log2) < Ho; ++p.ph_log2) {}
    return dispatch_gemm(p, block_n, nsplit, stream);
}

extern "C" int gdrn_gemm_fwd(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, void* y_hi,
                             void* y_lo, float* y_f32, const float* bias, float* stats, int M, int N, int N_pad, int K,
                             int ldc, int act, int nsplit, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (nsplit != 1 && nsplit != 3) return set_error(GDRN_ERR_ARG, "gemm_fwd: nsplit must be 1 or 3");
    if (nsplit == 3 && (a_lo == nullptr || w_lo == nullptr)) return set_error(GDRN_ERR_ARG, "gemm_fwd: lo planes missing");
    if (K % 64 != 0) return set_error(GDRN_ERR_ARG, "gemm_fwd: K=%d must be a multiple of 64", K);
    if (ldc % 8 != 0 || ldc < N) return set_error(GDRN_ERR_ARG, "gemm_fwd: bad ldc=%d", ldc);
    const int block_n = pick_block_n(N_pad, nsplit, (M + 127) / 128);
    if (block_n < 0) return set_error(GDRN_ERR_ARG, "gemm_fwd: N_pad=%d must be a multiple of 64", N_pad);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    const int cluster = 1;
    const int npl = nsplit == 1 ? 1 : 2;
    const void* as[2] = {a_hi, a_lo};
    const void* ws[2] = {w_hi, w_lo};
    for (int pl = 0; pl < npl; ++pl) {
        uint64_t adims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t astr[1] = {(uint64_t)K * 2};
        uint32_t abox[2] = {64, 128};
        if (make_tmap(&p.tmA[pl][0], as[pl], 2, adims, astr, abox)) return GDRN_ERR_CUDA;
        uint64_t wdims[2] = {(uint64_t)K, (uint64_t)N_pad};
        uint32_t wbox[2] = {64, (uint32_t)(block_n / cluster)};
        if (make_tmap(&p.tmB[pl], ws[pl], 2, wdims, astr, wbox)) return GDRN_ERR_CUDA;
    }
    p.cluster = cluster;
    p.mode = 0;
    p.M = M;
    p.N = N;
    p.num_m_tiles = (M + 127) / 128;
    p.num_n_tiles = N_pad / block_n;
    p.num_kb = K / 64;
    p.out_hi = y_hi;
    p.out_lo = y_lo;
    p.out_f32 = y_f32;
    p.ldc = ldc;
    p.bias = bias;
    p.act = act;
    p.stats = stats;
    return dispatch_gemm(p, block_n, nsplit, stream);
}

// BLOCK_N * 10 + NSPLIT (+ 10000 for the cta_group::2 kernel) of the most recent gdrn_conv_fwd / gdrn_gemm_fwd call on this thread
extern "C" int gdrn_last_gemm_variant() { return g_last_variant; }

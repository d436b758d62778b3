// tcgen05 weight-gradient kernel (sm_100a):  dW[co][tap][ci] = sum_p dY[p][co] * X[p (+) tap][ci]
//
// Both operands are "MN-major" for the tensor core: the reduction index is the pixel p, which is the
// SLOW index of the NHWC tensors, so the 64-pixel x 64-channel TMA boxes land in shared memory as
// [k = pixel][64 channels = 128 B] rows -- exactly the MN-major 128B-swizzled canonical layout
// (descriptor a_major = b_major = 1, SBO = 1024 B between 8-pixel groups, LBO = 8192 B between
// 64-channel chunks).  No transposes are ever materialised.
//
//   * A tile = dY  [64 pixels][128 co]      (2 boxes of 64 channels, plain 2-D view [P][Cout])
//   * B tile = X   [64 pixels][<=256 ci]    (<=4 boxes; for a conv the box is the tap-shifted 4-D
//                  window of the NHWC input, zero-filled outside the image; stride-2 convs read the
//                  four parity sub-lattices, like the forward kernel)
//   * one CTA = (co tile, ci tile, tap, K-split); fp32 accumulator in TMEM (<=256 columns);
//     partial results go to a [ksplit][Cout][taps*Cin] fp32 workspace that the unpack kernel
//     (pack.cu) reduces and transposes into the reference's OIHW gradient layout.
// Backward call sites replaced: autograd wgrad of every nn.Conv2d / ConvTranspose2d / nn.Linear in
// reference resnet_backbone.py, cdpn_rot_head_region.py, conv_pnp_net.py (cuDNN/cuBLAS today).
#include <stdlib.h>

#include <atomic>

#include "gdrn_internal.h"
#include "ptx.cuh"

namespace gdrn {

struct WgradParams {
    CUtensorMap tmA[2];     // dY  [plane]
    CUtensorMap tmB[2][4];  // X   [plane][phase]
    int mode;               // 0 = 2-D GEMM (B = [P][Ntot]), 1 = conv
    int Mvalid;             // Cout
    int Ntot;               // Cin (conv) or total B columns (gemm)
    int num_m_tiles, num_n_tiles, taps, ksplit;
    int kb_total;           // number of 64-pixel blocks
    int KW, pad, stride;
    int blocks_per_img, THk, TNk;
    float* ws;              // [ksplit][num_m_tiles*128][taps*Ntot]
    int tap_pack;           // 1: Cin == 64, stride 1: one work item covers a whole filter ROW, its KW taps packed along N
    int items;              // work items per (m_tile, n_tile, ksplit): taps, or KH when tap_pack
    int ws_ld;              // workspace row length in floats (= KH*KW*Cin)
};

constexpr int kWgStageA = 2 * 8192;
constexpr int kWgThreads = 64 + 256;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-9: two epilogue warpgroups (column halves)

template <int N_TILE, int NSPLIT>
struct WgradCfg {
    static constexpr int NPL = (NSPLIT == 1) ? 1 : 2;
    static constexpr int NCH = N_TILE / 64;
    static constexpr int A_BYTES = kWgStageA;
    static constexpr int B_BYTES = NCH * 8192;
    static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
    static constexpr int STAGES_RAW = (227 * 1024 - 2048) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2048;
    // round-toward-zero accumulation (see gemm_fwd.cu): cross terms and three round-robin hi*hi partials
    static constexpr int NMAIN = (NSPLIT == 3) ? 3 : 1;
    static constexpr int NACC = (NSPLIT == 3) ? 4 : 1;
    static constexpr int TMEM_RAW = NACC * N_TILE;
    static constexpr int TMEM_COLS = TMEM_RAW <= 32 ? 32 : TMEM_RAW <= 64 ? 64 : TMEM_RAW <= 128 ? 128 : TMEM_RAW <= 256 ? 256 : 512;
    static_assert(NACC * N_TILE <= 512, "TMEM overflow");
    static_assert(STAGES >= 2, "pipeline too shallow");
};

template <int N_TILE, int NSPLIT>
__global__ void __launch_bounds__(kWgThreads, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
    using Cfg = WgradCfg<N_TILE, NSPLIT>;
    constexpr int NPL = Cfg::NPL;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NCH = Cfg::NCH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // work item decode: m fastest, then n, tap, ksplit
    int item = blockIdx.x;
    const int m_tile = item % p.num_m_tiles;
    item /= p.num_m_tiles;
    const int n_tile = item % p.num_n_tiles;
    item /= p.num_n_tiles;
    const int tap = item % p.items;  // filter tap, or filter row when tap_pack
    const int ks = item / p.items;
    const int kb_per = (p.kb_total + p.ksplit - 1) / p.ksplit;
    const int kb_begin = ks * kb_per;
    const int kb_end = min(p.kb_total, kb_begin + kb_per);
    const int nkb = kb_end - kb_begin;  // may be <= 0 for a ragged last split: then we store zeros

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue under our tail
    pdl_wait();               // nothing above touched global memory; from here on we read what the predecessor wrote

    // Producer and MMA issuer run their loops with the WHOLE warp (warp-uniform control flow, operands in uniform registers);
    // one elected lane issues the TMA / tcgen05 instructions.  Per k-block nothing but the barrier wait, the expect_tx and the
    // issues remains (the first version rebuilt coordinates with divisions and eight descriptors per k-block on one thread,
    // which -- not the tensor pipe -- set the pace; see gemm_fwd.cu).
    if (warp == 0) {
        if (nkb > 0) {
            int dh = 0, dw = 0, map = 0;
            if (p.mode == 1 && p.tap_pack) {
                dh = tap - p.pad;  // `tap` is the filter row; the column shift varies per 64-channel chunk below
            } else if (p.mode == 1) {
                const int r = tap / p.KW;
                const int s = tap - r * p.KW;
                dh = r - p.pad;
                dw = s - p.pad;
                if (p.stride == 2) {
                    map = ((dh & 1) << 1) | (dw & 1);
                    dh >>= 1;
                    dw >>= 1;
                }
            }
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            const uint64_t tmA0 = reinterpret_cast<uint64_t>(&p.tmA[0]), tmA1 = reinterpret_cast<uint64_t>(&p.tmA[1]);
            const uint64_t tmB0 = reinterpret_cast<uint64_t>(&p.tmB[0][map]), tmB1 = reinterpret_cast<uint64_t>(&p.tmB[1][map]);
            const int mode = p.mode, tap_pack = p.tap_pack, padv = p.pad, bpi = p.blocks_per_img, THk = p.THk, TNk = p.TNk;
            const int m0 = m_tile * 128, nb0 = n_tile * N_TILE;
            // (image, 64-pixel block inside the image) of the first k-block, advanced incrementally
            int n0 = 0, hb = 0;
            if (mode == 1) {
                if (TNk == 1) {
                    n0 = kb_begin / bpi;
                    hb = kb_begin - n0 * bpi;
                } else {
                    n0 = kb_begin * TNk;
                }
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                const int hh = hb * THk + dh, prow = kb * 64;
                mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx_u32(fb, Cfg::STAGE_BYTES);
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) {
                        const uint32_t a_dst = dst + pl * Cfg::A_BYTES;
                        const uint64_t tmA = pl ? tmA1 : tmA0, tmB = pl ? tmB1 : tmB0;
                        tma_load_2d_u32(a_dst, tmA, fb, m0, prow);
                        tma_load_2d_u32(a_dst + 8192, tmA, fb, m0 + 64, prow);
                        const uint32_t b_dst = dst + NPL * Cfg::A_BYTES + pl * Cfg::B_BYTES;
                        if (mode == 1) {
                            if (tap_pack) {
#pragma unroll
                                for (int c = 0; c < NCH; ++c)  // chunk c = filter column c (all 64 input channels)
                                    tma_load_4d_u32(b_dst + c * 8192, tmB, fb, 0, c - padv, hh, n0);
                            } else {
#pragma unroll
                                for (int c = 0; c < NCH; ++c)
                                    tma_load_4d_u32(b_dst + c * 8192, tmB, fb, nb0 + c * 64, dw, hh, n0);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < NCH; ++c) tma_load_2d_u32(b_dst + c * 8192, tmB, fb, nb0 + c * 64, prow);
                        }
                    }
                }
                __syncwarp();
                if (mode == 1) {
                    if (TNk == 1) {
                        if (++hb == bpi) {
                            hb = 0;
                            ++n0;
                        }
                    } else {
                        n0 += TNk;
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (nkb > 0) {
            constexpr uint32_t idesc = make_idesc(128, N_TILE, 1, 1);
            const uint64_t desc_const = make_smem_desc(0, 8192, 1024);  // only the 14-bit address field varies
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            int stage = 0, main_idx = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
                const uint64_t da = desc_const | (uint64_t)(a_addr >> 4);
                const uint64_t db = da + ((NPL * Cfg::A_BYTES) >> 4);
                mbar_wait_u32(full0 + stage * 8, phase);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 16 pixels (= 16 rows of 128 B = 2048 B) per MMA
                        if (NSPLIT == 3) {
                            umma_bf16(tmem_base + main_idx * N_TILE, da + 128 * k, db + 128 * k, idesc,
                                      (kb >= Cfg::NMAIN || k != 0) ? 1u : 0u);
                            umma_bf16(tmem_base + Cfg::NMAIN * N_TILE, da + 128 * k, db + (Cfg::B_BYTES >> 4) + 128 * k, idesc,
                                      (kb | k) != 0 ? 1u : 0u);
                            umma_bf16(tmem_base + Cfg::NMAIN * N_TILE, da + (Cfg::A_BYTES >> 4) + 128 * k, db + 128 * k, idesc, 1u);
                        } else {
                            umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit_u32(empty0 + stage * 8);
                }
                __syncwarp();
                if (++main_idx == Cfg::NMAIN) main_idx = 0;
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit(done_bar);
            __syncwarp();
        }
    } else {
        // epilogue warps 2..9: TMEM lane quadrant = warp % 4; the CTA's single tile is drained by two warpgroups, each
        // taking half of the 32-column chunks (one work item per CTA: the epilogue is fully exposed)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int co = m_tile * 128 + row;
        if (nkb > 0) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        const size_t ld = (size_t)p.ws_ld;
        // tap_pack: columns (filter column, ci) of filter row `tap` are contiguous in the [co][tap][ci] workspace row
        const size_t col0 = p.tap_pack ? (size_t)tap * N_TILE : (size_t)tap * p.Ntot + (size_t)n_tile * N_TILE;
        float* dst_row = p.ws + ((size_t)ks * p.num_m_tiles * 128 + co) * ld + col0;
#pragma unroll 1
        for (int c = half * (N_TILE / 64); c < (half + 1) * (N_TILE / 64); ++c) {
            uint32_t raw[32];
            if (nkb > 0) {
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + c * 32;
                tmem_ld_32x32(t0, raw);
                tmem_ld_wait();
                if (NSPLIT == 3) {
                    const int nmain = nkb < Cfg::NMAIN ? nkb : Cfg::NMAIN;
                    for (int a2 = 1; a2 <= Cfg::NMAIN; ++a2) {
                        if (a2 < Cfg::NMAIN && a2 >= nmain) continue;
                        uint32_t r2[32];
                        tmem_ld_32x32(t0 + a2 * N_TILE, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            raw[j] = __float_as_uint(fmaf(__uint_as_float(r2[j]), a2 == Cfg::NMAIN ? kLoInvScale : 1.f, __uint_as_float(raw[j])));
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) raw[j] = 0u;
            }
            if (p.tap_pack || n_tile * N_TILE + c * 32 < p.Ntot) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(dst_row + c * 32 + j) = make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 variant (single-plane operands): a CTA PAIR owns one 256 (co) x N_TILE (ci) tile of a (tap, K-split) work item.
// Same reasoning as gemm_fwd2.cu: per 64-pixel k-block the 1-CTA kernel's shared memory serves 48 KB of TMA writes + 4 x 12 KB
// of MMA operand reads = 96 KB for 512 MMA cycles (187 B/clk against 128 B/clk); in a pair each CTA stages its own 128 output
// channels of dY but only HALF of the X columns: 32 KB + 4 x 8 KB = 64 KB (125 B/clk).  Eligible: Cout a multiple of 256 (head,
// layer3, layer4, deconv), no tap packing.  Identical arithmetic / accumulation order to the 1-CTA kernel (bit-equal, tested).
// MEASURED (profiles/r2_wgrad_2cta_ab.txt): alone under ncu the two long-K head layers go from 76 % to 84 % tensor pipe (210 ->
// 197 us), every short-K layer (layer3 / layer4, 32 k-blocks per CTA) stays at 33 % -- those are bound by the fixed cost per
// work item (prologue + 128 KB fp32 partial per CTA + its unpack), not by shared memory.  In the real step the pair kernel runs
// on the weight-gradient side stream next to main-stream kernels, where a cluster needs BOTH SMs of a TPC free at once: A/B
// 10.99 / 10.93 ms (off) vs 11.05 / 11.00 ms (on).  Hence default OFF; GDRN_WGRAD_2CTA=1 / gdrn_set_wgrad_2cta(1) enables it.
template <int N_TILE>
struct Wgrad2Cfg {
    static constexpr int NCH = N_TILE / 64;      // 64-column chunks of the pair tile
    static constexpr int NCH_H = NCH / 2;        // chunks staged by one CTA
    static constexpr int A_BYTES = kWgStageA;    // dY [64 px][128 co]
    static constexpr int B_BYTES = NCH_H * 8192;  // X  [64 px][N_TILE / 2 ci]
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (227 * 1024 - 2048) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2048;
    static_assert(NCH % 2 == 0 && N_TILE <= 256, "pair tile: 128 or 256 input channels");
};

template <int N_TILE>
__global__ void __launch_bounds__(kWgThreads, 1) gemm_wgrad2_kernel(const __grid_constant__ WgradParams p) {
    using Cfg = Wgrad2Cfg<N_TILE>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NCH_H = Cfg::NCH_H;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);  // leader only
    uint64_t* empty_bar = full_bar + STAGES;                // per CTA
    uint64_t* done_bar = empty_bar + STAGES;                // per CTA (multicast commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    // work item decode (per PAIR): m pair fastest, then n, tap, ksplit
    int item = blockIdx.x >> 1;
    const int m_pairs = p.num_m_tiles / 2;
    const int m_tile = (item % m_pairs) * 2 + (int)crank;
    item /= m_pairs;
    const int n_tile = item % p.num_n_tiles;
    item /= p.num_n_tiles;
    const int tap = item % p.items;
    const int ks = item / p.items;
    const int kb_per = (p.kb_total + p.ksplit - 1) / p.ksplit;
    const int kb_begin = ks * kb_per;
    const int kb_end = min(p.kb_total, kb_begin + kb_per);
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(N_TILE <= 128 ? 128 : 256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue under our tail
    pdl_wait();               // nothing above touched global memory; from here on we read what the predecessor wrote

    if (warp == 0) {
        if (nkb > 0) {
            int dh = 0, dw = 0, map = 0;
            if (p.mode == 1) {
                const int r = tap / p.KW;
                const int s2 = tap - r * p.KW;
                dh = r - p.pad;
                dw = s2 - p.pad;
                if (p.stride == 2) {
                    map = ((dh & 1) << 1) | (dw & 1);
                    dh >>= 1;
                    dw >>= 1;
                }
            }
            const uint32_t smem_base = smem_u32(smem), empty0 = smem_u32(empty_bar);
            const uint32_t full0_local = smem_u32(full_bar), full0 = mapa_cta0(full0_local);
            const uint64_t tmA0 = reinterpret_cast<uint64_t>(&p.tmA[0]);
            const uint64_t tmB0 = reinterpret_cast<uint64_t>(&p.tmB[0][map]);
            const int mode = p.mode, bpi = p.blocks_per_img, THk = p.THk, TNk = p.TNk;
            const int m0 = m_tile * 128, nb0 = n_tile * N_TILE + (int)crank * (N_TILE / 2);
            int n0 = 0, hb = 0;
            if (mode == 1) {
                if (TNk == 1) {
                    n0 = kb_begin / bpi;
                    hb = kb_begin - n0 * bpi;
                } else {
                    n0 = kb_begin * TNk;
                }
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const uint32_t fb = full0 + stage * 8, dst = smem_base + stage * Cfg::STAGE_BYTES;
                const int hh = hb * THk + dh, prow = kb * 64;
                mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx_u32(full0_local + stage * 8, 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
                    tma2_load_2d_u32(dst, tmA0, fb, m0, prow);
                    tma2_load_2d_u32(dst + 8192, tmA0, fb, m0 + 64, prow);
                    const uint32_t b_dst = dst + Cfg::A_BYTES;
                    if (mode == 1) {
#pragma unroll
                        for (int c = 0; c < NCH_H; ++c) tma2_load_4d_u32(b_dst + c * 8192, tmB0, fb, nb0 + c * 64, dw, hh, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < NCH_H; ++c) tma2_load_2d_u32(b_dst + c * 8192, tmB0, fb, nb0 + c * 64, prow);
                    }
                }
                __syncwarp();
                if (mode == 1) {
                    if (TNk == 1) {
                        if (++hb == bpi) {
                            hb = 0;
                            ++n0;
                        }
                    } else {
                        n0 += TNk;
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (leader && nkb > 0) {
            constexpr uint32_t idesc = make_idesc(256, N_TILE, 1, 1);
            const uint64_t desc_const = make_smem_desc(0, 8192, 1024);
            const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
                const uint64_t da = desc_const | (uint64_t)(a_addr >> 4);
                const uint64_t db = da + (Cfg::A_BYTES >> 4);
                mbar_wait_u32(full0 + stage * 8, phase);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma2(tmem_base, da + 128 * k, db + 128 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma2_commit_u32(empty0 + stage * 8);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma2_commit_u32(smem_u32(done_bar));
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int co = m_tile * 128 + row;
        if (nkb > 0) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        const size_t ld = (size_t)p.ws_ld;
        const size_t col0 = (size_t)tap * p.Ntot + (size_t)n_tile * N_TILE;
        float* dst_row = p.ws + ((size_t)ks * p.num_m_tiles * 128 + co) * ld + col0;
#pragma unroll 1
        for (int c = half * (N_TILE / 64); c < (half + 1) * (N_TILE / 64); ++c) {
            uint32_t raw[32];
            if (nkb > 0) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, raw);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) raw[j] = 0u;
            }
            if (n_tile * N_TILE + c * 32 < p.Ntot) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(dst_row + c * 32 + j) = make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may not exit (or free TMEM) while the leader's MMAs still read its smem / write its TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(N_TILE <= 128 ? 128 : 256) : "memory");
    }
}

template <int N_TILE>
static int launch_wgrad2(const WgradParams& p, cudaStream_t stream) {
    using Cfg = Wgrad2Cfg<N_TILE>;
    auto kern = gemm_wgrad2_kernel<N_TILE>;
    static bool attr_set = false;
    if (!attr_set) {
        GDRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.num_m_tiles * p.num_n_tiles * p.items * p.ksplit);  // (m_tiles / 2) pairs x 2 CTAs
    cfg.blockDim = dim3(kWgThreads);
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

static int g_wgrad_2cta = -1;  // GDRN_WGRAD_2CTA=1 enables the pair kernel (default off: neutral-to-slower in situ, see header comment)
static std::atomic<long> g_wgrad_2cta_launches{0};
static bool want_wgrad_2cta(const WgradParams& p, int n_tile, int nsplit) {
    if (g_wgrad_2cta < 0) {
        const char* e = getenv("GDRN_WGRAD_2CTA");
        g_wgrad_2cta = e ? atoi(e) : 0;
    }
    return g_wgrad_2cta == 1 && nsplit == 1 && !p.tap_pack && (p.num_m_tiles % 2) == 0 && (n_tile == 256 || n_tile == 128) &&
           (p.Mvalid % 256) == 0;
}

template <int N_TILE, int NSPLIT>
static int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
    using Cfg = WgradCfg<N_TILE, NSPLIT>;
    auto kern = gemm_wgrad_kernel<N_TILE, NSPLIT>;
    static bool attr_set = false;
    if (!attr_set) {
        GDRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.num_m_tiles * p.num_n_tiles * p.items * p.ksplit);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
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

static int dispatch_wgrad(const WgradParams& p, int n_tile, int nsplit, cudaStream_t stream) {
    if (want_wgrad_2cta(p, n_tile, nsplit)) {
        g_wgrad_2cta_launches.fetch_add(1, std::memory_order_relaxed);
        return n_tile == 256 ? launch_wgrad2<256>(p, stream) : launch_wgrad2<128>(p, stream);
    }
    if (nsplit == 1) {
        if (n_tile == 256) return launch_wgrad<256, 1>(p, stream);
        if (n_tile == 192) return launch_wgrad<192, 1>(p, stream);
        if (n_tile == 128) return launch_wgrad<128, 1>(p, stream);
        if (n_tile == 64) return launch_wgrad<64, 1>(p, stream);
    } else if (nsplit == 3) {
        if (n_tile == 128) return launch_wgrad<128, 3>(p, stream);
        if (n_tile == 64) return launch_wgrad<64, 3>(p, stream);
    }
    return set_error(GDRN_ERR_ARG, "wgrad: unsupported n_tile=%d nsplit=%d", n_tile, nsplit);
}

static int pick_n_tile(int ntot, int nsplit) {
    if (nsplit == 1 && ntot % 256 == 0) return 256;
    if (ntot % 128 == 0) return 128;
    if (ntot % 64 == 0) return 64;
    return -1;
}

}  // namespace gdrn

using namespace gdrn;

// Workspace size query + K-split choice shared by conv and gemm wgrad.
static int choose_ksplit(int base_items, int kb_total, int requested) {
    if (requested > 0) return requested < kb_total ? requested : kb_total;
    // one full wave of CTAs (each CTA is one (tile, tap, K-split) work item and owns an SM: 1 CTA/SM by smem), with at
    // least 8 k-blocks per split so the pipeline fill / TMEM drain is amortised
    int ks = num_sms() / base_items;
    int max_ks = kb_total / 8;
    if (max_ks < 1) max_ks = 1;
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    return ks;
}

extern "C" int gdrn_set_wgrad_2cta(int on) {
    gdrn::g_wgrad_2cta = on ? 1 : 0;
    return 0;
}
extern "C" long gdrn_wgrad_2cta_launch_count() { return gdrn::g_wgrad_2cta_launches.load(); }

extern "C" int gdrn_conv_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* ws,
                               long ws_floats, long* ws_need_out, int* ksplit_out, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                               int stride, int pad, int ksplit, int nsplit, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (nsplit != 1 && nsplit != 3) return set_error(GDRN_ERR_ARG, "conv_wgrad: nsplit must be 1 or 3");
    if (ws != nullptr && nsplit == 3 && (dy_lo == nullptr || x_lo == nullptr)) return set_error(GDRN_ERR_ARG, "conv_wgrad: lo planes missing");
    if (Cin % 64 || Cout % 64) return set_error(GDRN_ERR_ARG, "conv_wgrad: Cin/Cout must be multiples of 64");
    if (stride != 1 && stride != 2) return set_error(GDRN_ERR_ARG, "conv_wgrad: stride must be 1 or 2");
    const int Ho = H / stride, Wo = W / stride;
    if (Wo > 64 || 64 % Wo != 0) return set_error(GDRN_ERR_ARG, "conv_wgrad: unsupported output width %d", Wo);
    int THk = 64 / Wo;
    if (THk > Ho) THk = Ho;
    const int TNk = 64 / (Wo * THk);
    if (Ho % THk != 0 || (N % TNk) != 0) return set_error(GDRN_ERR_ARG, "conv_wgrad: shape not tileable by 64 pixels");
    const long P = (long)N * Ho * Wo;
    if (P % 64 != 0) return set_error(GDRN_ERR_ARG, "conv_wgrad: N*Ho*Wo must be a multiple of 64");
    // Cin == 64 (layer1): a 64-wide N tile leaves the tensor core smem/ingress-bound; pack the KW taps of a filter row
    // along N instead (N = 192): one dY tile load feeds three taps, 3x fewer MMAs / A loads per FLOP
    const int tap_pack = (nsplit == 1 && Cin == 64 && stride == 1 && KW == 3) ? 1 : 0;
    const int n_tile = tap_pack ? 192 : pick_n_tile(Cin, nsplit);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 1;
    p.tap_pack = tap_pack;
    p.items = tap_pack ? KH : KH * KW;
    p.ws_ld = KH * KW * Cin;
    p.Mvalid = Cout;
    p.Ntot = Cin;
    p.num_m_tiles = (Cout + 127) / 128;
    p.num_n_tiles = tap_pack ? 1 : Cin / n_tile;
    p.taps = KH * KW;
    p.kb_total = (int)(P / 64);
    p.ksplit = choose_ksplit(p.num_m_tiles * p.num_n_tiles * p.items, p.kb_total, ksplit);
    p.KW = KW;
    p.pad = pad;
    p.stride = stride;
    p.THk = THk;
    p.TNk = TNk;
    p.blocks_per_img = (TNk == 1) ? Ho / THk : 0;
    p.ws = ws;
    if (ksplit_out) *ksplit_out = p.ksplit;
    const long need = (long)p.ksplit * p.num_m_tiles * 128 * p.taps * Cin;
    if (ws_need_out) *ws_need_out = need;
    if (ws == nullptr) return 0;  // size query only
    if (ws_floats < need) return set_error(GDRN_ERR_ARG, "conv_wgrad: workspace too small (%ld < %ld floats)", ws_floats, need);
    const int npl = nsplit == 1 ? 1 : 2;
    const void* dys[2] = {dy_hi, dy_lo};
    const void* xs[2] = {x_hi, x_lo};
    for (int pl = 0; pl < npl; ++pl) {
        uint64_t adims[2] = {(uint64_t)Cout, (uint64_t)P};
        uint64_t astr[1] = {(uint64_t)Cout * 2};
        uint32_t abox[2] = {64, 64};
        if (make_tmap(&p.tmA[pl], dys[pl], 2, adims, astr, abox)) return GDRN_ERR_CUDA;
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xs[pl]);
        for (int ph = 0; ph < (stride == 2 ? 4 : 1); ++ph) {
            const int phh = ph >> 1, phw = ph & 1;
            uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
            uint64_t strides[3] = {(uint64_t)stride * Cin * 2, (uint64_t)stride * W * Cin * 2, (uint64_t)H * W * Cin * 2};
            uint32_t box[4] = {64, (uint32_t)Wo, (uint32_t)THk, (uint32_t)TNk};
            if (make_tmap(&p.tmB[pl][ph], xb + ((size_t)phh * W + phw) * Cin, 4, dims, strides, box)) return GDRN_ERR_CUDA;
        }
    }
    return dispatch_wgrad(p, n_tile, nsplit, stream);
}

// dW[M][Ntot] = sum_p A[p][M] * B[p][Ntot]   (A = dY [P][M], B = X [P][Ntot]; P zero-padded to 64 by TMA)
extern "C" int gdrn_gemm_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* ws,
                               long ws_floats, long* ws_need_out, int* ksplit_out, long P, int M, int Ntot, int ksplit, int nsplit,
                               void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (nsplit != 1 && nsplit != 3) return set_error(GDRN_ERR_ARG, "gemm_wgrad: nsplit must be 1 or 3");
    if (ws != nullptr && nsplit == 3 && (dy_lo == nullptr || x_lo == nullptr)) return set_error(GDRN_ERR_ARG, "gemm_wgrad: lo planes missing");
    if (M % 8 || Ntot % 64) return set_error(GDRN_ERR_ARG, "gemm_wgrad: M %% 8 and Ntot %% 64 required");
    const int n_tile = pick_n_tile(Ntot, nsplit);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.mode = 0;
    p.Mvalid = M;
    p.Ntot = Ntot;
    p.num_m_tiles = (M + 127) / 128;
    p.num_n_tiles = Ntot / n_tile;
    p.taps = 1;
    p.items = 1;
    p.ws_ld = Ntot;
    p.kb_total = (int)((P + 63) / 64);
    p.ksplit = choose_ksplit(p.num_m_tiles * p.num_n_tiles, p.kb_total, ksplit);
    p.ws = ws;
    if (ksplit_out) *ksplit_out = p.ksplit;
    const long need = (long)p.ksplit * p.num_m_tiles * 128 * Ntot;
    if (ws_need_out) *ws_need_out = need;
    if (ws == nullptr) return 0;
    if (ws_floats < need) return set_error(GDRN_ERR_ARG, "gemm_wgrad: workspace too small (%ld < %ld floats)", ws_floats, need);
    const int npl = nsplit == 1 ? 1 : 2;
    const void* dys[2] = {dy_hi, dy_lo};
    const void* xs[2] = {x_hi, x_lo};
    for (int pl = 0; pl < npl; ++pl) {
        uint64_t adims[2] = {(uint64_t)M, (uint64_t)P};
        uint64_t astr[1] = {(uint64_t)M * 2};
        uint32_t abox[2] = {64, 64};
        if (make_tmap(&p.tmA[pl], dys[pl], 2, adims, astr, abox)) return GDRN_ERR_CUDA;
        uint64_t bdims[2] = {(uint64_t)Ntot, (uint64_t)P};
        uint64_t bstr[1] = {(uint64_t)Ntot * 2};
        if (make_tmap(&p.tmB[pl][0], xs[pl], 2, bdims, bstr, abox)) return GDRN_ERR_CUDA;
    }
    return dispatch_wgrad(p, n_tile, nsplit, stream);
}

// attn_sm100.cu — flash-style joint attention forward on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces F.scaled_dot_product_attention(attn_mask=None, dropout_p=0, is_causal=False) at the MoT joint
// attention call sites (transformer_wan_mot.py:637-644, cogvideox_transformer_3d_mot.py:424-431) and the
// per-stream cross-attention (transformer_wan_mot.py:163-179).  q/k/v/o are addressed through explicit
// (batch, head, token) element strides, so the kernel reads Q, K and V straight out of the fused QKV GEMM
// output [tokens, 3, H, D] of BOTH streams (no torch.cat, no head-major transpose) and writes O token-major
// [tokens, H*D], which is the A operand of the output projection.
//
// One CTA = one (batch, head, 256 query rows) work item = two 128-row Q tiles that ping-pong on the tensor core: while
// the softmax warps of tile i work, the MMAs of tile 1-i run.  18 warps:
//   warp 0     : TMEM allocator, then TMA producer — Q once, then K_j / V_j tiles into a ring of 128xD bf16 stages (SWIZZLE_128B)
//   warp 1     : MMA issuer   — S_i = Q_i K_j^T (SS, both K-major, N = 128), O_i += P_i V_j (A = P from TMEM, B = V MN-major)
//   TMEM       : S0 | S1 | O0 | O1, fp32 columns; P_i is written as packed bf16 over columns [0, 64) of S_i
//   warps 2-9  : softmax of Q tile 0;  warps 10-17: softmax of Q tile 1.
// Measured floors that shape this (tools/mma_rate_probe.cu, tools/softmax_pipe_probe.cu): an M128 N<=128 K16 MMA costs
// ~71 clk whatever N is (so N = 128 tiles; a double-buffered N = 64 variant ran at half the tensor rate), and ONE softmax
// warp per SM sub-partition issues at IPC ~0.35, two at ~0.5, four at ~0.6 — a tile handled by four warps (one thread per
// row) takes longer than the MMAs it overlaps with.  So each tile's softmax runs on EIGHT warps, two per sub-partition, using the 16-lane TMEM
// shapes: warp (q, hl) owns TMEM lanes 32 q + 16 hl .. +15 (tcgen05.ld.16x256b — the mma.sync accumulator layout, a row
// lives in one quad; tcgen05.st.16x128b writes the packed P in the layout the MMA reads).  No cross-warp exchange exists:
//   * the common step has NO max reduction: p = 2^(s c - m_used c) is computed speculatively against the running reference
//     m_used, and the warp votes afterwards on whether a score sat more than 2^8 above it (a thread's partial row sum
//     exceeds 2^8; ex2.approx overflows cleanly to +inf).  Only when the vote fires (first half-tile, or a new maximum
//     far above the reference) the quad reduces the row maxima by shuffle, the warp rescales its 16 rows of O (lazy
//     rescale) and repeats the pass;
//   * P is published in two 64-column halves (p_full(i, 0/1)), so the first four PV MMAs of a step overlap the second
//     half of the softmax.
// The last KV tile is masked against Lkv (TMA zero-fills out-of-range K/V rows).
#include <cstdlib>
#include <type_traits>
#include "vap_kernels.cuh"

namespace vap {

constexpr int kSoftmaxWarps = 16;  // eight per Q tile
constexpr int kFirstSoftmaxWarp = 2;
constexpr int kAttnThreads = (kFirstSoftmaxWarp + kSoftmaxWarps) * 32;
constexpr int kBlockM = 128;  // rows per Q tile
constexpr int kBlockN = 128;  // kv rows per tile
constexpr float kRescaleThreshold = 8.0f;
constexpr float kSumTrigger = 256.0f;  // 2^kRescaleThreshold
// Of every 8 (p0,p1) pairs, how many take the FMA-pipe polynomial exp2 instead of MUFU.EX2.  Measured (tools/attn_variants.sh):
// at D = 128 the loop is issue-bound and every polynomial pair costs TFLOP/s; at D = 64 (half the MMA work per score) the
// MUFU is the bound and 1 of 8 is best.
#ifndef VAP_ATTN_POLY_PAIRS_D128
#define VAP_ATTN_POLY_PAIRS_D128 0
#endif
#ifndef VAP_ATTN_POLY_PAIRS_D64
#define VAP_ATTN_POLY_PAIRS_D64 1
#endif
#ifndef VAP_ATTN_LEADER_WAIT
#define VAP_ATTN_LEADER_WAIT 1
#endif
// Which waits of the critical chain busy-poll (mbarrier.test_wait) instead of suspending in try_wait: bit 0 = the MMA issuer's waits for P,
// bit 1 = the softmax warps' wait for S.
#ifndef VAP_ATTN_SPIN
#define VAP_ATTN_SPIN 0
#endif
#define VAP_MMA_WAIT(bar, par) do { if (VAP_ATTN_SPIN & 1) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)
#define VAP_SM_WAIT(bar, par) do { if (VAP_ATTN_SPIN & 2) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)
#ifndef VAP_ATTN_MMA_ORDER
#define VAP_ATTN_MMA_ORDER 1  // 1: the issuer's bookkeeping waits / releases are kept off the P -> PV path (see attn_mma_warp)
#endif
// Three-stage publish of P (16-lane kernel): half 0 | the first VAP_ATTN_P3_PAIRS pairs per thread of half 1 (4 kv columns each: 12 -> 48
// columns) | the rest.  What follows the softmax of a tile on its dependency chain is then the PV MMAs of the LAST stage + QK^T of the next
// step: 1 + 8 MMAs instead of 4 + 8 (D = 128).
#ifndef VAP_ATTN_P3
#define VAP_ATTN_P3 0
#endif
#ifndef VAP_ATTN_P3_PAIRS
#define VAP_ATTN_P3_PAIRS 12  // 8 or 12
#endif
#ifndef VAP_ATTN_TRACE
#define VAP_ATTN_TRACE 0  // 1: clock64() stamps of CTA (0,0,0) when a trace buffer is installed (tools/attn_trace.py)
#endif

// T = Q tiles per CTA: 2 (the ping-pong kernels: 256 query rows, all 512 TMEM columns, one CTA per SM) or 1 (the short-KV kernel: 128 query rows,
// 256 TMEM columns, a two-slot K / V ring, so that TWO CTAs share an SM).
template <int D, int T = 2>
struct AttnCfg {
    static constexpr int kQTiles = T;
    static constexpr int kTileBytes = 128 * D * 2;   // one Q / K / V tile
    static constexpr int kHalfBytes = 128 * 64 * 2;  // one 64-column (128-byte) swizzle slab
    static constexpr int kHalves = D / 64;
    static constexpr int kKvStages = (T == 2) ? ((D == 128) ? 5 : 8) : ((D == 128) ? 2 : 4);
    static constexpr int kBarBytes = 512;
    static constexpr int kSmemBytes = T * kTileBytes + kKvStages * kTileBytes + kBarBytes + 1024;
    static constexpr int kTmemCols = (T == 2) ? 512 : 256;
    static constexpr int kColS0 = 0, kColS1 = 128, kColO0 = (T == 2) ? 256 : 128, kColO1 = 256 + D;
};

// bf16(float(a) + float(b)) per element of a packed pair: what a bf16 tensor add computes
__device__ __forceinline__ uint32_t add_bf16x2_as_tensors(uint32_t a, uint32_t b) {
    const float2 fa = bf16x2_to_float2(a), fb = bf16x2_to_float2(b);
    return pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
}

// Shared-memory map of one CTA: Q tiles | K/V ring | mbarriers.
template <int D, int T = 2>
struct AttnSmem {
    using Cfg = AttnCfg<D, T>;
    uint32_t q_smem, kv_smem, bar_base;
    __device__ __forceinline__ explicit AttnSmem(uint32_t smem_base)
        : q_smem(smem_base), kv_smem(smem_base + T * Cfg::kTileBytes), bar_base(smem_base + T * Cfg::kTileBytes + Cfg::kKvStages * Cfg::kTileBytes) {}
    __device__ __forceinline__ uint32_t kv_full(int s) const { return bar_base + 8u * s; }
    __device__ __forceinline__ uint32_t kv_empty(int s) const { return bar_base + 8u * (Cfg::kKvStages + s); }
    __device__ __forceinline__ uint32_t q_full() const { return bar_base + 8u * (2 * Cfg::kKvStages); }
    __device__ __forceinline__ uint32_t s_full(int i) const { return bar_base + 8u * (2 * Cfg::kKvStages + 1 + i); }            // S_i(j) in TMEM (and PV_i(j-1) done)
    __device__ __forceinline__ uint32_t p_full(int i, int c) const { return bar_base + 8u * (2 * Cfg::kKvStages + 3 + 2 * i + c); }  // half c of P_i(j) in TMEM
    __device__ __forceinline__ uint32_t pv_half(int i) const { return bar_base + 8u * (2 * Cfg::kKvStages + 7 + i); }           // the PV MMAs of half 0 have completed
    __device__ __forceinline__ uint32_t o_done(int i) const { return bar_base + 8u * (2 * Cfg::kKvStages + 9 + i); }
    __device__ __forceinline__ uint32_t tmem_ptr_addr() const { return bar_base + 8u * (2 * Cfg::kKvStages + 11); }
    __device__ __forceinline__ uint32_t p_last(int i) const { return bar_base + 8u * (2 * Cfg::kKvStages + 12 + i); }  // three-stage publish: the last stage of P_i(j)
    __device__ __forceinline__ uint32_t pv_mid(int i) const { return bar_base + 8u * (2 * Cfg::kKvStages + 14 + i); }  // three-stage publish: the PV MMAs of stage 1 have completed
    // one thread: p_arrivals = softmax warps per Q tile (each arrives once per published half)
    __device__ __forceinline__ void init_barriers(int cluster_size, int p_arrivals) const {
        for (int s = 0; s < Cfg::kKvStages; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), cluster_size);  // one tcgen05.commit per CTA of the cluster
        }
        mbar_init(q_full(), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(p_full(i, 0), p_arrivals);
            mbar_init(p_full(i, 1), p_arrivals);
            mbar_init(pv_half(i), 1);
            mbar_init(o_done(i), 1);
            mbar_init(p_last(i), p_arrivals);
            mbar_init(pv_mid(i), 1);
        }
        fence_mbar_init();
    }
};

// The producer and MMA warps run their loops warp-wide (all lanes wait on the mbarriers, one elected lane issues): control flow and
// descriptors stay warp-uniform, so ptxas keeps them in uniform registers instead of wrapping every UTCHMMA / UTMALDG in an R2UR
// waterfall loop.
// CL = 2: clusters of two CTAs (adjacent 256-row query blocks of the same head) share every K / V tile: each CTA fetches half of
// the tile's rows and TMA multicasts them into both CTAs' shared memory, so K / V cross the L2 -> SM fabric once per 512 query
// rows; a ring slot is reusable when BOTH CTAs' MMAs have read it (multicast tcgen05.commit on the empty barriers).
template <int D, int CL, int T = 2>
__device__ __forceinline__ void attn_producer_warp(const AttnSmem<D, T>& sm, const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV, int q0, int head,
                                                   int batch, int j0, int n_kv, int cta_rank) {
    using Cfg = AttnCfg<D, T>;
    if (elect_one()) {
        mbar_arrive_expect_tx(sm.q_full(), T * Cfg::kTileBytes);
        for (int t = 0; t < T; ++t)
            for (int h = 0; h < Cfg::kHalves; ++h)
                tma_load_4d(sm.q_smem + t * Cfg::kTileBytes + h * Cfg::kHalfBytes, tmQ, sm.q_full(), h * 64, q0 + t * kBlockM, head, batch);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_kv; ++j) {
        for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
            mbar_wait(sm.kv_empty(stage), phase ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(sm.kv_full(stage), Cfg::kTileBytes);
                const uint32_t dst = sm.kv_smem + stage * Cfg::kTileBytes;
                for (int h = 0; h < Cfg::kHalves; ++h) {
                    if (CL == 2)  // my 64 rows of the tile, into both CTAs
                        tma_load_4d_multicast(dst + h * Cfg::kHalfBytes + cta_rank * (kBlockN / 2) * 128, kv == 0 ? tmK : tmV, sm.kv_full(stage), h * 64,
                                              (j0 + j) * kBlockN + cta_rank * (kBlockN / 2), head, batch, 3);
                    else
                        tma_load_4d(dst + h * Cfg::kHalfBytes, kv == 0 ? tmK : tmV, sm.kv_full(stage), h * 64, (j0 + j) * kBlockN, head, batch);
                }
            }
            __syncwarp();
            if (++stage == Cfg::kKvStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    }
}

#if VAP_ATTN_TRACE
#define TRM(k) do { if (trm && j < 64) trm[j * 8 + (k)] = clock64(); } while (0)
#else
#define TRM(k) do { } while (0)
#endif

template <int D, int CL, bool kP3 = false>
__device__ __forceinline__ void attn_mma_warp(const AttnSmem<D>& sm, uint32_t tmem_base, int n_kv, long long* trm) {
    using Cfg = AttnCfg<D>;
    constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);  // S = Q K^T : A, B K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, D, 0, 1);        // O = P V   : A (TMEM) K-major, B MN-major
    const uint32_t col_s[2] = {tmem_base + Cfg::kColS0, tmem_base + Cfg::kColS1};
    const uint32_t col_o[2] = {tmem_base + Cfg::kColO0, tmem_base + Cfg::kColO1};
    const uint32_t q_smem = sm.q_smem, kv_smem = sm.kv_smem;

    auto issue_qk = [&](int i, uint32_t k_addr) {
        const uint32_t q_addr = q_smem + i * Cfg::kTileBytes;
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                const uint32_t off = (k >> 2) * Cfg::kHalfBytes + (k & 3) * 32;
                umma_ss(col_s[i], make_smem_desc(q_addr + off, 0, 1024, kLayoutSw128),
                        make_smem_desc(k_addr + off, 0, 1024, kLayoutSw128), idesc_qk, k != 0 ? 1u : 0u);
            }
        }
        __syncwarp();
    };
    // K-steps [k0, k1) of O_i += P_i V (16 kv rows each)
    auto issue_pv = [&](int i, int k0, int k1, uint32_t v_addr, uint32_t accumulate) {
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockN / 16; ++k) {
                // A: P_i, packed bf16 pairs, 8 TMEM columns per 16 kv;  B: V rows [16k, 16k+16) (2048 B apart),
                // MN-major: 64-column slabs kHalfBytes apart (LBO), 8-row groups 1024 B apart (SBO)
                if (k >= k0 && k < k1)
                    umma_ts(col_o[i], col_s[i] + 8 * k, make_smem_desc(v_addr + k * 2048, Cfg::kHalfBytes, 1024, kLayoutSw128), idesc_pv,
                            k != 0 ? 1u : accumulate);
            }
        }
        __syncwarp();
    };
    auto issue_pv_half = [&](int i, int c, uint32_t v_addr, uint32_t accumulate) { issue_pv(i, 4 * c, 4 * c + 4, v_addr, accumulate); };
    constexpr int kMid = 4 + VAP_ATTN_P3_PAIRS / 4;  // three-stage publish: stage 1 covers K-steps [4, kMid), the last stage [kMid, 8)
    auto commit = [&](uint32_t bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
    };
    auto release = [&](uint32_t bar) {  // a K / V ring slot: in a cluster the arrive goes to both CTAs' empty barriers
        if (elect_one()) {
            if (CL == 2) umma_commit_multicast(bar, 3);
            else umma_commit(bar);
        }
        __syncwarp();
    };

    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
        if (++stage == Cfg::kKvStages) {
            stage = 0;
            phase ^= 1;
        }
    };
    mbar_wait(sm.q_full(), 0);
    mbar_wait(sm.kv_full(stage), phase);  // K_0
    tc_fence_after();
    issue_qk(0, kv_smem + stage * Cfg::kTileBytes);
    commit(sm.s_full(0));
    issue_qk(1, kv_smem + stage * Cfg::kTileBytes);
    commit(sm.s_full(1));
    release(sm.kv_empty(stage));
    advance();
#if VAP_ATTN_MMA_ORDER == 0
    static_assert(!kP3, "the three-stage publish is implemented for VAP_ATTN_MMA_ORDER=1 only");
    for (int j = 0; j < n_kv; ++j) {
        const uint32_t par = j & 1;
        const int v_stage = stage;
        TRM(0);
        mbar_wait(sm.kv_full(stage), phase);  // V_j
        advance();
        const int k_stage = stage;
        const bool has_next = (j + 1 < n_kv);
        if (has_next) {
            mbar_wait(sm.kv_full(stage), phase);  // K_{j+1}
            advance();
        }
        TRM(1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            VAP_MMA_WAIT(sm.p_full(i, 0), par);
            tc_fence_after();
            TRM(2 + 2 * i);
            issue_pv_half(i, 0, kv_smem + v_stage * Cfg::kTileBytes, j > 0 ? 1u : 0u);
            commit(sm.pv_half(i));
            VAP_MMA_WAIT(sm.p_full(i, 1), par);
            tc_fence_after();
            issue_pv_half(i, 1, kv_smem + v_stage * Cfg::kTileBytes, 1u);
            if (has_next) {
                issue_qk(i, kv_smem + k_stage * Cfg::kTileBytes);
                commit(sm.s_full(i));  // also covers PV_i(j): O_i is quiescent when the softmax sees S_i(j+1)
            } else {
                commit(sm.o_done(i));
            }
            TRM(3 + 2 * i);
        }
        release(sm.kv_empty(v_stage));
        if (has_next) release(sm.kv_empty(k_stage));
    }
#else
    // Issue order with the issuer's own serial latencies (mbarrier round trips, commits — ~100 clk each) taken off the path that decides when
    // the tensor pipe gets its next MMAs.  A clock64 timeline of the order above (profiles/r02_attn_trace_d128.txt) shows the pipe idle for
    // ~650 of every ~2980 clk between the last QK of tile 1 and the first PV of tile 0 of the next step, although that P had been ready for
    // ~700 clk: the issuer was releasing ring slots, waiting (successfully, but one round trip each) for the next V and K, then for P.  Here
    //   * the operands of step j+1 (V_{j+1}, K_{j+2}) are waited for inside step j, in the slack before tile 1's second P half;
    //   * the ring slots of step j are released after the FIRST MMAs of step j+1 have been issued (the commit then covers them as well);
    // so that between "QK_1(j+1) issued" and "PV_0(j+1) half 0 issued" there is one wait — for P itself.
    int v_stage = stage;
    mbar_wait(sm.kv_full(stage), phase);  // V_0
    advance();
    int k_stage = stage;
    if (n_kv > 1) {
        mbar_wait(sm.kv_full(stage), phase);  // K_1
        advance();
    }
    int prev_v = -1, prev_k = -1;
    for (int j = 0; j < n_kv; ++j) {
        const uint32_t par = j & 1;
        const bool has_next = (j + 1 < n_kv);
        int next_v = 0, next_k = 0;
        TRM(0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            VAP_MMA_WAIT(sm.p_full(i, 0), par);
            tc_fence_after();
            TRM(2 + 2 * i);
            issue_pv_half(i, 0, kv_smem + v_stage * Cfg::kTileBytes, j > 0 ? 1u : 0u);
            commit(sm.pv_half(i));
            if (i == 0) {  // the previous step's slots: every MMA that read them was issued before this commit
                if (prev_v >= 0) release(sm.kv_empty(prev_v));
                if (prev_k >= 0) release(sm.kv_empty(prev_k));
            } else if (has_next) {  // next step's operands, while the softmax of tile 1 works on its second half
                next_v = stage;
                mbar_wait(sm.kv_full(stage), phase);  // V_{j+1}
                advance();
                next_k = stage;
                if (j + 2 < n_kv) {
                    mbar_wait(sm.kv_full(stage), phase);  // K_{j+2}
                    advance();
                }
            }
            VAP_MMA_WAIT(sm.p_full(i, 1), par);
            tc_fence_after();
            if constexpr (kP3) {
                issue_pv(i, 4, kMid, kv_smem + v_stage * Cfg::kTileBytes, 1u);
                commit(sm.pv_mid(i));
                VAP_MMA_WAIT(sm.p_last(i), par);
                tc_fence_after();
                issue_pv(i, kMid, 8, kv_smem + v_stage * Cfg::kTileBytes, 1u);
            } else {
                issue_pv_half(i, 1, kv_smem + v_stage * Cfg::kTileBytes, 1u);
            }
            if (has_next) {
                issue_qk(i, kv_smem + k_stage * Cfg::kTileBytes);
                commit(sm.s_full(i));  // also covers PV_i(j): O_i is quiescent when the softmax sees S_i(j+1)
            } else {
                commit(sm.o_done(i));
            }
            TRM(3 + 2 * i);
        }
        prev_v = v_stage, prev_k = has_next ? k_stage : -1;
        v_stage = next_v, k_stage = next_k;
    }
    if (prev_v >= 0) release(sm.kv_empty(prev_v));
    if (prev_k >= 0) release(sm.kv_empty(prev_k));
#endif
}

// Work-item coordinates of a CTA: (batch, head, 256 query rows) and — split-KV: grid z = batch * kv_splits + split — its KV tile range.
struct AttnWork {
    int q0, head, batch, j0, n_kv;
    unsigned split;
    __device__ __forceinline__ explicit AttnWork(const AttnParams& p, int q_tiles = 2) {
        q0 = blockIdx.x * (q_tiles * kBlockM);
        head = blockIdx.y;
        batch = static_cast<int>(blockIdx.z) / p.kv_splits;
        split = blockIdx.z - static_cast<unsigned>(batch * p.kv_splits);
        const int n_kv_all = (p.Lkv + kBlockN - 1) / kBlockN;
        j0 = static_cast<int>(split * static_cast<unsigned>(n_kv_all) / static_cast<unsigned>(p.kv_splits));
        n_kv = static_cast<int>((split + 1u) * static_cast<unsigned>(n_kv_all) / static_cast<unsigned>(p.kv_splits)) - j0;
    }
};

// The softmax + epilogue warps of the 16-lane organisation (warps kFirstSoftmaxWarp ..), shared by the one-CTA kernel and the CTA-pair kernel
// (kRemoteP: P is published on the LEADER CTA's barriers, where the pair's only MMA issuer waits).
template <int D, bool kRemoteP, typename Smem>
__device__ __forceinline__ void attn_softmax_lane16(const Smem& sm, const AttnParams& p, const AttnWork& wk, uint32_t tmem_base, int warp, int lane) {
    using Cfg = typename Smem::Cfg;  // TMEM column map of the kernel this runs in (two Q tiles, or one in the short-KV kernel)
    constexpr bool kP3 = (VAP_ATTN_P3 != 0) && !kRemoteP && Cfg::kQTiles == 2;  // three-stage publish of P (the CTA-pair and short-KV kernels keep two)
    static_assert(VAP_ATTN_P3_PAIRS == 8 || VAP_ATTN_P3_PAIRS == 12, "VAP_ATTN_P3_PAIRS");
    constexpr int kPolyPairs = (D == 128) ? VAP_ATTN_POLY_PAIRS_D128 : VAP_ATTN_POLY_PAIRS_D64;
    const int q0 = wk.q0, head = wk.head, batch = wk.batch, j0 = wk.j0, n_kv = wk.n_kv;
    const unsigned split = wk.split;
    // ===== softmax + epilogue warps =====
    const int sw = warp - kFirstSoftmaxWarp;
    const int i = sw >> 3;         // Q tile
    const int q = warp & 3;        // TMEM lane quarter this warp may touch
    const int hl = (sw >> 2) & 1;  // which 16 lanes of the quarter
    const int cp = lane & 3;       // column phase inside a quad
    const int row0_in_tile = q * 32 + hl * 16 + (lane >> 2);  // this thread's rows: row0 and row0 + 8
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32 + hl * 16) << 16;
    const uint32_t s_col = tmem_base + lane_addr + (i == 0 ? Cfg::kColS0 : Cfg::kColS1);
    const uint32_t o_col = tmem_base + lane_addr + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
    const float c = p.scale_log2;
    const uint64_t c2 = pack_f32x2(c, c);
    const float thr_off = kRescaleThreshold / c;
    float m_used[2] = {-INFINITY, -INFINITY};  // the reference the accumulators of row r are scaled by (lazily updated)
    float thr[2] = {-INFINITY, -INFINITY};     // m_used + 8 / c : a score above it forces a reference update
    uint64_t nmc2[2] = {0ull, 0ull};           // packed (-m_used c, -m_used c)
    uint64_t l2[2] = {0ull, 0ull};             // packed partial row sums of this thread's columns
    long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && hl == 0 && lane == 0) ? p.trace + i * 512 : nullptr;
#if VAP_ATTN_TRACE
#define TR(k) do { if (tr && j < 64) tr[j * 8 + (k)] = clock64(); } while (0)
#else
#define TR(k) do { } while (0)
#endif
    for (int j = 0; j < n_kv; ++j) {
        TR(0);
        // s_full(i) phase j: QK_i(j) is complete, and with it PV_i(j-1) (issued earlier by the same thread), so O_i is
        // quiescent until the first p_full arrive below
#if VAP_ATTN_LEADER_WAIT
        // Only ONE warp of the tile's eight polls the mbarrier; the others block on a named barrier, which costs no issue
        // slots.  (Eight polling warps executed 39 % of the kernel's instructions and competed with the other tile's
        // softmax for the same schedulers, profiles/r01_attn_v5_in_step.json.)
        if ((sw & 7) == 0) VAP_SM_WAIT(sm.s_full(i), j & 1);
        named_bar_sync(1 + i, 256);
#else
        VAP_SM_WAIT(sm.s_full(i), j & 1);
#endif
        tc_fence_after();
        TR(1);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t sr[32];  // sr[4g + e]: columns 64 ch + 8 g + 2 cp + (e & 1) of row0 (e < 2) / row0 + 8 (e >= 2)
            tmem_ld_16x256b_x8(s_col + 64 * ch, sr);
            tmem_ld_wait();
            if (ch == 0) TR(2);
            const int valid = p.Lkv - (j0 + j) * kBlockN - 64 * ch - 2 * cp;  // this thread's column 8 g + e is inside the sequence iff 8 g + e < valid
            if (valid < 58) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (8 * g + (e & 1) >= valid) sr[4 * g + e] = __float_as_uint(-INFINITY);
            }
#define SF(x) __uint_as_float(sr[x])
            // p = 2^(s*c - m*c) is computed SPECULATIVELY against the current reference; the warp votes afterwards on whether
            // any score sat more than 2^8 above it.  Only when the vote fires (always at the
            // very first half-tile, rarely later) the reference is updated and the pass repeated — so the common step has
            // no max -> exp dependency.  Packed FFMA2 for the scale/shift, MUFU.EX2 for most pairs and the FMA-pipe
            // polynomial for kPolyPairs of every 8 pairs (tools/softmax_pipe_probe.cu: MUFU.EX2 costs 8 clk per warp
            // instruction, the polynomial 9 clk per element of FMA pipe: they only pay off side by side).
            uint32_t pk[16];  // pk[2g] = row0, pk[2g+1] = row0 + 8 : packed P column 32 ch + 4 g + cp
            uint64_t ls[2];
            // pairs [E0, E1) of this half (pair e: group g = e / 2, row r = e & 1) against the current reference; ls = their partial row sums
            auto pairs = [&](auto e0c, auto e1c) {
                ls[0] = 0ull, ls[1] = 0ull;
#pragma unroll
                for (int e = decltype(e0c)::value; e < decltype(e1c)::value; ++e) {
                    const int r = e & 1;
                    const uint64_t x2 = fma_f32x2(pack_f32x2(SF(2 * e), SF(2 * e + 1)), c2, nmc2[r]);
                    float x0, x1, p0, p1;
                    unpack_f32x2(x2, x0, x1);
                    if ((e & 7) < kPolyPairs) {
                        ex2_poly_x2(x0, x1, p0, p1);
                    } else {
                        p0 = ex2_approx(x0);
                        p1 = ex2_approx(x1);
                    }
                    ls[r] = add_f32x2(ls[r], pack_f32x2(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
            };
            // Vote: does any score of these pairs sit more than 2^8 above the reference?  The MUFU pairs are judged by
            // their results (a p > 2^8 makes this thread's partial row sum > 2^8; ex2.approx overflows cleanly to +inf), the
            // polynomial pairs by their scores (the exponent-field arithmetic is only valid for x < 128).  A false positive
            // merely refreshes the reference.  The very first half-tile always votes yes (no reference yet).
            auto vote = [&](auto e0c, auto e1c) -> bool {
                float lo0, hi0, lo1, hi1;
                unpack_f32x2(ls[0], lo0, hi0);
                unpack_f32x2(ls[1], lo1, hi1);
                bool need = (lo0 + hi0 > kSumTrigger) || (lo1 + hi1 > kSumTrigger) || (m_used[0] == -INFINITY);
                if constexpr (kPolyPairs > 0) {
                    float pm0 = -INFINITY, pm1 = -INFINITY;
#pragma unroll
                    for (int e = decltype(e0c)::value; e < decltype(e1c)::value; ++e)
                        if ((e & 7) < kPolyPairs) {
                            if (e & 1) pm1 = fmax3(pm1, SF(2 * e), SF(2 * e + 1));
                            else pm0 = fmax3(pm0, SF(2 * e), SF(2 * e + 1));
                        }
                    need = need || (pm0 > thr[0]) || (pm1 > thr[1]);
                }
                return __any_sync(0xffffffffu, need);
            };
            // Reference update (rare): the row maxima over the WHOLE half (a superset is as good), quad-reduced; the warp rescales its 16 rows
            // of O (lazy rescale) — after the PV MMAs already issued on this step's published P have left O (`pv_bar`, 0 = none in flight).
            auto update = [&](uint32_t pv_bar) {
                float mx0 = fmax3(SF(0), SF(1), SF(4)), mx1 = fmax3(SF(2), SF(3), SF(6));
                mx0 = fmax3(mx0, SF(5), SF(8)), mx1 = fmax3(mx1, SF(7), SF(10));
#pragma unroll
                for (int g = 2; g < 7; ++g) {
                    mx0 = fmax3(mx0, SF(4 * g + 1), SF(4 * g + 4));
                    mx1 = fmax3(mx1, SF(4 * g + 3), SF(4 * g + 6));
                }
                mx0 = fmaxf(mx0, SF(29));
                mx1 = fmaxf(mx1, SF(31));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                const float mn0 = fmaxf(m_used[0], mx0), mn1 = fmaxf(m_used[1], mx1);
                if (j > 0 || ch > 0) {
                    if (pv_bar != 0u) {
                        mbar_wait(pv_bar, j & 1);
                        tc_fence_after();
                    }
                    const float f0 = ex2_approx((m_used[0] - mn0) * c), f1 = ex2_approx((m_used[1] - mn1) * c);
                    l2[0] = mul_f32x2(l2[0], pack_f32x2(f0, f0));
                    l2[1] = mul_f32x2(l2[1], pack_f32x2(f1, f1));
#pragma unroll 1
                    for (int g = 0; g < D / 32; ++g) {
                        uint32_t ov[16];
                        tmem_ld_16x256b_x4(o_col + 32 * g, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * ((e & 2) ? f1 : f0));
                        tmem_st_16x256b_x4(o_col + 32 * g, ov);
                    }
                }
                m_used[0] = mn0, m_used[1] = mn1;
                thr[0] = mn0 + thr_off, thr[1] = mn1 + thr_off;
                nmc2[0] = pack_f32x2(-mn0 * c, -mn0 * c), nmc2[1] = pack_f32x2(-mn1 * c, -mn1 * c);
            };
            auto publish = [&](uint32_t bar) {  // everything stored so far (P, and rescaled O columns) is visible to the MMA issuer
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kRemoteP) mbar_arrive_remote(bar, 0);  // the pair's MMA issuer lives in the leader CTA
                    else mbar_arrive(bar);
                }
            };
            using I0 = std::integral_constant<int, 0>;
            using I16 = std::integral_constant<int, 16>;
            using IA = std::integral_constant<int, VAP_ATTN_P3_PAIRS>;
            if (!(kP3 && ch == 1)) {
#pragma unroll 1
                for (int pass = 0;; ++pass) {  // at most two passes: the second one runs against the refreshed reference
                    pairs(I0{}, I16{});
                    if (pass == 1 || !vote(I0{}, I16{})) break;
                    update(ch > 0 ? sm.pv_half(i) : 0u);  // ch 1: the PV MMAs of this step's first half must have left O_i
                }
                if (ch == 0) TR(3);
                l2[0] = add_f32x2(l2[0], ls[0]);
                l2[1] = add_f32x2(l2[1], ls[1]);
                // P is published the moment it is stored: PV of the first half has to be out of the way before the second half ends.  (Tried on a
                // B200 and slower: loading the second half's scores before this store, 1404 -> 1312 TFLOP/s, and publishing half 0 from inside the
                // second half's pass to hide the store latency, -> 1185: every clock P half 0 is late moves its PV into the tail, profiles/r02_attn_ab.json.)
                tmem_st_16x128b_x8(s_col + 32 * ch, pk);
                if (ch == 1) TR(5);
                publish(sm.p_full(i, ch));
            } else if constexpr (kP3) {
                // second half in two stages: the first VAP_ATTN_P3_PAIRS pairs are stored at once; their publish (store wait + fence + arrive) is
                // issued after the LAST pairs' exp work has been issued, which hides the store latency; the last stage is a quarter (or less)
                // of the half, so that the MMAs behind the softmax on the tile's dependency chain are 8 - kMid PV steps + QK^T.
#pragma unroll 1
                for (int pass = 0;; ++pass) {
                    pairs(I0{}, IA{});
                    if (pass == 1 || !vote(I0{}, IA{})) break;
                    update(sm.pv_half(i));
                }
                l2[0] = add_f32x2(l2[0], ls[0]);
                l2[1] = add_f32x2(l2[1], ls[1]);
                tmem_st_16x128b_x4(s_col + 32, pk);
                if (VAP_ATTN_P3_PAIRS == 12) tmem_st_16x128b_x2(s_col + 32 + 16, pk + 8);
#pragma unroll 1
                for (int pass = 0;; ++pass) {
                    pairs(IA{}, I16{});
                    if (pass == 0) publish(sm.p_full(i, 1));
                    if (pass == 1 || !vote(IA{}, I16{})) break;
                    update(sm.pv_mid(i));  // stage 1 is published: its PV MMAs (and with them half 0's) must have left O_i
                }
                l2[0] = add_f32x2(l2[0], ls[0]);
                l2[1] = add_f32x2(l2[1], ls[1]);
                if (VAP_ATTN_P3_PAIRS == 12) tmem_st_16x128b_x2(s_col + 32 + 24, pk + 12);
                else tmem_st_16x128b_x4(s_col + 32 + 16, pk + 8);
                TR(5);
                publish(sm.p_last(i));
            }
#undef SF
        }
        TR(6);
    }
    // ===== epilogue: O / l -> bf16 -> global; the quad of a row writes 16 contiguous bytes per 8-column group =====
    {
        float l[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float lo, hi;
            unpack_f32x2(l2[r], lo, hi);
            l[r] = lo + hi;
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
        }
#if VAP_ATTN_TRACE
        if (tr && i == 0) tr[1024 + 14] = clock64();  // last P published
#endif
        mbar_wait(sm.o_done(i), 0);
        tc_fence_after();
#if VAP_ATTN_TRACE
        if (tr && i == 0) tr[1024 + 15] = clock64();  // O complete
#endif
        const float inv_l[2] = {1.f / l[0], 1.f / l[1]};
        const int row[2] = {q0 + i * kBlockM + row0_in_tile, q0 + i * kBlockM + row0_in_tile + 8};
        // plain mode: one output tensor; peer mode (Ulysses exchange #2 fused into the epilogue): the query rows of rank r are
        // stored straight into rank r's output buffer over NVLink
        __nv_bfloat16* obase[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (p.o_rows_per_peer > 0) {
                const int peer = row[r] / p.o_rows_per_peer;
                obase[r] = (peer < 8 ? p.o_peer[peer] : p.o_peer[0]) + batch * p.o_sb + head * p.o_sh + 2 * cp +
                           static_cast<int64_t>(row[r] - peer * p.o_rows_per_peer) * p.o_sl;
            } else {
                obase[r] = p.o + static_cast<uint64_t>(split) * static_cast<uint64_t>(p.o_split_stride) + batch * p.o_sb + head * p.o_sh + 2 * cp + static_cast<int64_t>(row[r]) * p.o_sl;
            }
        }
#pragma unroll 1
        for (int g4 = 0; g4 < D / 32; ++g4) {
            uint32_t ov[16];
            tmem_ld_16x256b_x4(o_col + 32 * g4, ov);
            tmem_ld_wait();
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (row[r] < p.Lq) {
                    __nv_bfloat16* orow = obase[r] + 32 * g4;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t w = pack_bf16x2(__uint_as_float(ov[4 * g + 2 * r]) * inv_l[r], __uint_as_float(ov[4 * g + 2 * r + 1]) * inv_l[r]);
                        if (p.accumulate) w = add_bf16x2_as_tensors(*reinterpret_cast<const uint32_t*>(orow + 8 * g), w);
                        *reinterpret_cast<uint32_t*>(orow + 8 * g) = w;
                    }
                }
            }
        }
        if (p.lse && cp == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (row[r] < p.Lq)
                    p.lse[static_cast<uint64_t>(split) * static_cast<uint64_t>(p.lse_split_stride) + (static_cast<int64_t>(batch) * p.H + head) * p.Lq + row[r]] = m_used[r] * p.scale + logf(l[r]);
        }
    }
}

template <int D, int CL>
__global__ void __launch_bounds__(kAttnThreads, 1)  // registers are granted per 4 warps: 18 warps cost 20 -> 96 per thread
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = AttnCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const AttnSmem<D> sm((smem_u32(smem_raw) + 1023u) & ~1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const AttnWork wk(p);
    const int q0 = wk.q0, head = wk.head, batch = wk.batch, j0 = wk.j0, n_kv = wk.n_kv;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) sm.init_barriers(CL, kSoftmaxWarps / 2);
    if (warp == 0) {
        tmem_alloc(sm.tmem_ptr_addr(), Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_ptr_addr()));
    const int cta_rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;

    if (warp < kFirstSoftmaxWarp) {
        if (warp == 0) {
            attn_producer_warp<D, CL>(sm, &tmQ, &tmK, &tmV, q0, head, batch, j0, n_kv, cta_rank);
        } else if (warp == 1) {
            attn_mma_warp<D, CL, VAP_ATTN_P3 != 0>(sm, tmem_base, n_kv, (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace + 1024 : nullptr);
        }
    } else {
        attn_softmax_lane16<D, false>(sm, p, wk, tmem_base, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// SHORT-KV kernel: ONE 128-row Q tile per CTA, TWO CTAs per SM.
// The Wan cross-attention (Lkv = 512 text / 257 image tokens, transformer_wan_mot.py:163-179) gives a CTA 3-4 KV tiles of work: with one CTA
// per SM its prologue (TMEM allocation, barrier init, Q + first K load) and its epilogue (O -> global) are fully exposed — the ncu launch list
// shows 324 / 367 us per launch at 20 280 query rows, ~460 TFLOP/s, 2.8 % of a denoise step for 1.1 % of its FLOPs.  Here a CTA owns 128 query
// rows, half of the TMEM (S | O = 256 columns), a Q tile and a TWO-slot K / V ring (96 KB of shared memory at D = 128), ten warps (producer,
// MMA issuer, the eight softmax warps of the 16-lane organisation above — same code), so two CTAs are resident per SM and one's prologue /
// epilogue / load bubbles overlap the other's tile steps.  Ring protocol with two slots: K_j and V_j alternate between them; the issuer waits
// for V_j before the PV MMAs and for K_{j+1} only right before QK^T(j+1), releases V_j's slot behind PV(j) and K_{j+1}'s slot behind QK^T(j+1)
// (a commit covers every earlier MMA), so the producer refills a slot while the softmax of the next step runs.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kShortThreads = (kFirstSoftmaxWarp + kSoftmaxWarps / 2) * 32;  // 10 warps

template <int D>
__device__ __forceinline__ void attn_mma_warp_short(const AttnSmem<D, 1>& sm, uint32_t tmem_base, int n_kv) {
    using Cfg = AttnCfg<D, 1>;
    constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, D, 0, 1);
    const uint32_t col_s = tmem_base + Cfg::kColS0, col_o = tmem_base + Cfg::kColO0;
    const uint32_t q_addr = sm.q_smem, kv_smem = sm.kv_smem;
    auto issue_qk = [&](uint32_t k_addr) {
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                const uint32_t off = (k >> 2) * Cfg::kHalfBytes + (k & 3) * 32;
                umma_ss(col_s, make_smem_desc(q_addr + off, 0, 1024, kLayoutSw128), make_smem_desc(k_addr + off, 0, 1024, kLayoutSw128), idesc_qk, k != 0 ? 1u : 0u);
            }
        }
        __syncwarp();
    };
    auto issue_pv_half = [&](int c, uint32_t v_addr, uint32_t accumulate) {
        if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kBlockN / 32; ++kk) {
                const int k = 4 * c + kk;
                umma_ts(col_o, col_s + 8 * k, make_smem_desc(v_addr + k * 2048, Cfg::kHalfBytes, 1024, kLayoutSw128), idesc_pv, k != 0 ? 1u : accumulate);
            }
        }
        __syncwarp();
    };
    auto commit = [&](uint32_t bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
    };
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
        if (++stage == Cfg::kKvStages) {
            stage = 0;
            phase ^= 1;
        }
    };
    mbar_wait(sm.q_full(), 0);
    mbar_wait(sm.kv_full(stage), phase);  // K_0
    tc_fence_after();
    issue_qk(kv_smem + stage * Cfg::kTileBytes);
    commit(sm.s_full(0));
    commit(sm.kv_empty(stage));
    advance();
    for (int j = 0; j < n_kv; ++j) {
        const uint32_t par = j & 1;
        const bool has_next = (j + 1 < n_kv);
        const int v_stage = stage;
        mbar_wait(sm.kv_full(stage), phase);  // V_j
        advance();
        mbar_wait(sm.p_full(0, 0), par);
        tc_fence_after();
        issue_pv_half(0, kv_smem + v_stage * Cfg::kTileBytes, j > 0 ? 1u : 0u);
        commit(sm.pv_half(0));
        mbar_wait(sm.p_full(0, 1), par);
        tc_fence_after();
        issue_pv_half(1, kv_smem + v_stage * Cfg::kTileBytes, 1u);
        commit(sm.kv_empty(v_stage));  // V_j's slot is free as soon as PV(j) has read it: V_{j+1} gets QK^T(j+1) + half a softmax pass to arrive
        if (has_next) {
            const int k_stage = stage;
            mbar_wait(sm.kv_full(stage), phase);  // K_{j+1}: its slot was released behind QK^T(j), it has had a whole softmax pass to arrive
            advance();
            tc_fence_after();
            issue_qk(kv_smem + k_stage * Cfg::kTileBytes);
            commit(sm.s_full(0));  // also covers PV(j): O is quiescent when the softmax sees S(j+1)
            commit(sm.kv_empty(k_stage));
        } else {
            commit(sm.o_done(0));
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kShortThreads, 2)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = AttnCfg<D, 1>;
    extern __shared__ uint8_t smem_raw[];
    const AttnSmem<D, 1> sm((smem_u32(smem_raw) + 1023u) & ~1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const AttnWork wk(p, 1);
#if VAP_ATTN_TRACE
    // kernel-level stamps of CTA (0,0,0) in the unused columns 6 / 7 of the issuer's rows (tools/attn_short_trace.py): [1024 + 6] entry, [+ 7] set-up
    // done, [+ 14] last P published, [+ 15] O complete (both by the softmax code), [+ 22] O stored, [+ 23] exit
    long long* trk = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == kFirstSoftmaxWarp && lane == 0) ? p.trace + 1024 : nullptr;
    if (trk) trk[6] = clock64();
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) sm.init_barriers(1, kSoftmaxWarps / 2);
    if (warp == 0) {
        tmem_alloc(sm.tmem_ptr_addr(), Cfg::kTmemCols);  // half of the SM's TMEM: the co-resident CTA allocates the other half
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_ptr_addr()));
#if VAP_ATTN_TRACE
    if (trk) trk[7] = clock64();
#endif

    if (warp == 0) {
        attn_producer_warp<D, 1, 1>(sm, &tmQ, &tmK, &tmV, wk.q0, wk.head, wk.batch, wk.j0, wk.n_kv, 0);
    } else if (warp == 1) {
        attn_mma_warp_short<D>(sm, tmem_base, wk.n_kv);
    } else {
        attn_softmax_lane16<D, false>(sm, p, wk, tmem_base, warp, lane);
    }
#if VAP_ATTN_TRACE
    if (trk) trk[22] = clock64();
#endif

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
#if VAP_ATTN_TRACE
    if (trk) trk[23] = clock64();
#endif
}


// ------------------------------------------------------------------------------------------------------------------------
// "row" organisation of the softmax: ONE THREAD PER QUERY ROW, four warps per Q tile, 12 warps per CTA.
//   warp 0: TMEM allocator + TMA producer, warp 1: MMA issuer (both shared with the kernel above), warps 2-3: idle (they only give their
//   registers away), warps 4-7: softmax of Q tile 0, warps 8-11: softmax of Q tile 1.
// Thread t of warp (4 + 4 i + q) owns TMEM lane 32 q + t = query row 128 i + 32 q + t: tcgen05.ld.32x32b hands it whole 32-column
// chunks of its score row, so there is no shuffle, no quad, no vote except the (rare) reference update; its packed-bf16 P row goes back
// with tcgen05.st.32x32b in the layout the TS MMA reads.  Same arithmetic as above: speculative p = 2^(s c - m_used c), partial-sum
// trigger at 2^8, lazy O rescale, P published in two 64-column halves.  Fewer warps (8 instead of 16 softmax warps share the four
// schedulers with nothing but each other) leave issue slots for the FMA-pipe exp2 polynomial (kPoly of every 8 pairs), which is what
// takes load off the MUFU — at D = 128 the MUFU needs 2048 clk per pair of tile steps against 2048 clk of MMAs.
// Register budget: the kernel launches with 168 registers per thread (384 threads); setmaxnreg.dec takes warps 0-3 down to 72 and the
// registers they RELEASE — 128 x (168 - 72) = 12288, the only pool setmaxnreg.inc can draw from (the SM's unallocated remainder does not
// count: a first version asking for 2 x 128 x 40 = 10240 against 9216 released spun forever in USETMAXREG.TRY_ALLOC) — take the eight
// softmax warps up to 216: 256 x (216 - 168) = 12288.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kRowThreads = 384;
constexpr int kRowFirstSoftmaxWarp = 4;
#ifndef VAP_ATTN_ROW_POLY_D128
#define VAP_ATTN_ROW_POLY_D128 1
#endif
#ifndef VAP_ATTN_ROW_POLY_D64
#define VAP_ATTN_ROW_POLY_D64 2
#endif
#ifndef VAP_ATTN_ROW_PREFETCH
#define VAP_ATTN_ROW_PREFETCH 0  // 1: the whole 128-column score row is loaded at the top of a step (128 registers), 0: 64 columns per half
#endif

template <int D, int CL>
__global__ void __launch_bounds__(kRowThreads, 1)
attn_fwd_row_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = AttnCfg<D>;
    constexpr int kPoly = (D == 128) ? VAP_ATTN_ROW_POLY_D128 : VAP_ATTN_ROW_POLY_D64;
    extern __shared__ uint8_t smem_raw[];
    const AttnSmem<D> sm((smem_u32(smem_raw) + 1023u) & ~1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) sm.init_barriers(CL, 4);
    if (warp == 0) {
        tmem_alloc(sm.tmem_ptr_addr(), Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_ptr_addr()));
    const int cta_rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;

    if (warp < kRowFirstSoftmaxWarp) {
        setmaxnreg_dec<72>();
        const AttnWork wk(p);  // recomputed per role: values carried across the setmaxnreg split get spilled
        if (warp == 0) {
            attn_producer_warp<D, CL>(sm, &tmQ, &tmK, &tmV, wk.q0, wk.head, wk.batch, wk.j0, wk.n_kv, cta_rank);
        } else if (warp == 1) {
            attn_mma_warp<D, CL>(sm, tmem_base, wk.n_kv, nullptr);
        }
    } else {
        setmaxnreg_inc<216>();
        const AttnWork wk(p);
        const int i = (warp - kRowFirstSoftmaxWarp) >> 2;  // Q tile
        const int q = warp & 3;                            // TMEM lane quarter this warp may touch
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t s_col = tmem_base + lane_addr + (i == 0 ? Cfg::kColS0 : Cfg::kColS1);
        const uint32_t o_col = tmem_base + lane_addr + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
        const int row = wk.q0 + i * kBlockM + q * 32 + lane;
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        const float thr_off = kRescaleThreshold / c;
        float m_used = -INFINITY;  // the reference this row's accumulators are scaled by (lazily updated)
        float thr = -INFINITY;     // m_used + 8 / c (only the polynomial pairs are judged by their scores)
        uint64_t nmc2 = 0ull;      // packed (-m_used c, -m_used c)
        uint64_t l2 = 0ull;        // packed partial sums of this row

        for (int j = 0; j < wk.n_kv; ++j) {
            VAP_SM_WAIT(sm.s_full(i), j & 1);  // QK_i(j) complete, and with it PV_i(j-1): O_i is quiescent until our first p_full arrive
            tc_fence_after();
            uint32_t sc[VAP_ATTN_ROW_PREFETCH ? 4 : 2][32];
            if (VAP_ATTN_ROW_PREFETCH) {
#pragma unroll
                for (int t = 0; t < 4; ++t) tmem_ld_x32(s_col + 32 * t, sc[t]);
                tmem_ld_wait();
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t(&sa)[32] = sc[VAP_ATTN_ROW_PREFETCH ? 2 * ch : 0];
                uint32_t(&sb)[32] = sc[VAP_ATTN_ROW_PREFETCH ? 2 * ch + 1 : 1];
                if (!VAP_ATTN_ROW_PREFETCH) {
                    tmem_ld_x32(s_col + 64 * ch, sa);
                    tmem_ld_x32(s_col + 64 * ch + 32, sb);
                    tmem_ld_wait();
                }
                const int valid = p.Lkv - (wk.j0 + j) * kBlockN - 64 * ch;  // column e of this half is inside the sequence iff e < valid
                if (valid < 64) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (e >= valid) sa[e] = __float_as_uint(-INFINITY);
                        if (32 + e >= valid) sb[e] = __float_as_uint(-INFINITY);
                    }
                }
#define SF(x) __uint_as_float((x) < 32 ? sa[(x) & 31] : sb[(x) & 31])
                uint32_t pk[32];  // packed P columns 32 ch + e  (kv 64 ch + 2 e, 2 e + 1)
                uint64_t ls[4];
#pragma unroll 1
                for (int pass = 0;; ++pass) {  // at most two passes: the second one runs against the refreshed reference
                    ls[0] = 0ull, ls[1] = 0ull, ls[2] = 0ull, ls[3] = 0ull;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const uint64_t x2 = fma_f32x2(pack_f32x2(SF(2 * e), SF(2 * e + 1)), c2, nmc2);
                        float x0, x1, p0, p1;
                        unpack_f32x2(x2, x0, x1);
                        if ((e & 7) < kPoly) {
                            ex2_poly_x2(x0, x1, p0, p1);
                        } else {
                            p0 = ex2_approx(x0);
                            p1 = ex2_approx(x1);
                        }
                        ls[e & 3] = add_f32x2(ls[e & 3], pack_f32x2(p0, p1));
                        pk[e] = pack_bf16x2(p0, p1);
                    }
                    ls[0] = add_f32x2(add_f32x2(ls[0], ls[1]), add_f32x2(ls[2], ls[3]));
                    float lo, hi;
                    unpack_f32x2(ls[0], lo, hi);
                    // a MUFU result above 2^8 (ex2.approx overflows cleanly to +inf) shows in the partial sum; the polynomial's exponent
                    // arithmetic is only valid for x < 128, so its pairs are judged by their scores; the first half-tile has no reference yet
                    bool need = (lo + hi > kSumTrigger) || (m_used == -INFINITY);
                    if constexpr (kPoly > 0) {
                        float pm = -INFINITY;
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if ((e & 7) < kPoly) pm = fmax3(pm, SF(2 * e), SF(2 * e + 1));
                        need = need || (pm > thr);
                    }
                    if (pass == 1 || !__any_sync(0xffffffffu, need)) break;
                    // ---- reference update: this row's maximum, rescale the warp's 32 rows of O (factor 1 where nothing changed), repeat ----
                    float mx = fmax3(SF(0), SF(1), SF(2));
#pragma unroll
                    for (int e = 3; e < 63; e += 2) mx = fmax3(mx, SF(e), SF(e + 1));
                    mx = fmaxf(mx, SF(63));
                    const float mn = fmaxf(m_used, mx);
                    if (j > 0 || ch > 0) {
                        if (ch > 0) {  // the PV MMAs of this step's first half must have left O_i
                            mbar_wait(sm.pv_half(i), j & 1);
                            tc_fence_after();
                        }
                        const float f = ex2_approx((m_used - mn) * c);
                        l2 = mul_f32x2(l2, pack_f32x2(f, f));
#pragma unroll 1
                        for (int g = 0; g < D / 32; ++g) {
                            uint32_t ov[32];
                            tmem_ld_x32(o_col + 32 * g, ov);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * f);
                            tmem_st_x32(o_col + 32 * g, ov);
                        }
                    }
                    m_used = mn;
                    thr = mn + thr_off;
                    nmc2 = pack_f32x2(-mn * c, -mn * c);
                }
#undef SF
                l2 = add_f32x2(l2, ls[0]);
                tmem_st_x32(s_col + 32 * ch, pk);
                tmem_st_wait();  // covers the rescaled O columns too
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sm.p_full(i, ch));
            }
        }
        // ===== epilogue: O / l -> bf16 -> global, 16-byte stores (a thread writes its row's 32-column chunk = 64 contiguous bytes) =====
        {
            float lo, hi;
            unpack_f32x2(l2, lo, hi);
            const float l = lo + hi;
            mbar_wait(sm.o_done(i), 0);
            tc_fence_after();
            const float inv_l = 1.f / l;
            const bool row_ok = row < p.Lq;
            // plain mode: one output tensor; peer mode (Ulysses exchange #2 fused into the epilogue): the query rows of rank r are
            // stored straight into rank r's output buffer over NVLink; split-KV: partial O of this KV range
            __nv_bfloat16* orow;
            if (p.o_rows_per_peer > 0) {
                const int peer = row / p.o_rows_per_peer;
                orow = (peer < 8 ? p.o_peer[peer] : p.o_peer[0]) + wk.batch * p.o_sb + wk.head * p.o_sh + static_cast<int64_t>(row - peer * p.o_rows_per_peer) * p.o_sl;
            } else {
                orow = p.o + static_cast<uint64_t>(wk.split) * static_cast<uint64_t>(p.o_split_stride) + wk.batch * p.o_sb + wk.head * p.o_sh + static_cast<int64_t>(row) * p.o_sl;
            }
#pragma unroll 1
            for (int g = 0; g < D / 32; ++g) {
                uint32_t ov[32];
                tmem_ld_x32(o_col + 32 * g, ov);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(ov[8 * v + 0]) * inv_l, __uint_as_float(ov[8 * v + 1]) * inv_l);
                        o.y = pack_bf16x2(__uint_as_float(ov[8 * v + 2]) * inv_l, __uint_as_float(ov[8 * v + 3]) * inv_l);
                        o.z = pack_bf16x2(__uint_as_float(ov[8 * v + 4]) * inv_l, __uint_as_float(ov[8 * v + 5]) * inv_l);
                        o.w = pack_bf16x2(__uint_as_float(ov[8 * v + 6]) * inv_l, __uint_as_float(ov[8 * v + 7]) * inv_l);
                        if (p.accumulate) {
                            const uint4 prev = *reinterpret_cast<const uint4*>(orow + 32 * g + 8 * v);
                            o.x = add_bf16x2_as_tensors(prev.x, o.x), o.y = add_bf16x2_as_tensors(prev.y, o.y);
                            o.z = add_bf16x2_as_tensors(prev.z, o.z), o.w = add_bf16x2_as_tensors(prev.w, o.w);
                        }
                        st_v4(orow + 32 * g + 8 * v, o);
                    }
                }
            }
            if (p.lse && row_ok)
                p.lse[static_cast<uint64_t>(wk.split) * static_cast<uint64_t>(p.lse_split_stride) + (static_cast<int64_t>(wk.batch) * p.H + wk.head) * p.Lq + row] =
                    m_used * p.scale + logf(l);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// CTA-PAIR kernel (D = 128): two CTAs of a cluster — the two SMs of a TPC — work on 512 query rows of one head with tcgen05.mma
// cta_group::2 (M = 256).  Pair tile i = (Q tile i of CTA 0, Q tile i of CTA 1); each CTA keeps the rows of ITS Q tiles in its own TMEM with
// the one-CTA layout S0 | S1 | O0 | O1, runs its own 16 softmax warps on them and stores its own rows.  What the pair shares is the B
// operand: CTA r stages kv rows [64 r, 64 r + 64) of every K tile and head-dim columns [64 r, 64 r + 64) of every V tile — HALF the
// shared-memory operand bytes per MMA (A 4 KB + B 2 KB instead of 4 + 4: a 128 x 128 x 16 MMA fed from 8 KB of shared memory runs at
// ~71 clk instead of 64 because it needs the SM's whole 128 B/clk) and half the L2 -> SM traffic, with 16 KB ring stages (eight of them).
// Protocol (the GEMM pair kernel's): kv_full / q_full / p_full live in the LEADER (CTA 0) and collect both CTAs' TMA bytes resp. the P
// arrivals of both CTAs' softmax warps (remote mbarrier.arrive); the leader's single MMA warp issues every MMA and its tcgen05.commit
// multicasts to both CTAs' s_full / pv_half / o_done / kv_empty.
// ------------------------------------------------------------------------------------------------------------------------
struct AttnPairSmem {
    using Cfg = AttnCfg<128>;                          // TMEM column map: the one-CTA layout S0 | S1 | O0 | O1
    static constexpr int kTileBytes = 128 * 128 * 2;  // one Q tile
    static constexpr int kSlabBytes = 128 * 64 * 2;   // Q: 128 rows x 64 columns (one 128-byte swizzle slab)
    static constexpr int kStageBytes = 64 * 128 * 2;  // half a K tile (64 rows x 128 cols = two 8 KB slabs) or half a V tile (128 rows x 64 cols)
    static constexpr int kStages = 8;
    static constexpr int kBarBytes = 512;
    static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kStageBytes + kBarBytes + 1024;
    uint32_t q_smem, kv_smem, bar_base;
    __device__ __forceinline__ explicit AttnPairSmem(uint32_t smem_base)
        : q_smem(smem_base), kv_smem(smem_base + 2 * kTileBytes), bar_base(smem_base + 2 * kTileBytes + kStages * kStageBytes) {}
    __device__ __forceinline__ uint32_t kv_full(int s) const { return bar_base + 8u * s; }                       // leader: both CTAs' halves have landed
    __device__ __forceinline__ uint32_t kv_empty(int s) const { return bar_base + 8u * (kStages + s); }          // per CTA: the pair's MMAs have read the slot
    __device__ __forceinline__ uint32_t q_full() const { return bar_base + 8u * (2 * kStages); }                 // leader
    __device__ __forceinline__ uint32_t s_full(int i) const { return bar_base + 8u * (2 * kStages + 1 + i); }    // per CTA (multicast commit)
    __device__ __forceinline__ uint32_t p_full(int i, int c) const { return bar_base + 8u * (2 * kStages + 3 + 2 * i + c); }  // leader: 16 warps
    __device__ __forceinline__ uint32_t pv_half(int i) const { return bar_base + 8u * (2 * kStages + 7 + i); }   // per CTA
    __device__ __forceinline__ uint32_t o_done(int i) const { return bar_base + 8u * (2 * kStages + 9 + i); }    // per CTA
    __device__ __forceinline__ uint32_t tmem_ptr_addr() const { return bar_base + 8u * (2 * kStages + 11); }
};
static_assert(2 * AttnPairSmem::kStages + 12 <= AttnPairSmem::kBarBytes / 8, "barrier area");

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    constexpr int D = 128;
    using Cfg = AttnCfg<D>;
    using SM = AttnPairSmem;
    extern __shared__ uint8_t smem_raw[];
    const SM sm((smem_u32(smem_raw) + 1023u) & ~1023u);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const AttnWork wk(p);  // q0 = blockIdx.x * 256: this CTA's own 256 query rows
    const int cta_rank = static_cast<int>(cluster_ctarank());
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SM::kStages; ++s) {
            mbar_init(sm.kv_full(s), 1);   // the leader's producer (arrive + expect_tx of both halves); unused in the peer
            mbar_init(sm.kv_empty(s), 1);  // the leader's multicast commit
        }
        mbar_init(sm.q_full(), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(sm.s_full(i), 1);
            mbar_init(sm.p_full(i, 0), kSoftmaxWarps);  // eight softmax warps of EACH CTA; unused in the peer
            mbar_init(sm.p_full(i, 1), kSoftmaxWarps);
            mbar_init(sm.pv_half(i), 1);
            mbar_init(sm.o_done(i), 1);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc_pair(sm.tmem_ptr_addr(), Cfg::kTmemCols);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything crosses the pair
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_ptr_addr()));

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own Q tiles, own half of every K and V tile; all bytes complete on the LEADER's barriers =====
        if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(sm.q_full(), 2 * 2 * SM::kTileBytes);
            for (int t = 0; t < 2; ++t)
                for (int h = 0; h < 2; ++h)
                    tma_load_4d_pair(sm.q_smem + t * SM::kTileBytes + h * SM::kSlabBytes, &tmQ, sm.q_full(), h * 64, wk.q0 + t * kBlockM, wk.head, wk.batch);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < wk.n_kv; ++j) {
            for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
                mbar_wait(sm.kv_empty(stage), phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(sm.kv_full(stage), 2 * SM::kStageBytes);
                    const uint32_t dst = sm.kv_smem + stage * SM::kStageBytes;
                    const int row = (wk.j0 + j) * kBlockN;
                    if (kv == 0) {  // my 64 kv rows, both 64-column slabs (K-major B operand: N split across the pair)
                        tma_load_4d_pair(dst, &tmK, sm.kv_full(stage), 0, row + cta_rank * (kBlockN / 2), wk.head, wk.batch);
                        tma_load_4d_pair(dst + SM::kStageBytes / 2, &tmK, sm.kv_full(stage), 64, row + cta_rank * (kBlockN / 2), wk.head, wk.batch);
                    } else {  // all 128 kv rows of my 64 head-dim columns (MN-major B operand: N = D split across the pair)
                        tma_load_4d_pair(dst, &tmV, sm.kv_full(stage), cta_rank * 64, row, wk.head, wk.batch);
                    }
                }
                __syncwarp();
                if (++stage == SM::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer (leader CTA only), same order as attn_mma_warp =====
            constexpr uint32_t idesc_qk = make_idesc_bf16(2 * kBlockM, kBlockN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(2 * kBlockM, D, 0, 1);
            const uint32_t col_s[2] = {tmem_base + Cfg::kColS0, tmem_base + Cfg::kColS1};
            const uint32_t col_o[2] = {tmem_base + Cfg::kColO0, tmem_base + Cfg::kColO1};
            const uint32_t q_smem = sm.q_smem, kv_smem = sm.kv_smem;
            auto issue_qk = [&](int i, uint32_t k_addr) {
                const uint32_t q_addr = q_smem + i * SM::kTileBytes;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < D / 16; ++k)
                        umma_ss_pair(col_s[i], make_smem_desc(q_addr + (k >> 2) * SM::kSlabBytes + (k & 3) * 32, 0, 1024, kLayoutSw128),
                                     make_smem_desc(k_addr + (k >> 2) * (SM::kStageBytes / 2) + (k & 3) * 32, 0, 1024, kLayoutSw128), idesc_qk, k != 0 ? 1u : 0u);
                }
                __syncwarp();
            };
            auto issue_pv_half = [&](int i, int c, uint32_t v_addr, uint32_t accumulate) {
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < kBlockN / 32; ++kk) {
                        const int k = 4 * c + kk;
                        umma_ts_pair(col_o[i], col_s[i] + 8 * k, make_smem_desc(v_addr + k * 2048, SM::kStageBytes, 1024, kLayoutSw128), idesc_pv, k != 0 ? 1u : accumulate);
                    }
                }
                __syncwarp();
            };
            auto commit = [&](uint32_t bar) {  // arrives on this barrier in BOTH CTAs
                if (elect_one()) umma_commit_pair(bar, 3);
                __syncwarp();
            };
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == SM::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            const int n_kv = wk.n_kv;
            mbar_wait(sm.q_full(), 0);
            mbar_wait(sm.kv_full(stage), phase);  // K_0
            tc_fence_after();
            issue_qk(0, kv_smem + stage * SM::kStageBytes);
            commit(sm.s_full(0));
            issue_qk(1, kv_smem + stage * SM::kStageBytes);
            commit(sm.s_full(1));
            commit(sm.kv_empty(stage));
            advance();
            int v_stage = stage;
            mbar_wait(sm.kv_full(stage), phase);  // V_0
            advance();
            int k_stage = stage;
            if (n_kv > 1) {
                mbar_wait(sm.kv_full(stage), phase);  // K_1
                advance();
            }
            int prev_v = -1, prev_k = -1;
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t par = j & 1;
                const bool has_next = (j + 1 < n_kv);
                int next_v = 0, next_k = 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(sm.p_full(i, 0), par);
                    tc_fence_after();
                    issue_pv_half(i, 0, kv_smem + v_stage * SM::kStageBytes, j > 0 ? 1u : 0u);
                    commit(sm.pv_half(i));
                    if (i == 0) {
                        if (prev_v >= 0) commit(sm.kv_empty(prev_v));
                        if (prev_k >= 0) commit(sm.kv_empty(prev_k));
                    } else if (has_next) {
                        next_v = stage;
                        mbar_wait(sm.kv_full(stage), phase);  // V_{j+1}
                        advance();
                        next_k = stage;
                        if (j + 2 < n_kv) {
                            mbar_wait(sm.kv_full(stage), phase);  // K_{j+2}
                            advance();
                        }
                    }
                    mbar_wait(sm.p_full(i, 1), par);
                    tc_fence_after();
                    issue_pv_half(i, 1, kv_smem + v_stage * SM::kStageBytes, 1u);
                    if (has_next) {
                        issue_qk(i, kv_smem + k_stage * SM::kStageBytes);
                        commit(sm.s_full(i));
                    } else {
                        commit(sm.o_done(i));
                    }
                }
                prev_v = v_stage, prev_k = has_next ? k_stage : -1;
                v_stage = next_v, k_stage = next_k;
            }
            if (prev_v >= 0) commit(sm.kv_empty(prev_v));
            if (prev_k >= 0) commit(sm.kv_empty(prev_k));
        }
    } else {
        attn_softmax_lane16<D, true>(sm, p, wk, tmem_base, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair's MMAs, loads or barrier arrivals may still touch it
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    }
}


static int make_attn_tmap(CUtensorMap* tm, const AttnTensor& t, int B, int H, int L, int D, int box_rows, const char* name) {
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "attention: %s must be 16-byte aligned", name);
    VAP_REQUIRE(t.sl % 8 == 0 && t.sh % 8 == 0 && t.sb % 8 == 0, "attention: %s strides must be multiples of 8 elements", name);
    const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(L), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    // a size-1 dim may come with stride 0 from the caller; TMA wants a positive multiple of 16 bytes
    const uint64_t strides[3] = {static_cast<uint64_t>(t.sl > 0 ? t.sl : D), static_cast<uint64_t>(t.sh > 0 ? t.sh : D),
                                 static_cast<uint64_t>(t.sb > 0 ? t.sb : D)};
    const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
    return make_tmap_bf16(tm, t.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Developer switches, read per call (cheap: two getenv), so tools can A/B variants inside one process:
//   VAP_ATTN_CLUSTER=2   two-CTA clusters sharing the K / V tiles by TMA multicast
//   VAP_ATTN_SOFTMAX     "row" (one thread per query row, 12 warps) or "lane16" (16-lane TMEM shapes, 18 warps)
static int attn_cluster_mode() {
    const char* e = getenv("VAP_ATTN_CLUSTER");
    return e ? atoi(e) : 0;
}
#ifndef VAP_ATTN_DEFAULT_ROW
#define VAP_ATTN_DEFAULT_ROW 0
#endif
#ifndef VAP_ATTN_DEFAULT_PAIR
#define VAP_ATTN_DEFAULT_PAIR 0
#endif
static bool attn_pair_mode() {  // VAP_ATTN_PAIR = 1: the CTA-pair kernel (cta_group::2) for D = 128
    const char* e = getenv("VAP_ATTN_PAIR");
    if (!e || !*e) return VAP_ATTN_DEFAULT_PAIR != 0;
    return e[0] == '1';
}
static bool attn_row_mode() {
    const char* e = getenv("VAP_ATTN_SOFTMAX");
    if (!e || !*e) return VAP_ATTN_DEFAULT_ROW != 0;
    return e[0] == 'r';
}

template <int D, int CL, bool ROW>
static int launch_attn_d(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t stream) {
    using Cfg = AttnCfg<D>;
    static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget");
    static_assert(2 * Cfg::kKvStages + 16 <= Cfg::kBarBytes / 8, "barrier area");
    auto kernel = ROW ? attn_fwd_row_kernel<D, CL> : attn_fwd_kernel<D, CL>;
    constexpr int threads = ROW ? kRowThreads : kAttnThreads;
    static bool opted_in[64] = {};
    if (int rc = smem_opt_in(kernel, Cfg::kSmemBytes, opted_in)) return rc;
    const unsigned q_blocks = static_cast<unsigned>((p.Lq + 2 * kBlockM - 1) / (2 * kBlockM));
    if (CL == 1) {
        const dim3 grid(q_blocks, p.H, p.B * p.kv_splits);
        kernel<<<grid, threads, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, p);
        VAP_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((q_blocks + 1) / 2 * 2, p.H, p.B * p.kv_splits);  // a trailing CTA without query rows still loads and consumes its share of K / V
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    VAP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmQ, tmK, tmV, p));
    return 0;
}

// Short-KV launches (see attn_fwd_short_kernel): chosen when a CTA would see at most kShortMaxKvTiles KV tiles.  VAP_ATTN_SHORT=0 / 1 forces it off / on.
constexpr int kShortMaxKvTiles = 8;  // measured (tools/attn_short_ab.py, profiles/r02_attn_short_ab.json): ahead up to 1024 KV rows, behind from 2048
static int attn_short_mode() {
    const char* e = getenv("VAP_ATTN_SHORT");
    return (e && *e) ? (e[0] == '1' ? 1 : 0) : -1;
}
template <int D>
static int launch_attn_short(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t stream) {
    using Cfg = AttnCfg<D, 1>;
    static_assert(2 * (Cfg::kSmemBytes + 1024) <= 233472, "two CTAs per SM");
    static bool opted_in[64] = {};
    if (int rc = smem_opt_in(attn_fwd_short_kernel<D>, Cfg::kSmemBytes, opted_in)) return rc;
    const dim3 grid(static_cast<unsigned>((p.Lq + kBlockM - 1) / kBlockM), p.H, p.B * p.kv_splits);
    attn_fwd_short_kernel<D><<<grid, kShortThreads, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int launch_attn_pair(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t stream) {
    static_assert(AttnPairSmem::kSmemBytes <= 232448, "shared memory budget");
    static bool opted_in[64] = {};
    if (int rc = smem_opt_in(attn_fwd_pair_kernel, AttnPairSmem::kSmemBytes, opted_in)) return rc;
    const unsigned q_blocks = static_cast<unsigned>((p.Lq + 2 * kBlockM - 1) / (2 * kBlockM));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((q_blocks + 1) / 2 * 2, p.H, p.B * p.kv_splits);  // a trailing CTA without query rows still stages its half of K / V and runs its softmax on zeros
    cfg.blockDim = dim3(kAttnThreads);
    cfg.dynamicSmemBytes = AttnPairSmem::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    VAP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_pair_kernel, tmQ, tmK, tmV, p));
    return 0;
}

template <int D>
static int launch_attn_variant(bool cluster, bool row, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t stream) {
    if (row) return cluster ? launch_attn_d<D, 2, true>(tmQ, tmK, tmV, p, stream) : launch_attn_d<D, 1, true>(tmQ, tmK, tmV, p, stream);
    return cluster ? launch_attn_d<D, 2, false>(tmQ, tmK, tmV, p, stream) : launch_attn_d<D, 1, false>(tmQ, tmK, tmV, p, stream);
}

int launch_attention_fwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, AttnParams p, int D, cudaStream_t stream) {
    VAP_REQUIRE(D == 64 || D == 128, "attention: head_dim=%d must be 64 or 128", D);
    VAP_REQUIRE(p.B > 0 && p.H > 0 && p.Lq >= 0 && p.Lkv > 0, "attention: bad shape B=%d H=%d Lq=%d Lkv=%d", p.B, p.H, p.Lq, p.Lkv);
    if (p.kv_splits < 1) p.kv_splits = 1;
    VAP_REQUIRE(p.H <= 65535 && static_cast<int64_t>(p.B) * p.kv_splits <= 65535, "attention: H and B * kv_splits must be <= 65535");
    if (p.kv_splits > 1) {
        VAP_REQUIRE(p.kv_splits <= 8 && (p.Lkv + kBlockN - 1) / kBlockN >= p.kv_splits, "attention: kv_splits=%d needs at least that many %d-row KV tiles (Lkv=%d)",
                    p.kv_splits, kBlockN, p.Lkv);
        VAP_REQUIRE(p.lse != nullptr && p.o_rows_per_peer == 0, "attention: split-KV writes local partials and needs their log-sum-exp buffer");
    }
    if (p.o_rows_per_peer > 0) {
        const int npeer = (p.Lq + p.o_rows_per_peer - 1) / p.o_rows_per_peer;
        VAP_REQUIRE(npeer <= 8, "attention: at most 8 output peers");
        for (int r = 0; r < npeer; ++r) VAP_REQUIRE(p.o_peer[r] && (reinterpret_cast<uintptr_t>(p.o_peer[r]) & 15) == 0, "attention: bad output peer %d", r);
        VAP_REQUIRE(p.o_sl % 8 == 0 && p.o_sh % 8 == 0 && p.o_sb % 8 == 0, "attention: output strides must be multiples of 8 elements");
    } else {
        VAP_REQUIRE((reinterpret_cast<uintptr_t>(p.o) & 15) == 0 && p.o_sl % 8 == 0 && p.o_sh % 8 == 0 && p.o_sb % 8 == 0,
                    "attention: output must be 16-byte aligned with strides that are multiples of 8 elements");
    }
    VAP_REQUIRE(!p.accumulate || (p.kv_splits == 1 && p.o_rows_per_peer == 0), "attention: accumulate needs the plain output mode (no split-KV, no peers)");
    if (p.Lq == 0) return 0;
    const bool cluster = attn_cluster_mode() == 2 && p.Lq > 2 * kBlockM;
    const bool pair = attn_pair_mode() && D == 128 && p.Lq > 2 * kBlockM;
    CUtensorMap tmQ, tmK, tmV;
    if (pair) {  // CTA r stages 64 kv rows of K (box 64 x 64) and 64 head-dim columns of V (box 64 x 128)
        if (make_attn_tmap(&tmQ, q, p.B, p.H, p.Lq, D, kBlockM, "q")) return -3;
        if (make_attn_tmap(&tmK, k, p.B, p.H, p.Lkv, D, kBlockN / 2, "k")) return -3;
        if (make_attn_tmap(&tmV, v, p.B, p.H, p.Lkv, D, kBlockN, "v")) return -3;
        return launch_attn_pair(tmQ, tmK, tmV, p, stream);
    }
    if (make_attn_tmap(&tmQ, q, p.B, p.H, p.Lq, D, kBlockM, "q")) return -3;
    if (make_attn_tmap(&tmK, k, p.B, p.H, p.Lkv, D, cluster ? kBlockN / 2 : kBlockN, "k")) return -3;
    if (make_attn_tmap(&tmV, v, p.B, p.H, p.Lkv, D, cluster ? kBlockN / 2 : kBlockN, "v")) return -3;
    const bool row = attn_row_mode();
    const int kv_tiles_per_cta = ((p.Lkv + kBlockN - 1) / kBlockN + p.kv_splits - 1) / p.kv_splits;
    const int short_mode = attn_short_mode();
    if (!cluster && !row && (short_mode == 1 || (short_mode == -1 && kv_tiles_per_cta <= kShortMaxKvTiles)))
        return D == 128 ? launch_attn_short<128>(tmQ, tmK, tmV, p, stream) : launch_attn_short<64>(tmQ, tmK, tmV, p, stream);
    return D == 128 ? launch_attn_variant<128>(cluster, row, tmQ, tmK, tmV, p, stream) : launch_attn_variant<64>(cluster, row, tmQ, tmK, tmV, p, stream);
}

// ------------------------------------------------------------------------------------------------------------------------
// split-KV merge: one thread per 8 output channels (16 bytes).  HBM-bound: reads `splits` partial rows + their log-sum-exps,
// writes one row — to the plain output tensor or straight into the owning peer's buffer (Ulysses exchange #2), exactly like the
// attention epilogue.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_combine_kernel(const AttnCombineParams p) {
    const AttnParams& d = p.dst;
    const int vec_per_head = p.D / 8;
    const int64_t vec_per_row = static_cast<int64_t>(d.H) * vec_per_head;
    const int64_t total = static_cast<int64_t>(d.B) * d.Lq * vec_per_row;
    const int64_t part_stride = static_cast<int64_t>(d.B) * d.Lq * d.H * p.D;  // elements between splits of o_part
    const int64_t lse_stride = static_cast<int64_t>(d.B) * d.H * d.Lq;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(idx % vec_per_head);
        const int head = static_cast<int>((idx / vec_per_head) % d.H);
        const int64_t brow = idx / vec_per_row;  // batch * Lq + row
        const int row = static_cast<int>(brow % d.Lq);
        const int batch = static_cast<int>(brow / d.Lq);
        const int64_t lse_idx = (static_cast<int64_t>(batch) * d.H + head) * d.Lq + row;
        float ls[8];
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < p.splits) {
                ls[s] = p.lse_part[s * lse_stride + lse_idx];
                mx = fmaxf(mx, ls[s]);
            }
        float wsum = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < p.splits) {
                ls[s] = __expf(ls[s] - mx);
                wsum += ls[s];
            }
        const float inv = 1.f / wsum;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const __nv_bfloat16* src = p.o_part + (brow * d.H + head) * p.D + 8 * c8;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < p.splits) {
                const uint4 u = ld_nc_v4(src + s * part_stride);
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
                const float ws = ls[s] * inv;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = bf16x2_to_float2(w[j]);
                    acc[2 * j] += ws * f.x;
                    acc[2 * j + 1] += ws * f.y;
                }
            }
        __nv_bfloat16* out;
        if (d.o_rows_per_peer > 0) {
            const int peer = row / d.o_rows_per_peer;
            out = (peer < 8 ? d.o_peer[peer] : d.o_peer[0]) + batch * d.o_sb + head * d.o_sh + static_cast<int64_t>(row - peer * d.o_rows_per_peer) * d.o_sl;
        } else {
            out = d.o + batch * d.o_sb + head * d.o_sh + static_cast<int64_t>(row) * d.o_sl;
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]);
        o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]);
        o.w = pack_bf16x2(acc[6], acc[7]);
        st_v4(out + 8 * c8, o);
        if (d.lse && c8 == 0) d.lse[lse_idx] = mx + __logf(wsum);
    }
}

int launch_attention_combine(const AttnCombineParams& p, cudaStream_t stream) {
    const AttnParams& d = p.dst;
    VAP_REQUIRE(p.D == 64 || p.D == 128, "attention combine: head_dim=%d must be 64 or 128", p.D);
    VAP_REQUIRE(p.splits >= 1 && p.splits <= 8, "attention combine: splits=%d must be in [1, 8]", p.splits);
    VAP_REQUIRE(d.B > 0 && d.H > 0 && d.Lq >= 0, "attention combine: bad shape B=%d H=%d Lq=%d", d.B, d.H, d.Lq);
    VAP_REQUIRE(p.o_part && p.lse_part && (reinterpret_cast<uintptr_t>(p.o_part) & 15) == 0, "attention combine: partials must be 16-byte aligned");
    VAP_REQUIRE(d.o_sl % 8 == 0 && d.o_sh % 8 == 0 && d.o_sb % 8 == 0, "attention combine: output strides must be multiples of 8 elements");
    if (d.o_rows_per_peer > 0) {
        const int npeer = (d.Lq + d.o_rows_per_peer - 1) / d.o_rows_per_peer;
        VAP_REQUIRE(npeer <= 8, "attention combine: at most 8 output peers");
        for (int r = 0; r < npeer; ++r) VAP_REQUIRE(d.o_peer[r] && (reinterpret_cast<uintptr_t>(d.o_peer[r]) & 15) == 0, "attention combine: bad output peer %d", r);
    } else {
        VAP_REQUIRE(d.o && (reinterpret_cast<uintptr_t>(d.o) & 15) == 0, "attention combine: output must be 16-byte aligned");
    }
    const int64_t total = static_cast<int64_t>(d.B) * d.Lq * d.H * (p.D / 8);
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    attn_combine_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vap

// gemm_sm100.cu — bf16 GEMM  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)  on tcgen05 / TMEM / TMA (sm_100a).
//
// The MoT projections (fused QKV, O, FFN-in, FFN-out; reference call sites transformer_wan_mot.py:214-216,
// 241-243, attention.py:1245-1251, attention_processor.py:2923-2925, 2952) are nn.Linear layers:
// activations [M,K] row-major and weights [N,K] row-major, i.e. both operands K-major — exactly the
// layout UMMA wants, so neither side is transposed or repacked.
//
// Two persistent, warp-specialised kernels; a per-shape wave model (gemm_use_pair) picks one:
//   gemm_bf16_pair_kernel : 256 x 256 output tile per CLUSTER OF TWO CTAs on tcgen05.mma cta_group::2 (M = 256 across the SM pair:
//                           each CTA stages its 128 rows of A and half of the W tile) — the large MoT projections
//   gemm_bf16_kernel      : 128 x BN tile per CTA (cta_group::1) — small / narrow problems, described next
// One-CTA kernel, one CTA per SM:
//   warp 0   : TMA producer  (A tile 128x64, W tile BNx64 per stage, SWIZZLE_128B, kStages-deep mbarrier ring)
//   warp 1   : MMA issuer    (one elected thread, tcgen05.mma cta_group::1 kind::f16, M=128 N=BN K=16)
//   warp 2   : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11: epilogue     (tcgen05.ld -> bias / GELU / gated residual with the reference's bf16 rounding
//                             points -> 16-byte global stores), overlapped with the next tile's main loop; two warps per
//                             TMEM lane quarter, each draining half of the row's columns (VAP_GEMM_EPI_WARPS = 8)
#include <cstdlib>
#include <cstring>
#include "vap_kernels.cuh"

namespace vap {

enum GemmEpilogue : int {
    kEpiBias = 0,         // C = bf16(acc + bias)
    kEpiBiasGelu = 1,     // C = bf16(gelu_tanh(bf16(acc + bias)))                          activations.py:65-91
    kEpiGateResF32 = 2,   // C = bf16(float(R) + bf16(acc + bias) * gate[n])   gate fp32      transformer_wan_mot.py:658-663, 684-697
    kEpiResAdd = 3,       // C = bf16(float(R) + bf16(acc + bias))                            transformer_wan_mot.py:675-676
    kEpiGateResBf16 = 4,  // C = bf16(float(R) + bf16(gate[n] * bf16(acc + bias)))  gate bf16  cogvideox_transformer_3d_mot.py:445-446
};


constexpr int kBM = 128;
constexpr int kBK = 64;
// Epilogue warps per CTA: 4 (one per TMEM lane quarter, a thread drains a whole accumulator row) or 8 (two per quarter, each draining half of the
// row's columns): the epilogue of a tile overlaps the main loop of the next one, and at K <= 5120 a four-warp residual / GELU epilogue is the longer of the two.
#ifndef VAP_GEMM_EPI_WARPS
#define VAP_GEMM_EPI_WARPS 8
#endif
constexpr int kEpiWarps = VAP_GEMM_EPI_WARPS;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "VAP_GEMM_EPI_WARPS");
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;
#ifndef VAP_GEMM_GROUP_M
#define VAP_GEMM_GROUP_M 16
#endif
#ifndef VAP_GEMM_PAIR_STAGES
#define VAP_GEMM_PAIR_STAGES 7
#endif
constexpr int kGroupM = VAP_GEMM_GROUP_M;  // M-blocks (pair kernel: 256-row units) that share one sweep over N

// MT = M-tiles (128 rows each) per CTA tile.  MT = 2 (256 x 256 tile, two MMAs per k-step sharing one W tile) streams a third less
// operand data from L2 per FLOP than MT = 1 (128 x 256) but its two accumulators fill all 512 TMEM columns (the epilogue of a
// tile is not overlapped with the next tile's main loop) and only three 64 KB stages fit: measured slower, kept as an experiment.
template <int BN, int MT>
struct GemmCfg {
    static constexpr int kABytes = MT * kBM * kBK * 2;
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (MT == 2) ? 3 : ((BN == 256) ? 4 : 6);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kAccStages = (MT == 2) ? 1 : 2;
    static constexpr int kTmemCols = (MT * BN * kAccStages >= 512) ? 512 : 256;  // powers of two
};

__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int& mb, int& nb) {
    // grouped rasterisation: kGroupM M-blocks share the same sweep over N so A and W tiles are reused from L2
    const int per_group = kGroupM * n_blocks;
    const int g = tile / per_group;
    const int r = tile - g * per_group;
    const int m0 = g * kGroupM;
    const int gm = min(kGroupM, m_blocks - m0);
    mb = m0 + r % gm;
    nb = r / gm;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
    const float kBeta = 0.7978845608028654f;  // sqrt(2/pi)
    const float kKappa = 0.044715f;
    const float inner = kBeta * (x + kKappa * x * x * x);
    return 0.5f * x * (1.f + tanhf(inner));
}

// Epilogue of one accumulator row: thread `lane` of the warp that owns TMEM lane quarter q reads its row (BN fp32 columns at
// `taddr`, 32 at a time), applies bias / GELU / gated residual with the reference's bf16 rounding points and stores 16-byte vectors.
template <int BN, bool kRes>
__device__ __forceinline__ void epilogue_rows_impl(const GemmParams& p, uint32_t taddr, int row, int nb, int c_begin, int c_end) {
    const bool row_ok = row < p.M;
    const float* gate = nullptr;
    if (p.gate) gate = p.gate + (p.rows_per_batch > 0 ? (row_ok ? row / p.rows_per_batch : 0) : 0) * p.gate_stride;
    __nv_bfloat16* crow = p.C + static_cast<int64_t>(row) * p.ldc;
    const __nv_bfloat16* rrow = p.R ? p.R + static_cast<int64_t>(row) * p.ldr : nullptr;
    constexpr bool has_res = kRes;  // two instantiations: the bias / GELU epilogues keep their short loop body (a shared body cost the GELU epilogue 25 % at K = 3072)
    // The residual row is read one 32-column chunk AHEAD of the accumulator (a thread's 64 bytes per chunk are two full sectors of its own row, but
    // 32 different rows per warp instruction: latency-bound).  Read right before use, the eight chunks of a residual epilogue took longer than the
    // main loop of the next tile at K = 5120 (ncu launch list, Wan O-projection 20 280 x 5120 x 5120: 820 us against 687 us for the same GEMM
    // with the bias-only epilogue); one chunk ahead, the loads fly during the tcgen05.ld + arithmetic + stores of the chunk before.
    uint4 rnext[4];
    auto load_res = [&](int c0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int n = nb * BN + c0 + 8 * g;
            if (row_ok && n < p.N) rnext[g] = *reinterpret_cast<const uint4*>(rrow + n);
        }
    };
    if (has_res) load_res(c_begin);
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint4 rcur[4];
        if (has_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
            if (c0 + 32 < c_end) load_res(c0 + 32);
        }
        uint32_t v[32];
        tmem_ld_x32(taddr + c0, v);
        tmem_ld_wait();
        const int n0 = nb * BN + c0;
        if (row_ok && n0 < p.N) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // 4 groups of 8 columns = 16-byte stores
                const int n = n0 + 8 * g;
                if (n < p.N) {
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[8 * g + j]);
                    if (p.bias) {
                        const uint4 bb = *reinterpret_cast<const uint4*>(p.bias + n);
                        const uint32_t* bu = reinterpret_cast<const uint32_t*>(&bb);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = bf16x2_to_float2(bu[j]);
                            y[2 * j] += f.x;
                            y[2 * j + 1] += f.y;
                        }
                    }
                    if (kRes || p.epilogue != kEpiBias) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);
                        if (!kRes) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) y[j] = gelu_tanh_f(y[j]);
                        } else {
                            const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rcur[g]);
                            float r[8];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = bf16x2_to_float2(ru[j]);
                                r[2 * j] = f.x;
                                r[2 * j + 1] = f.y;
                            }
                            if (p.epilogue == kEpiResAdd) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) y[j] = r[j] + y[j];
                            } else {
                                const float4 g0 = *reinterpret_cast<const float4*>(gate + n);
                                const float4 g1 = *reinterpret_cast<const float4*>(gate + n + 4);
                                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                                if (p.epilogue == kEpiGateResF32) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) y[j] = r[j] + y[j] * gg[j];
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) y[j] = r[j] + bf16_round(gg[j] * y[j]);
                                }
                            }
                        }
                    }
                    uint4 o;
                    o.x = pack_bf16x2(y[0], y[1]);
                    o.y = pack_bf16x2(y[2], y[3]);
                    o.z = pack_bf16x2(y[4], y[5]);
                    o.w = pack_bf16x2(y[6], y[7]);
                    *reinterpret_cast<uint4*>(crow + n) = o;
                }
            }
        }
    }
}

// `part` = (warp - 4) / 4: which share of the row's columns this warp drains (kEpiWarps / 4 shares)
template <int BN>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, uint32_t taddr, int row, int nb, int part) {
    constexpr int kCols = BN / (kEpiWarps / 4);
    static_assert(kCols % 32 == 0, "column share of an epilogue warp");
    if (p.epilogue >= kEpiGateResF32) epilogue_rows_impl<BN, true>(p, taddr, row, nb, part * kCols, (part + 1) * kCols);
    else epilogue_rows_impl<BN, false>(p, taddr, row, nb, part * kCols, (part + 1) * kCols);
}

// CL = 2: the kernel runs as clusters of two CTAs that work on vertically adjacent M-blocks of the SAME N-block: each CTA fetches
// half of the W tile and TMA multicasts it into both CTAs' shared memory, so the W operand crosses the L2 -> SM fabric once per
// pair (a third less operand traffic per FLOP, same tile shape / pipeline depth / accumulator double-buffering).  A shared-memory
// slot is reusable when BOTH CTAs' MMAs have read it: every CTA's tcgen05.commit arrives on both CTAs' empty barriers.
template <int BN, int MT, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = GemmCfg<BN, MT>;
    static_assert(CL == 1 || (CL == 2 && MT == 1 && BN == 256), "the cluster variant is the 128 x 256 tile");
    const int cta_rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;
    // persistent loop: work unit u = one N-block x CL M-blocks; this CTA takes M-block CL * mbu + cta_rank of its cluster's units
    const int unit0 = (CL == 2) ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
    const int unit_stride = (CL == 2) ? static_cast<int>(cluster_nctaid_x()) : static_cast<int>(gridDim.x);
    const int m_units = (p.m_blocks + CL - 1) / CL;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
    // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = m_units * p.n_blocks;
    const int k_blocks = (p.K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CL);  // one tcgen05.commit per CTA of the cluster
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), kEpiWarps);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_addr, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                int mb, nb;
                tile_coords(tile, m_units, p.n_blocks, mb, nb);
                mb = mb * CL + cta_rank;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t b_dst = a_dst + Cfg::kABytes;
                    mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    tma_load_2d(a_dst, &tmA, full_bar(stage), kb * kBK, mb * (kBM * MT));
                    if (CL == 2)  // my half of the W tile, into both CTAs
                        tma_load_2d_multicast(b_dst + cta_rank * (BN / 2) * kBK * 2, &tmB, full_bar(stage), kb * kBK, nb * BN + cta_rank * (BN / 2), 3);
                    else
                        tma_load_2d(b_dst, &tmB, full_bar(stage), kb * kBK, nb * BN);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (MT * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t db = make_smem_desc(b_addr + k * 32, 0, 1024, kLayoutSw128);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {  // the M-tiles share the W operand
                            const uint64_t da = make_smem_desc(a_addr + mt * (kBM * kBK * 2) + k * 32, 0, 1024, kLayoutSw128);
                            umma_ss(d_tmem + mt * BN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (CL == 2) umma_commit_multicast(empty_bar(stage), 3);  // frees the slot in both CTAs once these MMAs have read it
                    else umma_commit(empty_bar(stage));
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                if (++acc == Cfg::kAccStages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps =====
        const int q = warp & 3;  // TMEM lane quarter owned by this warp
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
            int mb, nb;
            tile_coords(tile, m_units, p.n_blocks, mb, nb);
            mb = mb * CL + cta_rank;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt)
                epilogue_rows<BN>(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BN) + mt * BN, (mb * MT + mt) * kBM + q * 32 + lane, nb, (warp - 4) >> 2);
            // all TMEM reads of this accumulator stage are complete (tmem_ld_wait) -> release it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == Cfg::kAccStages) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant: ONE 256 x 256 output tile per cluster of two CTAs, computed by tcgen05.mma cta_group::2 (M = 256).
// CTA rank r of the pair stages its own 128 rows of A and HALF of the W tile (rows [128 r, 128 r + 128) of the 256) per k-step:
// 32 KB per stage and CTA instead of 48 KB, so every byte of shared memory feeds twice the MMA work of the one-CTA kernel
// (the operand read rate per FLOP halves for W), and seven stages fit.  Protocol (the CUTLASS 2-SM scheme):
//   full[s]   lives in the LEADER (even CTA): its producer arms it with the bytes of BOTH CTAs' loads; the peer's TMA completes
//             its transaction bytes there (cp.async.bulk.tensor.cta_group::2)
//   empty[s]  per CTA; the leader's tcgen05.commit.cta_group::2 arrives on both (multicast)
//   tfull[a]  per CTA; committed by the leader on both -> each CTA's epilogue warps read their own 128 accumulator rows
//   tempty[a] lives in the leader, 2 x kEpiWarps arrivals: the epilogue warps of BOTH CTAs (the peer's arrive remotely)
// ------------------------------------------------------------------------------------------------------------------------
struct GemmPairCfg {
    static constexpr int BN = 256;
    static constexpr int kABytes = kBM * kBK * 2;        // this CTA's 128 rows of A
    static constexpr int kBBytes = (BN / 2) * kBK * 2;   // this CTA's half of the W tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = VAP_GEMM_PAIR_STAGES;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = 512;  // two accumulator stages x 256 fp32 columns (x 128 lanes in each CTA)
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = GemmPairCfg;
    constexpr int BN = Cfg::BN;
    const int cta_rank = static_cast<int>(cluster_ctarank());
    const bool leader = cta_rank == 0;
    const int unit0 = static_cast<int>(cluster_id_x());
    const int unit_stride = static_cast<int>(cluster_nctaid_x());
    const int m_units = (p.m_blocks + 1) / 2;  // 256-row blocks
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = m_units * p.n_blocks;
    const int k_blocks = (p.K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full_bar(s), 1);   // the leader's producer (arrive + expect_tx); unused in the peer
            mbar_init(empty_bar(s), 1);  // the leader's commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 2 * kEpiWarps);  // the epilogue warps of each CTA; unused in the peer
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_ptr_addr, Cfg::kTmemCols);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything crosses the pair
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs) =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                int mu, nb;
                tile_coords(tile, m_units, p.n_blocks, mu, nb);
                const int row0 = (2 * mu + cta_rank) * kBM;
                const int col0 = nb * BN + cta_rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t b_dst = a_dst + Cfg::kABytes;
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                    tma_load_2d_pair(a_dst, &tmA, full_bar(stage), kb * kBK, row0);
                    tma_load_2d_pair(b_dst, &tmB, full_bar(stage), kb * kBK, col0);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // ===== MMA issuer (leader CTA only) =====
            constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_ss_pair(d_tmem, make_smem_desc(a_addr + k * 32, 0, 1024, kLayoutSw128), make_smem_desc(b_addr + k * 32, 0, 1024, kLayoutSw128), idesc,
                                     (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair(empty_bar(stage), 3);  // frees the slot in both CTAs once these MMAs have read it
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_pair(tfull_bar(acc), 3);  // accumulator complete -> both CTAs' epilogues
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps (both CTAs: each drains its own 128 rows of the accumulator) =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
            int mu, nb;
            tile_coords(tile, m_units, p.n_blocks, mu, nb);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            epilogue_rows<BN>(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, (2 * mu + cta_rank) * kBM + q * 32 + lane, nb, (warp - 4) >> 2);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair's MMAs, loads or barrier arrivals may still touch it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    }
}

template <int BN, int MT>
static int launch_gemm_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, MT>;
    static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget");
    static bool opted_in[1][64] = {};
    if (int rc = smem_opt_in(gemm_bf16_kernel<BN, MT, 1>, Cfg::kSmemBytes, opted_in[0])) return rc;
    p.m_blocks = (p.M + kBM * MT - 1) / (kBM * MT);
    p.n_blocks = (p.N + BN - 1) / BN;
    const int tiles = p.m_blocks * p.n_blocks;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    gemm_bf16_kernel<BN, MT, 1><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// 128 x 256 tiles in clusters of two CTAs sharing the W tile by TMA multicast
static int launch_gemm_cluster(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, cudaStream_t stream) {
    using Cfg = GemmCfg<256, 1>;
    static int max_clusters_dev[64];  // per device: 0 = not probed yet (the smem opt-in and the occupancy are per-device facts)
    int dev = 0;
    VAP_CHECK_CUDA(cudaGetDevice(&dev));
    int& max_clusters = max_clusters_dev[dev & 63];
    if (max_clusters == 0) {
        VAP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<256, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        cudaLaunchConfig_t probe{};
        probe.gridDim = dim3(static_cast<unsigned>(sm_count() / 2 * 2));
        probe.blockDim = dim3(kGemmThreads);
        probe.dynamicSmemBytes = Cfg::kSmemBytes;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
        probe.attrs = &attr, probe.numAttrs = 1;
        int n = 0;
        VAP_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_bf16_kernel<256, 1, 2>, &probe));
        max_clusters = n > 0 ? n : -1;
    }
    VAP_REQUIRE(max_clusters > 0, "gemm_bf16: no 2-CTA cluster fits on this device");
    p.m_blocks = (p.M + kBM - 1) / kBM;
    p.n_blocks = (p.N + 255) / 256;
    const int units = ((p.m_blocks + 1) / 2) * p.n_blocks;
    const int clusters = units < max_clusters ? units : max_clusters;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    VAP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<256, 1, 2>, tmA, tmB, p));
    return 0;
}

// 256 x 256 tiles on CTA pairs (tcgen05.mma cta_group::2)
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, cudaStream_t stream) {
    using Cfg = GemmPairCfg;
    static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget");
    static_assert(2 * Cfg::kStages + 5 <= 32, "barrier area");
    static int max_clusters_dev[64];  // per device: 0 = not probed yet
    int dev = 0;
    VAP_CHECK_CUDA(cudaGetDevice(&dev));
    int& max_clusters = max_clusters_dev[dev & 63];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    if (max_clusters == 0) {
        VAP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / 2 * 2));
        int n = 0;
        VAP_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_bf16_pair_kernel, &cfg));
        max_clusters = n > 0 ? n : -1;
    }
    VAP_REQUIRE(max_clusters > 0, "gemm_bf16: no 2-CTA cluster fits on this device");
    p.m_blocks = (p.M + kBM - 1) / kBM;
    p.n_blocks = (p.N + 255) / 256;
    const int units = ((p.m_blocks + 1) / 2) * p.n_blocks;
    const int clusters = units < max_clusters ? units : max_clusters;
    cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
    VAP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_pair_kernel, tmA, tmB, p));
    return 0;
}

// Which kernel runs a wide (N >= 256) shape.  Measured on the MoT shapes (tools/gemm_pair_ab.sh, B200): the CTA-pair kernel reaches
// 1470-1545 TFLOP/s where the one-CTA kernel reaches 1405-1440 (cuBLAS 1460-1645), but its 256 x 256 units quantise worse on small
// problems (M = 769: 673 vs 848).  So, per shape, a wave model decides: a pair unit (two tiles of work on two SMs) takes ~0.94 of a
// one-CTA tile time.  VAP_GEMM_PAIR = "auto" (default) | 0 (never) | 1 (every eligible shape).
static bool gemm_use_pair(int M, int N) {
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("VAP_GEMM_PAIR");
        mode = (!e || !strcmp(e, "auto")) ? -1 : atoi(e);
    }
    if (M <= kBM || N < 256 || mode == 0) return false;
    if (mode == 1) return true;
    const int sms = sm_count();
    const int n_blocks = (N + 255) / 256;
    const int64_t tiles = static_cast<int64_t>((M + kBM - 1) / kBM) * n_blocks;
    const int64_t units = static_cast<int64_t>((M + 2 * kBM - 1) / (2 * kBM)) * n_blocks;
    const int64_t waves_one = (tiles + sms - 1) / sms;
    const int64_t waves_pair = (units + sms / 2 - 1) / (sms / 2);
    return 0.94 * static_cast<double>(waves_pair) < static_cast<double>(waves_one);
}

static int gemm_cluster_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("VAP_GEMM_CLUSTER");
        mode = e ? atoi(e) : 0;
    }
    return mode;
}

// M-tiles per CTA tile.  Measured on the MoT shapes (tools/gemm_ab.sh): the 256 x 256 tile (MT = 2) loses to 128 x 256 (1330-1350 vs
// 1460-1520 TFLOP/s; only three pipeline stages fit and the epilogue is exposed), so it stays an experiment: VAP_GEMM_MT=2.
static int gemm_m_tiles(int M, int N) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("VAP_GEMM_MT");
        forced = e ? atoi(e) : 0;
    }
    (void)M;
    return (forced == 2 && N >= 256) ? 2 : 1;
}

int launch_gemm_bf16(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, GemmParams p, cudaStream_t stream) {
    VAP_REQUIRE(p.M >= 0 && p.N > 0 && p.K > 0, "gemm_bf16: bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
    VAP_REQUIRE(p.N % 8 == 0 && p.K % 8 == 0, "gemm_bf16: N=%d and K=%d must be multiples of 8", p.N, p.K);
    VAP_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && p.ldc % 8 == 0, "gemm_bf16: leading dimensions must be multiples of 8 elements");
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.C) & 15) == 0,
                "gemm_bf16: A, W and C must be 16-byte aligned");
    VAP_REQUIRE(p.epilogue >= kEpiBias && p.epilogue <= kEpiGateResBf16, "gemm_bf16: unknown epilogue %d", p.epilogue);
    if (p.epilogue >= kEpiGateResF32) {
        VAP_REQUIRE(p.R != nullptr && p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.R) & 15) == 0,
                    "gemm_bf16: residual epilogue needs a 16-byte aligned residual with ldr %% 8 == 0");
        if (p.epilogue != kEpiResAdd) VAP_REQUIRE(p.gate != nullptr, "gemm_bf16: gated epilogue needs a gate vector");
    }
    if (p.bias) VAP_REQUIRE((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "gemm_bf16: bias must be 16-byte aligned");
    if (p.M == 0) return 0;
    const bool wide = p.N >= 256;
    const int BN = wide ? 256 : 128;
    const int MT = gemm_m_tiles(p.M, p.N);
    const bool pair = wide && MT == 1 && gemm_cluster_mode() != 2 && gemm_use_pair(p.M, p.N);
    const bool cluster = !pair && wide && MT == 1 && gemm_cluster_mode() == 2 && p.M > 2 * kBM;
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
        const uint64_t strides[1] = {static_cast<uint64_t>(lda)};
        const uint32_t box[2] = {kBK, static_cast<uint32_t>(kBM * MT)};
        if (make_tmap_bf16(&tmA, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
        const uint64_t strides[1] = {static_cast<uint64_t>(ldw)};
        const uint32_t box[2] = {kBK, static_cast<uint32_t>((cluster || pair) ? BN / 2 : BN)};
        if (make_tmap_bf16(&tmB, W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    }
    if (pair) return launch_gemm_pair(tmA, tmB, p, stream);
    if (cluster) return launch_gemm_cluster(tmA, tmB, p, stream);
    if (MT == 2) return launch_gemm_bn<256, 2>(tmA, tmB, p, stream);
    return wide ? launch_gemm_bn<256, 1>(tmA, tmB, p, stream) : launch_gemm_bn<128, 1>(tmA, tmB, p, stream);
}

}  // namespace vap

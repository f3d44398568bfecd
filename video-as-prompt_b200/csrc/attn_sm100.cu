// attn_sm100.cu — flash-style joint attention forward on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces F.scaled_dot_product_attention(attn_mask=None, dropout_p=0, is_causal=False) at the MoT joint
// attention call sites (transformer_wan_mot.py:637-644, cogvideox_transformer_3d_mot.py:424-431) and the
// per-stream cross-attention (transformer_wan_mot.py:163-179).  q/k/v/o are addressed through explicit
// (batch, head, token) element strides, so the kernel reads Q, K and V straight out of the fused QKV GEMM
// output [tokens, 3, H, D] of BOTH streams (no torch.cat, no head-major transpose) and writes O token-major
// [tokens, H*D], which is the A operand of the output projection.
//
// One CTA = one (batch, head, 256 query rows) work item = two 128-row Q tiles that ping-pong: while the softmax
// warps work on tile i, the tensor core runs PV / QK^T of tile 1-i.  20 warps:
//   warp 0    : TMA producer — Q once, then K_j / V_j tiles into a ring of 128xD bf16 stages (SWIZZLE_128B)
//   warp 1    : MMA issuer   — S_i = Q_i K_j^T (SS, both K-major), O_i += P_i V_j (A = P from TMEM, B = V MN-major)
//   warp 2    : TMEM allocator (S0 | S1 | O0 | O1, fp32 columns; P_i is written as packed bf16 over S_i)
//   warps 4-19: softmax — ALL 16 warps take the same Q tile and alternate between the two tiles.  Warp (q, ch) owns
//               32 query rows (TMEM lane quarter q = warp % 4, one row per thread) x 32 kv columns (chunk ch):
//               tcgen05.ld, partial row max -> shared-memory exchange inside the quarter, online softmax in the exp2
//               domain, lazy rescale of its D/4 O columns (only when the running max grew by > 2^8), P -> TMEM.
// Why 16 warps on one tile: the softmax of a 128x128 tile costs ~1000 clk of MUFU / FMA / ALU pipe time per SM
// sub-partition (tools/softmax_pipe_probe.cu) against 1024 clk of MMA per tile; with warps dedicated to a tile the two
// softmaxes ran concurrently at half speed each and the tensor pipe idled 42 % of the time (profiles/r01).
// The last KV tile is masked against Lkv (TMA zero-fills out-of-range K/V rows).
#include "vap_kernels.cuh"

namespace vap {


constexpr int kSoftmaxWarps = 16;  // four per TMEM lane quarter, one per 32-column chunk of the S tile
constexpr int kAttnThreads = (4 + kSoftmaxWarps) * 32;
constexpr int kBlockM = 128;  // rows per Q tile
constexpr int kBlockN = 128;  // kv rows per tile
constexpr float kRescaleThreshold = 8.0f;
#ifndef VAP_ATTN_POLY_PAIRS
#define VAP_ATTN_POLY_PAIRS 3
#endif
constexpr int kPolyPairs = VAP_ATTN_POLY_PAIRS;  // of every 8 (p0,p1) pairs, how many take the software exp2

template <int D>
struct AttnCfg {
    static constexpr int kTileBytes = 128 * D * 2;  // one Q / K / V tile
    static constexpr int kHalfBytes = 128 * 64 * 2;  // one 64-column (128-byte) swizzle slab
    static constexpr int kHalves = D / 64;
    static constexpr int kKvStages = (D == 128) ? 4 : 8;
    static constexpr int kXchgBytes = 2 * 4 * 128 * 4;  // [Q tile][kv chunk][row] fp32: row-max / row-sum exchange between the warps of a lane quarter
    static constexpr int kSmemBytes = 2 * kTileBytes + kKvStages * kTileBytes + kXchgBytes + 1024 + 256;
    static constexpr int kTmemCols = 512;
    static constexpr int kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 256 + D;
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = AttnCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + 2 * Cfg::kTileBytes;
    const uint32_t xchg_smem = kv_smem + Cfg::kKvStages * Cfg::kTileBytes;
    const uint32_t bar_base = xchg_smem + Cfg::kXchgBytes;
    auto kv_full = [&](int s) { return bar_base + 8u * s; };
    auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::kKvStages + s); };
    const uint32_t q_full = bar_base + 8u * (2 * Cfg::kKvStages);
    auto s_full = [&](int i) { return bar_base + 8u * (2 * Cfg::kKvStages + 1 + i); };
    auto p_full = [&](int i) { return bar_base + 8u * (2 * Cfg::kKvStages + 3 + i); };
    auto o_done = [&](int i) { return bar_base + 8u * (2 * Cfg::kKvStages + 5 + i); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::kKvStages + 7);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * kBlockM);
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int n_kv = (p.Lkv + kBlockN - 1) / kBlockN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kKvStages; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(p_full(i), kSoftmaxWarps);  // one arrive per softmax warp
            mbar_init(o_done(i), 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_addr, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    // 640 threads x 96 registers: a softmax thread holds 32 S values + 16 packed P words, so no setmaxnreg re-balancing.
    // The producer and MMA warps run their loops warp-wide (all lanes wait on the mbarriers, one elected lane issues):
    // control flow and descriptors stay warp-uniform, so ptxas keeps them in uniform registers instead of wrapping
    // every UTCHMMA / UTMALDG in an R2UR waterfall loop.
    if (warp < 4) {
      if (warp == 0) {
            // ===== TMA producer =====
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, 2 * Cfg::kTileBytes);
                for (int t = 0; t < 2; ++t)
                    for (int h = 0; h < Cfg::kHalves; ++h)
                        tma_load_4d(q_smem + t * Cfg::kTileBytes + h * Cfg::kHalfBytes, &tmQ, q_full, h * 64, q0 + t * kBlockM, head, batch);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
                    mbar_wait(kv_empty(stage), phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(kv_full(stage), Cfg::kTileBytes);
                        const uint32_t dst = kv_smem + stage * Cfg::kTileBytes;
                        for (int h = 0; h < Cfg::kHalves; ++h)
                            tma_load_4d(dst + h * Cfg::kHalfBytes, kv == 0 ? &tmK : &tmV, kv_full(stage), h * 64, j * kBlockN, head, batch);
                    }
                    __syncwarp();
                    if (++stage == Cfg::kKvStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
      } else if (warp == 1) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);  // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, D, 0, 1);        // O = P V   : A (TMEM) K-major, B MN-major
            const uint32_t col_s[2] = {tmem_base + Cfg::kColS0, tmem_base + Cfg::kColS1};
            const uint32_t col_o[2] = {tmem_base + Cfg::kColO0, tmem_base + Cfg::kColO1};

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
            auto issue_pv = [&](int i, uint32_t v_addr, uint32_t accumulate) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBlockN / 16; ++k) {
                        // A: P_i, packed bf16 pairs, 8 TMEM columns per 16 kv;  B: V rows [16k, 16k+16) (2048 B apart),
                        // MN-major: 64-column slabs kHalfBytes apart (LBO), 8-row groups 1024 B apart (SBO)
                        umma_ts(col_o[i], col_s[i] + 8 * k, make_smem_desc(v_addr + k * 2048, Cfg::kHalfBytes, 1024, kLayoutSw128),
                                idesc_pv, k != 0 ? 1u : accumulate);
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
            mbar_wait(q_full, 0);
            mbar_wait(kv_full(stage), phase);  // K_0
            tc_fence_after();
            issue_qk(0, kv_smem + stage * Cfg::kTileBytes);
            commit(s_full(0));
            issue_qk(1, kv_smem + stage * Cfg::kTileBytes);
            commit(s_full(1));
            commit(kv_empty(stage));
            advance();
            long long* trm = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace + 1024 : nullptr;
#define TRM(k) do { if (trm && j < 64) trm[j * 8 + (k)] = clock64(); } while (0)
            for (int j = 0; j < n_kv; ++j) {
                const int v_stage = stage;
                TRM(0);
                mbar_wait(kv_full(stage), phase);  // V_j
                advance();
                const int k_stage = stage;
                const bool has_next = (j + 1 < n_kv);
                if (has_next) {
                    mbar_wait(kv_full(stage), phase);  // K_{j+1}
                    advance();
                }
                TRM(1);
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(p_full(i), j & 1);
                    tc_fence_after();
                    TRM(2 + 2 * i);
                    issue_pv(i, kv_smem + v_stage * Cfg::kTileBytes, j > 0 ? 1u : 0u);
                    if (!has_next) commit(o_done(i));  // the final PV; earlier ones are covered by the next s_full commit
                    if (has_next) {
                        issue_qk(i, kv_smem + k_stage * Cfg::kTileBytes);
                        commit(s_full(i));
                    }
                    TRM(3 + 2 * i);
                }
                commit(kv_empty(v_stage));
                if (has_next) commit(kv_empty(k_stage));
            }
      }
    } else {
        // ===== softmax + epilogue warps =====
        // All 16 warps work on ONE Q tile at a time and alternate between the two tiles (unit u = (kv tile j, Q tile i)), so
        // the softmax of tile i overlaps the MMAs of tile 1-i by construction.  Warp (q, ch): q = warp % 4 is the TMEM lane
        // quarter the hardware lets it touch (32 query rows, one per thread), ch = kv column chunk [32 ch, 32 ch + 32).
        // The row max needs all four chunks: the warps of a quarter exchange their partial maxima through shared memory
        // behind a 128-thread named barrier (they sit on the same SM sub-partition).
        const int sw = warp - 4;
        const int q = warp & 3;
        const int ch = sw >> 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t s_col[2] = {tmem_base + lane_addr + Cfg::kColS0, tmem_base + lane_addr + Cfg::kColS1};
        const uint32_t o_col[2] = {tmem_base + lane_addr + Cfg::kColO0, tmem_base + lane_addr + Cfg::kColO1};
        constexpr int kOCols = D / 4;  // O columns this warp rescales / writes out
        const float c = p.scale_log2;
        const uint64_t c2 = pack_f32x2(c, c);
        float m_used[2] = {-INFINITY, -INFINITY};  // the row maximum the accumulators of tile i are scaled by (lazily updated)
        uint64_t l2[2] = {0ull, 0ull};              // packed partial row sums of this thread's 32 columns
        long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && sw == 0 && lane == 0) ? p.trace : nullptr;
#define TR(k) do { if (tr && j < 64) tr[i * 512 + j * 8 + (k)] = clock64(); } while (0)
        for (int j = 0; j < n_kv; ++j) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                TR(0);
                // s_full(i) phase j: QK_i(j) is complete, and with it PV_i(j-1) (issued earlier by the same thread), so O_i is
                // quiescent until the p_full arrive below
                mbar_wait(s_full(i), j & 1);
                tc_fence_after();
                TR(1);
                uint32_t sr[32];
                tmem_ld_x32(s_col[i] + 32 * ch, sr);
                tmem_ld_wait();
                TR(2);
                const int valid = p.Lkv - j * kBlockN - 32 * ch;  // columns of this chunk inside the sequence (may be <= 0)
                if (valid < 32) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (e >= valid) sr[e] = __float_as_uint(-INFINITY);
                }
                float mx0 = fmax3(__uint_as_float(sr[0]), __uint_as_float(sr[1]), __uint_as_float(sr[2]));
                float mx1 = fmax3(__uint_as_float(sr[3]), __uint_as_float(sr[4]), __uint_as_float(sr[5]));
#pragma unroll
                for (int e = 6; e < 30; e += 4) {
                    mx0 = fmax3(mx0, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                    mx1 = fmax3(mx1, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                }
                const float m_loc = fmaxf(fmax3(mx0, __uint_as_float(sr[30]), __uint_as_float(sr[31])), mx1);
                const uint32_t xaddr = xchg_smem + static_cast<uint32_t>(i * 512 + row_in_tile) * 4u;
                st_shared_f32(xaddr + ch * 512u, m_loc);
                named_bar_sync(1 + q, 128);  // also orders every S load of the quarter before any P store below
                const float m_t = fmaxf(fmaxf(ld_shared_f32(xaddr), ld_shared_f32(xaddr + 512u)), fmaxf(ld_shared_f32(xaddr + 1024u), ld_shared_f32(xaddr + 1536u)));
                const float m_new = fmaxf(m_used[i], m_t);
                if (j == 0) {
                    m_used[i] = m_new;
                } else {
                    const bool need = (m_new - m_used[i]) * c > kRescaleThreshold;
                    if (__any_sync(0xffffffffu, need)) {  // identical in the four warps of the quarter (same rows, same maxima)
                        const float f = ex2_approx((m_used[i] - m_new) * c);
                        l2[i] = mul_f32x2(l2[i], pack_f32x2(f, f));
                        uint32_t ov[kOCols];
                        if constexpr (kOCols == 32) tmem_ld_x32(o_col[i] + kOCols * ch, ov);
                        else tmem_ld_x16(o_col[i] + kOCols * ch, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < kOCols; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * f);
                        if constexpr (kOCols == 32) tmem_st_x32(o_col[i] + kOCols * ch, ov);
                        else tmem_st_x16(o_col[i] + kOCols * ch, ov);
                        m_used[i] = m_new;
                    }
                }
                TR(3);
                // p = 2^(s*c - m*c): packed FFMA2 for the scale/shift, MUFU.EX2 for most pairs and the FMA-pipe polynomial for
                // kPolyPairs of every 8 pairs (the MUFU retires one warp-instruction per 8 clk and would otherwise pace the loop)
                const float nmc = -m_used[i] * c;
                const uint64_t nmc2 = pack_f32x2(nmc, nmc);
                uint32_t pk[16];
                uint64_t ls[2] = {0ull, 0ull};
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(sr[2 * e]), __uint_as_float(sr[2 * e + 1])), c2, nmc2);
                    float x0, x1, p0, p1;
                    unpack_f32x2(x2, x0, x1);
                    if ((e & 7) < kPolyPairs) {
                        ex2_poly_x2(x0, x1, p0, p1);
                    } else {
                        p0 = ex2_approx(x0);
                        p1 = ex2_approx(x1);
                    }
                    ls[e & 1] = add_f32x2(ls[e & 1], pack_f32x2(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
                l2[i] = add_f32x2(l2[i], add_f32x2(ls[0], ls[1]));
                tmem_st_x16(s_col[i] + 16 * ch, pk);
                TR(5);
                tmem_st_wait();  // covers the rescaled O columns too
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(i));
                TR(6);
            }
        }
        // ===== epilogue: O / l -> bf16 -> global; warp (q, ch) writes columns [ch D/4, (ch+1) D/4) of its 32 rows =====
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float lo, hi;
            unpack_f32x2(l2[i], lo, hi);
            const uint32_t xaddr = xchg_smem + static_cast<uint32_t>(i * 512 + row_in_tile) * 4u;
            named_bar_sync(1 + q, 128);  // the last row-max exchange of this slot has been read by everyone
            st_shared_f32(xaddr + ch * 512u, lo + hi);
            named_bar_sync(1 + q, 128);
            const float l = (ld_shared_f32(xaddr) + ld_shared_f32(xaddr + 512u)) + (ld_shared_f32(xaddr + 1024u) + ld_shared_f32(xaddr + 1536u));
            mbar_wait(o_done(i), 0);
            tc_fence_after();
            const float inv_l = 1.f / l;
            const int row = q0 + i * kBlockM + row_in_tile;
            uint32_t ov[kOCols];
            if constexpr (kOCols == 32) tmem_ld_x32(o_col[i] + kOCols * ch, ov);
            else tmem_ld_x16(o_col[i] + kOCols * ch, ov);
            tmem_ld_wait();
            if (row < p.Lq) {
                __nv_bfloat16* orow = p.o + batch * p.o_sb + head * p.o_sh + static_cast<int64_t>(row) * p.o_sl + kOCols * ch;
#pragma unroll
                for (int g = 0; g < kOCols / 8; ++g) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(ov[8 * g + 0]) * inv_l, __uint_as_float(ov[8 * g + 1]) * inv_l);
                    o.y = pack_bf16x2(__uint_as_float(ov[8 * g + 2]) * inv_l, __uint_as_float(ov[8 * g + 3]) * inv_l);
                    o.z = pack_bf16x2(__uint_as_float(ov[8 * g + 4]) * inv_l, __uint_as_float(ov[8 * g + 5]) * inv_l);
                    o.w = pack_bf16x2(__uint_as_float(ov[8 * g + 6]) * inv_l, __uint_as_float(ov[8 * g + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + 8 * g) = o;
                }
                if (p.lse && ch == 0) p.lse[(static_cast<int64_t>(batch) * p.H + head) * p.Lq + row] = m_used[i] * p.scale + logf(l);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}


static int make_attn_tmap(CUtensorMap* tm, const AttnTensor& t, int B, int H, int L, int D, const char* name) {
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "attention: %s must be 16-byte aligned", name);
    VAP_REQUIRE(t.sl % 8 == 0 && t.sh % 8 == 0 && t.sb % 8 == 0, "attention: %s strides must be multiples of 8 elements", name);
    const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(L), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    // a size-1 dim may come with stride 0 from the caller; TMA wants a positive multiple of 16 bytes
    const uint64_t strides[3] = {static_cast<uint64_t>(t.sl > 0 ? t.sl : D), static_cast<uint64_t>(t.sh > 0 ? t.sh : D),
                                 static_cast<uint64_t>(t.sb > 0 ? t.sb : D)};
    const uint32_t box[4] = {64, 128, 1, 1};
    return make_tmap_bf16(tm, t.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int D>
static int launch_attn_d(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t stream) {
    using Cfg = AttnCfg<D>;
    static bool attr_set = false;
    if (!attr_set) {
        VAP_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    const dim3 grid((p.Lq + 2 * kBlockM - 1) / (2 * kBlockM), p.H, p.B);
    attn_fwd_kernel<D><<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_attention_fwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, AttnParams p, int D, cudaStream_t stream) {
    VAP_REQUIRE(D == 64 || D == 128, "attention: head_dim=%d must be 64 or 128", D);
    VAP_REQUIRE(p.B > 0 && p.H > 0 && p.Lq >= 0 && p.Lkv > 0, "attention: bad shape B=%d H=%d Lq=%d Lkv=%d", p.B, p.H, p.Lq, p.Lkv);
    VAP_REQUIRE(p.H <= 65535 && p.B <= 65535, "attention: H and B must be <= 65535");
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(p.o) & 15) == 0 && p.o_sl % 8 == 0 && p.o_sh % 8 == 0 && p.o_sb % 8 == 0,
                "attention: output must be 16-byte aligned with strides that are multiples of 8 elements");
    if (p.Lq == 0) return 0;
    CUtensorMap tmQ, tmK, tmV;
    if (make_attn_tmap(&tmQ, q, p.B, p.H, p.Lq, D, "q")) return -3;
    if (make_attn_tmap(&tmK, k, p.B, p.H, p.Lkv, D, "k")) return -3;
    if (make_attn_tmap(&tmV, v, p.B, p.H, p.Lkv, D, "v")) return -3;
    return D == 128 ? launch_attn_d<128>(tmQ, tmK, tmV, p, stream) : launch_attn_d<64>(tmQ, tmK, tmV, p, stream);
}

}  // namespace vap

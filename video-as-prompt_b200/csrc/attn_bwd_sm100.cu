// attn_bwd_sm100.cu — joint attention BACKWARD on tcgen05 / TMEM / TMA (sm_100a): dQ, dK, dV from dO and the forward's
// log-sum-exp.  SURVEY §8(f) rank 4 (trainer integration): the gradient of F.scaled_dot_product_attention at the MoT joint
// attention call sites (transformer_wan_mot.py:637-644, cogvideox_transformer_3d_mot.py:424-431), which the reference gets from
// torch autograd through the SDPA backend selected by finetrainers' dispatcher (finetrainers/models/attention_dispatch.py:416-458).
//
//   P = exp(S * scale - lse),  S = Q K^T            (recomputed; no online softmax: the row statistics are known)
//   dV = P^T dO
//   dP = dO V^T,  dS = P o (dP - delta) * scale,    delta[row] = sum_d dO[row, d] * O[row, d]
//   dQ = dS K,    dK = dS^T Q
//
// Two kernels with ONE skeleton (template flag kDKV), so that every tcgen05 operand form is one the forward kernel already
// uses (SS with both operands K-major; A from TMEM as packed bf16 with B MN-major straight from a [rows, D] tile):
//   dQ  kernel: CTA owns a 128-row Q tile (Q_i, dO_i resident), streams (K_j, V_j):
//        S  = Q_i K_j^T, dP  = dO_i V_j^T  (SS)  ->  threads (one per q row): dS   ->  dQ_i += dS K_j           (TS, K_j MN-major)
//   dKV kernel: CTA owns a 128-row KV tile (K_j, V_j resident), streams (Q_i, dO_i):
//        S^T = K_j Q_i^T, dP^T = V_j dO_i^T (SS)  ->  threads (one per kv row): P^T, dS^T  ->  dV_j += P^T dO_i, dK_j += dS^T Q_i  (TS)
// i.e. S and dP are computed twice (7 instead of 5 tile products per tile pair) in exchange for: no atomics on dQ, no transposed
// operand staging through shared memory, deterministic results.  First version: correctness first — the MMAs of a step and its
// elementwise pass are serialised (one tile pair in flight); overlap (ping-pong as in the forward) is the next step.
//
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 elementwise + epilogue (thread t <-> TMEM lane t <-> tile row t).
// kSplit = 2 (VAP_ATTN_BWD_SPLIT=2, experimental): a second set of four elementwise warps (8-11) takes the upper 64 streamed columns;
// each set packs its bf16 results over ITS OWN already-consumed fp32 columns (set s: columns [64 s, 64 s + 32)), so the sets never touch
// each other's data and the TMEM-operand address of K-step k becomes 64 (k / 4) + 8 (k % 4) instead of 8 k.
// TMEM: [0,128) S / S^T (P^T written over it as packed bf16), [128,256) dP / dP^T (dS / dS^T over it), [256, 256+D) first
// accumulator (dQ or dV), [256+D, 256+2D) second accumulator (dK).
#include <cstdlib>
#include "vap_kernels.cuh"

namespace vap {

constexpr int kBwdTile = 128;
constexpr int bwd_threads(int split) { return 128 + 128 * split; }
constexpr float kLog2e = 1.4426950408889634f;

template <int D>
struct BwdCfg {
    static constexpr int kTileBytes = 128 * D * 2;
    static constexpr int kHalfBytes = 128 * 64 * 2;
    static constexpr int kHalves = D / 64;
    static constexpr int kStages = (D == 128) ? 2 : 3;  // each stage holds the two streamed tiles
    static constexpr int kStatBytes = 2 * 2 * 128 * 4;  // lse2 / delta of the streamed rows, double-buffered (dKV kernel)
    static constexpr int kBarBytes = 256;
    static constexpr int kSmemBytes = 2 * kTileBytes + kStages * 2 * kTileBytes + kStatBytes + kBarBytes + 1024;
    static constexpr int kColS = 0, kColDP = 128, kColAcc0 = 256, kColAcc1 = 256 + D;
};

struct BwdParams {
    int B, H, Lq, Lkv;
    const float* lse;    // [B, H, Lq] natural-log LSE of the scaled scores (the forward's output)
    const float* delta;  // [B, H, Lq] rowsum(dO o O)
    __nv_bfloat16* out0;  // dQ (dQ kernel) or dV (dKV kernel)
    int64_t out0_sb, out0_sh, out0_sl;
    __nv_bfloat16* out1;  // dK (dKV kernel)
    int64_t out1_sb, out1_sh, out1_sl;
    float scale, scale_log2;
    int prefetch_s;  // dQ kernel only (VAP_ATTN_BWD_PREFETCH=1, experimental): S double-buffered in the TMEM columns the dK accumulator would use
};

// tmOwnA / tmOwnB: the CTA's resident tiles (Q_i, dO_i) or (K_j, V_j); tmStrA / tmStrB: the streamed tiles (K_j, V_j) or (Q_i, dO_i).
template <int D, bool kDKV, int kSplit>
__global__ void __launch_bounds__(bwd_threads(kSplit), 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmOwnA, const __grid_constant__ CUtensorMap tmOwnB,
                const __grid_constant__ CUtensorMap tmStrA, const __grid_constant__ CUtensorMap tmStrB, const BwdParams p) {
    using Cfg = BwdCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t own_smem = smem_base;                                   // [A tile | B tile]
    const uint32_t str_smem = smem_base + 2 * Cfg::kTileBytes;             // stages of [A tile | B tile]
    const uint32_t stat_smem = str_smem + Cfg::kStages * 2 * Cfg::kTileBytes;  // [buf][lse2 128 | delta 128] fp32
    const uint32_t bar_base = stat_smem + Cfg::kStatBytes;
    auto st_full = [&](int s) { return bar_base + 8u * s; };
    auto st_empty = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    const uint32_t own_full = bar_base + 8u * (2 * Cfg::kStages);
    const uint32_t sdp_full = bar_base + 8u * (2 * Cfg::kStages + 1);  // S and dP of this step are in TMEM
    const uint32_t ds_full = bar_base + 8u * (2 * Cfg::kStages + 2);   // the bf16 operands of this step are in TMEM
    const uint32_t acc_done = bar_base + 8u * (2 * Cfg::kStages + 3);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int own0 = blockIdx.x * kBwdTile;  // first row of the resident tile (q rows: dQ kernel, kv rows: dKV kernel)
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int own_len = kDKV ? p.Lkv : p.Lq;
    const int str_len = kDKV ? p.Lq : p.Lkv;
    const int n_it = (str_len + kBwdTile - 1) / kBwdTile;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmOwnA);
        tma_prefetch_desc(&tmOwnB);
        tma_prefetch_desc(&tmStrA);
        tma_prefetch_desc(&tmStrB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(st_full(s), 1);
            mbar_init(st_empty(s), 1);
        }
        mbar_init(own_full, 1);
        mbar_init(sdp_full, 1);
        mbar_init(ds_full, 128 * kSplit);  // one arrive per elementwise thread
        mbar_init(acc_done, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_addr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (warp == 0) {
        // ===== TMA producer (warp-wide loop, one elected lane issues) =====
        if (elect_one()) {
            mbar_arrive_expect_tx(own_full, 2 * Cfg::kTileBytes);
            for (int h = 0; h < Cfg::kHalves; ++h) {
                tma_load_4d(own_smem + h * Cfg::kHalfBytes, &tmOwnA, own_full, h * 64, own0, head, batch);
                tma_load_4d(own_smem + Cfg::kTileBytes + h * Cfg::kHalfBytes, &tmOwnB, own_full, h * 64, own0, head, batch);
            }
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < n_it; ++it) {
            mbar_wait(st_empty(stage), phase ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(st_full(stage), 2 * Cfg::kTileBytes);
                const uint32_t dst = str_smem + stage * 2 * Cfg::kTileBytes;
                for (int h = 0; h < Cfg::kHalves; ++h) {
                    tma_load_4d(dst + h * Cfg::kHalfBytes, &tmStrA, st_full(stage), h * 64, it * kBwdTile, head, batch);
                    tma_load_4d(dst + Cfg::kTileBytes + h * Cfg::kHalfBytes, &tmStrB, st_full(stage), h * 64, it * kBwdTile, head, batch);
                }
            }
            __syncwarp();
            if (++stage == Cfg::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_ss = make_idesc_bf16(kBwdTile, kBwdTile, 0, 0);  // [own rows x streamed rows] = A B^T, both K-major (K = D)
        constexpr uint32_t idesc_ts = make_idesc_bf16(kBwdTile, D, 0, 1);         // [own rows x D] += A(TMEM) B, B = [streamed rows, D] MN-major
        const uint32_t col_s = tmem_base + Cfg::kColS, col_dp = tmem_base + Cfg::kColDP;
        const uint32_t col_acc0 = tmem_base + Cfg::kColAcc0, col_acc1 = tmem_base + Cfg::kColAcc1;
        auto issue_ss = [&](uint32_t col, uint32_t a_addr, uint32_t b_addr) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                const uint32_t off = (k >> 2) * Cfg::kHalfBytes + (k & 3) * 32;
                umma_ss(col, make_smem_desc(a_addr + off, 0, 1024, kLayoutSw128), make_smem_desc(b_addr + off, 0, 1024, kLayoutSw128), idesc_ss,
                        k != 0 ? 1u : 0u);
            }
        };
        auto issue_ts = [&](uint32_t col_acc, uint32_t col_a, uint32_t b_addr, uint32_t accumulate) {
#pragma unroll
            for (int k = 0; k < kBwdTile / 16; ++k) {  // A: packed bf16 pairs, 8 TMEM columns per 16 streamed rows; B rows [16k, 16k+16) are 2048 B apart
                constexpr int kStepsPerSet = (kBwdTile / 16) / kSplit;  // elementwise set s packed its columns at [128 / kSplit * s, ...)
                const uint32_t a_col = col_a + (kBwdTile / kSplit) * (k / kStepsPerSet) + 8 * (k % kStepsPerSet);
                umma_ts(col_acc, a_col, make_smem_desc(b_addr + k * 2048, Cfg::kHalfBytes, 1024, kLayoutSw128), idesc_ts, k != 0 ? 1u : accumulate);
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        // dQ kernel with prefetch: S(it + 1) = Q K_{it+1}^T is issued right after dP(it), into the other S buffer, so that it runs on the
        // tensor pipe while the threads work on step it (dP cannot follow: its columns hold dS(it) until dQ += dS K has read them)
        const bool pre = !kDKV && p.prefetch_s != 0;
        auto s_buf = [&](int i) { return col_s + ((pre && (i & 1)) ? 384u : 0u); };
        mbar_wait(own_full, 0);
        for (int it = 0; it < n_it; ++it) {
            mbar_wait(st_full(stage), phase);
            tc_fence_after();
            const uint32_t str_a = str_smem + stage * 2 * Cfg::kTileBytes, str_b = str_a + Cfg::kTileBytes;
            if (elect_one()) {
                if (!pre || it == 0) issue_ss(s_buf(it), own_smem, str_a);  // S = Q K^T      | S^T  = K Q^T
                issue_ss(col_dp, own_smem + Cfg::kTileBytes, str_b);        // dP = dO V^T    | dP^T = V dO^T
                umma_commit(sdp_full);
            }
            __syncwarp();
            if (pre && it + 1 < n_it) {
                const int nstage = (stage + 1 == Cfg::kStages) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == Cfg::kStages) ? (phase ^ 1) : phase;
                mbar_wait(st_full(nstage), nphase);  // K_{it+1} has landed (the loop waits on it again next iteration: already complete)
                tc_fence_after();
                if (elect_one()) issue_ss(s_buf(it + 1), own_smem, str_smem + nstage * 2 * Cfg::kTileBytes);
                __syncwarp();
            }
            mbar_wait(ds_full, it & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t acc = it > 0 ? 1u : 0u;
                if (kDKV) {
                    issue_ts(col_acc0, col_s, str_b, acc);   // dV += P^T dO
                    issue_ts(col_acc1, col_dp, str_a, acc);  // dK += dS^T Q
                } else {
                    issue_ts(col_acc0, col_dp, str_a, acc);  // dQ += dS K
                }
                umma_commit(st_empty(stage));
                if (it + 1 == n_it) umma_commit(acc_done);
            }
            __syncwarp();
            if (++stage == Cfg::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===== elementwise + epilogue: thread t <-> TMEM lane t <-> row own0 + t of the resident tile =====
        const int q4 = warp & 3;
        const int set = (warp - 4) >> 2;                  // which quarter / half of the streamed columns this warp handles
        constexpr int kColsPerSet = kBwdTile / kSplit;
        const int col_base = set * kColsPerSet;
        const int t = q4 * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
        const uint32_t s_col = tmem_base + lane_addr + Cfg::kColS;
        const uint32_t dp_col = tmem_base + lane_addr + Cfg::kColDP;
        const int own_row = own0 + t;
        const bool own_ok = own_row < own_len;
        const float c = p.scale_log2;
        const int64_t stat_base = (static_cast<int64_t>(batch) * p.H + head) * p.Lq;
        float row_lse2 = 0.f, row_delta = 0.f;  // dQ kernel: statistics of this thread's q row
        if (!kDKV && own_ok) {
            row_lse2 = p.lse[stat_base + own_row] * kLog2e;
            row_delta = p.delta[stat_base + own_row];
        }
        for (int it = 0; it < n_it; ++it) {
            const uint32_t stat = stat_smem + (it & 1) * (2 * 128 * 4);
            if (kDKV) {  // statistics of the streamed q rows: thread t fetches row t's, everybody reads all 128 (rows past Lq: 0, 0)
                if (set == 0) {
                    const int qrow = it * kBwdTile + t;
                    const bool ok = qrow < p.Lq;
                    st_shared_f32(stat + 4u * t, ok ? p.lse[stat_base + qrow] * kLog2e : 0.f);
                    st_shared_f32(stat + 512u + 4u * t, ok ? p.delta[stat_base + qrow] : 0.f);
                }
                named_bar_sync(1, 128 * kSplit);
            }
            mbar_wait(sdp_full, it & 1);
            tc_fence_after();
            const int valid = str_len - it * kBwdTile;  // streamed rows (TMEM columns) inside the sequence
#pragma unroll 1
            for (int ch = 0; ch < 4 / kSplit; ++ch) {
                const int c0 = col_base + 32 * ch;  // first of this chunk's 32 fp32 columns
                uint32_t sr[32], dr[32];
                tmem_ld_x32(s_col + ((!kDKV && p.prefetch_s != 0 && (it & 1)) ? 384u : 0u) + c0, sr);
                tmem_ld_x32(dp_col + c0, dr);
                tmem_ld_wait();
                uint32_t pk_p[16], pk_ds[16];
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {  // four columns per step
                    float lse2[4] = {row_lse2, row_lse2, row_lse2, row_lse2}, delta[4] = {row_delta, row_delta, row_delta, row_delta};
                    if (kDKV) {  // per-column statistics of the streamed q rows: every lane reads the same 16 bytes (broadcast)
                        const float4 l4 = ld_shared_v4_f32(stat + 4u * (c0 + 4 * e4));
                        const float4 d4 = ld_shared_v4_f32(stat + 512u + 4u * (c0 + 4 * e4));
                        lse2[0] = l4.x, lse2[1] = l4.y, lse2[2] = l4.z, lse2[3] = l4.w;
                        delta[0] = d4.x, delta[1] = d4.y, delta[2] = d4.z, delta[3] = d4.w;
                    }
                    float pv[4], dsv[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const int col = c0 + 4 * e4 + h;
                        float pe = ex2_approx(fmaf(__uint_as_float(sr[4 * e4 + h]), c, -lse2[h]));
                        // outside the sequences: streamed rows past the end (zero-filled by TMA: S = 0 would give p = exp(-lse)) and rows
                        // of the resident tile past its end contribute nothing
                        if (col >= valid || !own_ok) pe = 0.f;
                        pv[h] = pe;
                        dsv[h] = pe * (__uint_as_float(dr[4 * e4 + h]) - delta[h]) * p.scale;
                    }
                    pk_p[2 * e4] = pack_bf16x2(pv[0], pv[1]), pk_p[2 * e4 + 1] = pack_bf16x2(pv[2], pv[3]);
                    pk_ds[2 * e4] = pack_bf16x2(dsv[0], dsv[1]), pk_ds[2 * e4 + 1] = pack_bf16x2(dsv[2], dsv[3]);
                }
                // packed bf16 over the fp32 columns this thread has already consumed: chunk ch read columns [base + 32 ch, base + 32 ch + 32)
                // and writes [base + 16 ch, base + 16 ch + 16)
                if (kDKV) tmem_st_x16(s_col + col_base + 16 * ch, pk_p);
                tmem_st_x16(dp_col + col_base + 16 * ch, pk_ds);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(ds_full);
        }
        // ===== epilogue: accumulators -> bf16 -> global =====
        mbar_wait(acc_done, 0);
        tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < (kDKV ? 2 : 1); ++a) {
            __nv_bfloat16* base = (a == 0 ? p.out0 : p.out1);
            const int64_t sb = a == 0 ? p.out0_sb : p.out1_sb, sh = a == 0 ? p.out0_sh : p.out1_sh, sl = a == 0 ? p.out0_sl : p.out1_sl;
            __nv_bfloat16* orow = base + batch * sb + head * sh + static_cast<int64_t>(own_row) * sl;
            const uint32_t acc_col = tmem_base + lane_addr + (a == 0 ? Cfg::kColAcc0 : Cfg::kColAcc1);
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                if (((a * (D / 32) + c0 / 32) % kSplit) != set) continue;  // the sets share the epilogue's 32-column chunks
                uint32_t ov[32];
                tmem_ld_x32(acc_col + c0, ov);
                tmem_ld_wait();
                if (own_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(ov[8 * g + 0]), __uint_as_float(ov[8 * g + 1]));
                        o.y = pack_bf16x2(__uint_as_float(ov[8 * g + 2]), __uint_as_float(ov[8 * g + 3]));
                        o.z = pack_bf16x2(__uint_as_float(ov[8 * g + 4]), __uint_as_float(ov[8 * g + 5]));
                        o.w = pack_bf16x2(__uint_as_float(ov[8 * g + 6]), __uint_as_float(ov[8 * g + 7]));
                        *reinterpret_cast<uint4*>(orow + c0 + 8 * g) = o;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// delta[b, h, row] = sum_d dO[b, h, row, d] * O[b, h, row, d]  (fp32).  D / 8 lanes per row, 16-byte loads, sub-warp shuffle reduction.
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                                                             int B, int H, int Lq, int64_t o_sb, int64_t o_sh, int64_t o_sl, int64_t d_sb, int64_t d_sh,
                                                             int64_t d_sl) {
    constexpr int kLanes = D / 8;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t rows = static_cast<int64_t>(B) * H * Lq;
    const int64_t row_id = idx / kLanes;
    const int sub = static_cast<int>(idx % kLanes);
    float acc = 0.f;
    if (row_id < rows) {
        const int l = static_cast<int>(row_id % Lq);
        const int h = static_cast<int>((row_id / Lq) % H);
        const int b = static_cast<int>(row_id / (static_cast<int64_t>(Lq) * H));
        const uint4 ou = ld_nc_v4(o + b * o_sb + h * o_sh + static_cast<int64_t>(l) * o_sl + 8 * sub);
        const uint4 du = ld_nc_v4(dout + b * d_sb + h * d_sh + static_cast<int64_t>(l) * d_sl + 8 * sub);
        const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ou);
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(&du);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = bf16x2_to_float2(ow[j]), g = bf16x2_to_float2(dw[j]);
            acc = fmaf(a.x, g.x, acc);
            acc = fmaf(a.y, g.y, acc);
        }
    }
#pragma unroll
    for (int m = 1; m < kLanes; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (row_id < rows && sub == 0) delta[row_id] = acc;
}

static int make_bwd_tmap(CUtensorMap* tm, const AttnTensor& t, int B, int H, int L, int D, const char* name) {
    VAP_REQUIRE(t.ptr && (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "attention bwd: %s must be a 16-byte aligned device pointer", name);
    VAP_REQUIRE(t.sl % 8 == 0 && t.sh % 8 == 0 && t.sb % 8 == 0, "attention bwd: %s strides must be multiples of 8 elements", name);
    const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(L), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(t.sl > 0 ? t.sl : D), static_cast<uint64_t>(t.sh > 0 ? t.sh : D),
                                 static_cast<uint64_t>(t.sb > 0 ? t.sb : D)};
    const uint32_t box[4] = {64, kBwdTile, 1, 1};
    return make_tmap_bf16(tm, t.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Developer switches, read per call.  Defaults = the fastest configuration measured on a B200 (H 40, J 16 384, D 128, tools/kernel_bench.py --bwd,
// profiles/r02_pending_and_reference_checks.json): one elementwise set 628 TFLOP/s, two sets 661, two sets + S prefetch in the dQ kernel 678; all
// three are parity-green (tests/gpu_checks.py attn_bwd_*).  VAP_ATTN_BWD_SPLIT=1 / VAP_ATTN_BWD_PREFETCH=0 select the simpler variants.
static int bwd_prefetch_mode() {
    const char* e = getenv("VAP_ATTN_BWD_PREFETCH");
    return (e && atoi(e) == 0) ? 0 : 1;
}

static int bwd_split_mode() {
    const char* e = getenv("VAP_ATTN_BWD_SPLIT");
    return (e && atoi(e) == 1) ? 1 : 2;
}

template <int D, int kSplit>
static int launch_bwd_d(const AttnBwdArgs& a, cudaStream_t stream) {
    using Cfg = BwdCfg<D>;
    static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget");
    static_assert(8 * (2 * Cfg::kStages + 5) <= Cfg::kBarBytes, "barrier area");
    static bool opted_in[2][64] = {};
    if (int rc = smem_opt_in(attn_bwd_kernel<D, false, kSplit>, Cfg::kSmemBytes, opted_in[0])) return rc;
    if (int rc = smem_opt_in(attn_bwd_kernel<D, true, kSplit>, Cfg::kSmemBytes, opted_in[1])) return rc;
    // 1. delta = rowsum(dO o O)
    {
        const int64_t threads = static_cast<int64_t>(a.B) * a.H * a.Lq * (D / 8);
        const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
        attn_bwd_delta_kernel<D><<<blocks, 256, 0, stream>>>(a.o.ptr, a.dout.ptr, a.delta, a.B, a.H, a.Lq, a.o.sb, a.o.sh, a.o.sl, a.dout.sb, a.dout.sh,
                                                             a.dout.sl);
        VAP_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap tmQ, tmK, tmV, tmDO;
    if (make_bwd_tmap(&tmQ, a.q, a.B, a.H, a.Lq, D, "q")) return -3;
    if (make_bwd_tmap(&tmK, a.k, a.B, a.H, a.Lkv, D, "k")) return -3;
    if (make_bwd_tmap(&tmV, a.v, a.B, a.H, a.Lkv, D, "v")) return -3;
    if (make_bwd_tmap(&tmDO, a.dout, a.B, a.H, a.Lq, D, "dout")) return -3;
    BwdParams p{};
    p.B = a.B, p.H = a.H, p.Lq = a.Lq, p.Lkv = a.Lkv;
    p.lse = a.lse, p.delta = a.delta;
    p.scale = a.scale, p.scale_log2 = a.scale * kLog2e;
    p.prefetch_s = bwd_prefetch_mode();
    // 2. dQ: one CTA per 128 q rows, streams K / V
    p.out0 = a.dq.ptr, p.out0_sb = a.dq.sb, p.out0_sh = a.dq.sh, p.out0_sl = a.dq.sl;
    {
        const dim3 grid((a.Lq + kBwdTile - 1) / kBwdTile, a.H, a.B);
        attn_bwd_kernel<D, false, kSplit><<<grid, bwd_threads(kSplit), Cfg::kSmemBytes, stream>>>(tmQ, tmDO, tmK, tmV, p);
        VAP_CHECK_CUDA(cudaGetLastError());
    }
    // 3. dV, dK: one CTA per 128 kv rows, streams Q / dO
    p.out0 = a.dv.ptr, p.out0_sb = a.dv.sb, p.out0_sh = a.dv.sh, p.out0_sl = a.dv.sl;
    p.out1 = a.dk.ptr, p.out1_sb = a.dk.sb, p.out1_sh = a.dk.sh, p.out1_sl = a.dk.sl;
    {
        const dim3 grid((a.Lkv + kBwdTile - 1) / kBwdTile, a.H, a.B);
        attn_bwd_kernel<D, true, kSplit><<<grid, bwd_threads(kSplit), Cfg::kSmemBytes, stream>>>(tmK, tmV, tmQ, tmDO, p);
        VAP_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

int launch_attention_bwd(const AttnBwdArgs& a, int D, cudaStream_t stream) {
    VAP_REQUIRE(D == 64 || D == 128, "attention bwd: head_dim=%d must be 64 or 128", D);
    VAP_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lkv > 0, "attention bwd: bad shape B=%d H=%d Lq=%d Lkv=%d", a.B, a.H, a.Lq, a.Lkv);
    VAP_REQUIRE(a.H <= 65535 && a.B <= 65535, "attention bwd: H and B must be <= 65535");
    VAP_REQUIRE(a.lse && a.delta, "attention bwd: lse and the delta workspace are required");
    VAP_REQUIRE(a.o.ptr && (reinterpret_cast<uintptr_t>(a.o.ptr) & 15) == 0 && a.o.sl % 8 == 0 && a.o.sh % 8 == 0 && a.o.sb % 8 == 0,
                "attention bwd: o must be 16-byte aligned with strides that are multiples of 8 elements");
    const AttnGrad* outs[3] = {&a.dq, &a.dk, &a.dv};
    for (const AttnGrad* g : outs)
        VAP_REQUIRE(g->ptr && (reinterpret_cast<uintptr_t>(g->ptr) & 15) == 0 && g->sl % 8 == 0 && g->sh % 8 == 0 && g->sb % 8 == 0,
                    "attention bwd: gradients must be 16-byte aligned with strides that are multiples of 8 elements");
    if (bwd_split_mode() == 2) return D == 128 ? launch_bwd_d<128, 2>(a, stream) : launch_bwd_d<64, 2>(a, stream);
    return D == 128 ? launch_bwd_d<128, 1>(a, stream) : launch_bwd_d<64, 1>(a, stream);
}

}  // namespace vap

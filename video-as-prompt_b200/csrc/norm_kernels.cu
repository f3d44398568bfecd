// norm_kernels.cu — HBM-bound kernels of the MoT block (sm_100a).
//
//  * adaln_layernorm_kernel : (adaLN-modulated) LayerNorm, one or four warps per token row, the row lives in
//    registers (16-byte vector loads, read once / written once: 4 B per element of algorithmic traffic).
//      Wan : (LN_fp32(x) [*w+b]) * (1+scale) + shift -> bf16            transformer_wan_mot.py:620-623, 668-669, 680-689
//      Cog : bf16(bf16(bf16(LN_affine(x)) * bf16(1+scale)) + shift)     normalization.py:464-471
//  * qk_norm_rope_kernel    : q/k normalisation + temporally-biased RoPE, in place on the joint q|k|v buffer.
//      Wan : RMSNorm across all H*D channels (fp32 variance, bf16 rounding before the weight multiply,
//            normalization.py:554-568) + complex RoPE on interleaved pairs (transformer_wan_mot.py:229-236)
//      Cog : per-head LayerNorm(D, eps 1e-6, affine) (attention_processor.py:2934-2937) + real cos/sin RoPE
//            on video tokens only (attention_processor.py:2943-2945, embeddings.py:1229-1248)
#include "vap_kernels.cuh"

namespace vap {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// (adaLN) LayerNorm
// ------------------------------------------------------------------------------------------

// WPR warps share one token row (WPR = 1 for d <= 2048, 4 above): a thread then holds at most MAXV = 8 16-byte vectors, so
// the row stays in registers WITHOUT spilling (the one-warp-per-row version held 20 vectors per lane at d = 5120, spilled
// and reached 1.3 TB/s) and 1024 threads per SM keep ~80 KB of loads in flight.  Row statistics: two passes over the
// registers, warp shuffles, and for WPR = 4 a 4-float exchange through shared memory behind a 128-thread named barrier.
template <int MAXV, int WPR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (MAXV <= 5) ? 4 : 3) adaln_layernorm_kernel(LnParams p) {
    __shared__ float red[2][kWarpsPerBlock];
    const int warp = threadIdx.x >> 5;
    const int tl = threadIdx.x & (32 * WPR - 1);  // thread index inside the row group
    const int group = warp / WPR;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (kWarpsPerBlock / WPR) + group;
    if (row >= p.rows) return;  // whole row groups leave together (the named barrier below is per group)
    const int nvec = p.d >> 3;  // 8 bf16 per 16-byte vector
    const __nv_bfloat16* xr = p.x + row * p.x_stride;
    constexpr int kStride = 32 * WPR;
    const int lane = tl;  // vector index base

    uint4 raw[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + kStride * i;
        if (v < nvec) raw[i] = ld_nc_v4(xr + 8 * v);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + kStride * i < nvec) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = bf16x2_to_float2(u[j]);
                sum += f.x + f.y;
            }
        }
    }
    sum = warp_sum(sum);
    if constexpr (WPR > 1) {
        if ((threadIdx.x & 31) == 0) red[0][warp] = sum;
        named_bar_sync(1 + group, 32 * WPR);
        sum = 0.f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) sum += red[0][group * WPR + w];
    }
    const float mean = sum / static_cast<float>(p.d);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + kStride * i < nvec) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = bf16x2_to_float2(u[j]);
                const float a = f.x - mean, b = f.y - mean;
                sq += a * a + b * b;
            }
        }
    }
    sq = warp_sum(sq);
    if constexpr (WPR > 1) {
        if ((threadIdx.x & 31) == 0) red[1][warp] = sq;
        named_bar_sync(1 + group, 32 * WPR);
        sq = 0.f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) sq += red[1][group * WPR + w];
    }
    const float rstd = rsqrtf(sq / static_cast<float>(p.d) + p.eps);

    const int64_t batch = p.rows_per_batch > 0 ? row / p.rows_per_batch : 0;
    const float* s1p = p.scale1p ? p.scale1p + batch * p.mod_stride : nullptr;
    const float* shf = p.shift ? p.shift + batch * p.mod_stride : nullptr;
    __nv_bfloat16* orow = p.out + row * p.out_stride;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + kStride * i;
        if (v < nvec) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
            float y[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = bf16x2_to_float2(u[j]);
                y[2 * j] = (f.x - mean) * rstd;
                y[2 * j + 1] = (f.y - mean) * rstd;
            }
            const int c = 8 * v;
            if (p.ln_w) {
                const float4 w0 = *reinterpret_cast<const float4*>(p.ln_w + c), w1 = *reinterpret_cast<const float4*>(p.ln_w + c + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(p.ln_b + c), b1 = *reinterpret_cast<const float4*>(p.ln_b + c + 4);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = y[j] * w[j] + b[j];
            }
            if (p.cog_rounding) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);
            }
            if (s1p) {
                const float4 s0 = *reinterpret_cast<const float4*>(s1p + c), s1 = *reinterpret_cast<const float4*>(s1p + c + 4);
                const float4 h0 = *reinterpret_cast<const float4*>(shf + c), h1 = *reinterpret_cast<const float4*>(shf + c + 4);
                const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                if (p.cog_rounding) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j] * s[j]) + h[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = y[j] * s[j] + h[j];
                }
            }
            uint4 o;
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            o.z = pack_bf16x2(y[4], y[5]);
            o.w = pack_bf16x2(y[6], y[7]);
            st_v4(orow + c, o);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Staged variants for wide rows (d = 256 NV): persistent CTAs, one per SM, eight warps; every warp owns a private ring of kStages
// row buffers in shared memory that it fills itself with cp.async.bulk (one bulk copy per token row, completion on the warp's own
// mbarriers) — so the loads of row k+1 are in flight while row k is being reduced, normalised and stored, with no dependence between
// warps and ~80 KB of reads outstanding per SM all the time.  The per-channel fp32 vectors (modulation / affine / norm weights) are
// staged in shared memory ONCE per CTA instead of being re-read through L1 for every row.  The row itself is pulled from shared
// memory into registers once (NV 16-byte vectors per lane) and the mean / variance / normalise passes run on the registers.
// The CTAs of a launch never straddle a batch: CTA (b, c) works on the rows of batch b only, so its staged modulation vectors hold.
// ------------------------------------------------------------------------------------------
constexpr int kStagedWarps = 8;

__device__ __forceinline__ void bulk_load_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4_u32(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

struct StagedRows {  // the contiguous row range [r0, r1) of one batch this CTA works on
    int64_t r0, r1, batch;
    __device__ __forceinline__ StagedRows(int64_t rows, int64_t rows_per_batch, int ctas_per_batch) {
        batch = blockIdx.x / ctas_per_batch;
        const int c = blockIdx.x - static_cast<int>(batch) * ctas_per_batch;
        const int64_t b0 = batch * rows_per_batch;
        const int64_t nb = (rows - b0 < rows_per_batch) ? rows - b0 : rows_per_batch;
        const int64_t chunk = (nb + ctas_per_batch - 1) / ctas_per_batch;
        r0 = b0 + c * chunk;
        r1 = r0 + chunk < b0 + nb ? r0 + chunk : b0 + nb;
    }
};

template <int NV, int kStages>
__global__ void __launch_bounds__(kStagedWarps * 32, 1) adaln_layernorm_staged_kernel(LnParams p, int ctas_per_batch) {
    extern __shared__ __align__(128) uint8_t ln_smem[];
    constexpr int d = 256 * NV;
    constexpr uint32_t row_bytes = d * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const StagedRows rg(p.rows, p.rows_per_batch, ctas_per_batch);
    // layout: [fp32 vectors that exist: (scale1p | shift) (ln_w | ln_b)] [ring: warp x stage x row] [mbarriers]
    const uint32_t n_vec = (p.scale1p ? 2u : 0u) + (p.ln_w ? 2u : 0u);
    float* vec_mod = reinterpret_cast<float*>(ln_smem);                 // scale1p, shift
    float* vec_aff = vec_mod + (p.scale1p ? 2 * d : 0);                 // ln_w, ln_b
    const uint32_t ring = smem_u32(ln_smem) + n_vec * d * sizeof(float) + static_cast<uint32_t>(warp) * kStages * row_bytes;
    const uint32_t bars = smem_u32(ln_smem) + n_vec * d * sizeof(float) + kStagedWarps * kStages * row_bytes + static_cast<uint32_t>(warp) * kStages * 8u;
    const float* srcs[4] = {p.scale1p ? p.scale1p + rg.batch * p.mod_stride : nullptr, p.shift ? p.shift + rg.batch * p.mod_stride : nullptr, p.ln_w, p.ln_b};
    float* dsts[4] = {vec_mod, vec_mod + d, vec_aff, vec_aff + d};
    // Shared-memory layout of a staged vector: the 8 channels of 16-byte data vector v are split into two float4 halves, lo[v] at float
    // offset 4 v and hi[v] at d / 2 + 4 v, so that the 32 lanes of a warp read 32 CONSECUTIVE float4 (a lane-strided 32-byte read would hit
    // every bank twice: the vectors are 4x the bytes of the data and shared-memory bandwidth is what bounds this kernel after HBM).
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (srcs[t])
            for (int f = threadIdx.x; f < d / 4; f += kStagedWarps * 32)
                *reinterpret_cast<float4*>(dsts[t] + (f & 1) * (d / 2) + (f >> 1) * 4) = *reinterpret_cast<const float4*>(srcs[t] + 4 * f);
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + 8u * s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t first = rg.r0 + warp;
    auto issue = [&](int64_t k) {  // row first + 8 k into stage k % kStages
        const int64_t r = first + kStagedWarps * k;
        if (r < rg.r1 && lane == 0) {
            const uint32_t st = static_cast<uint32_t>(k % kStages);
            mbar_arrive_expect_tx(bars + 8u * st, row_bytes);
            bulk_load_row(ring + st * row_bytes, p.x + r * p.x_stride, row_bytes, bars + 8u * st);
        }
    };
#pragma unroll
    for (int k = 0; k < kStages; ++k) issue(k);
    const float inv_d = 1.f / static_cast<float>(d);
    for (int64_t k = 0;; ++k) {
        const int64_t row = first + kStagedWarps * k;
        if (row >= rg.r1) break;
        const uint32_t st = static_cast<uint32_t>(k % kStages);
        mbar_wait(bars + 8u * st, static_cast<uint32_t>(k / kStages) & 1u);
        uint4 raw[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) raw[i] = ld_shared_v4_u32(ring + st * row_bytes + (lane + 32 * i) * 16);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(u[j]);
                sum += f.x + f.y;
            }
        }
        // The row now lives in registers (the sum above consumed every vector, so the shared-memory reads have completed): the stage is
        // refilled at once with the row kStages ahead, which keeps kStages rows per warp in flight while this one is being processed.
        __syncwarp();
        issue(k + kStages);
        const float mean = warp_sum(sum) * inv_d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(u[j]);
                const float a = f.x - mean, b = f.y - mean;
                sq += a * a + b * b;
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_d + p.eps);
        __nv_bfloat16* orow = p.out + row * p.out_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 8 * (lane + 32 * i);
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
            float y[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(u[j]);
                y[2 * j] = (f.x - mean) * rstd;
                y[2 * j + 1] = (f.y - mean) * rstd;
            }
            if (p.ln_w) {
                const float4 w0 = *reinterpret_cast<const float4*>(vec_aff + c / 2), w1 = *reinterpret_cast<const float4*>(vec_aff + d / 2 + c / 2);
                const float4 b0 = *reinterpret_cast<const float4*>(vec_aff + d + c / 2), b1 = *reinterpret_cast<const float4*>(vec_aff + d + d / 2 + c / 2);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = y[j] * w[j] + b[j];
            }
            if (p.cog_rounding) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);
            }
            if (p.scale1p) {
                const float4 s0 = *reinterpret_cast<const float4*>(vec_mod + c / 2), s1 = *reinterpret_cast<const float4*>(vec_mod + d / 2 + c / 2);
                const float4 h0 = *reinterpret_cast<const float4*>(vec_mod + d + c / 2), h1 = *reinterpret_cast<const float4*>(vec_mod + d + d / 2 + c / 2);
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                if (p.cog_rounding) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j] * sc[j]) + h[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = y[j] * sc[j] + h[j];
                }
            }
            uint4 o;
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            o.z = pack_bf16x2(y[4], y[5]);
            o.w = pack_bf16x2(y[6], y[7]);
            st_v4(orow + c, o);
        }
    }
}

// Launch plan of a staged kernel: batches x CTAs per batch (about one CTA per SM in total), or 0 CTAs when the shape does not qualify.
static int staged_ctas_per_batch(int64_t rows, int64_t rows_per_batch, int64_t* nbatch) {
    const int64_t rpb = rows_per_batch > 0 && rows_per_batch < rows ? rows_per_batch : rows;
    *nbatch = (rows + rpb - 1) / rpb;
    if (*nbatch > 64 || rows < 16 * kStagedWarps * 8) return 0;  // few rows: the per-CTA staging of the vectors would not amortise
    int per = static_cast<int>(sm_count() / *nbatch);
    if (per < 1) per = 1;
    // developer switch, read per call (tools/norm_rows_ab.py): minimum rows per warp before another CTA is worth staging the vectors again
    const char* env = getenv("VAP_NORM_STAGED_ROWS_PER_WARP");
    // 2: measured on a B200 (profiles/r02_norm_rows_ab.json): at the 2 535 rows of one rank of 8-way Ulysses the q/k-norm + RoPE takes 50.6 us with
    // 2 rows per warp against 65.5 with 4 and 104 with 8; from 5 070 rows up the choice makes no difference
    const int rows_per_warp = (env && atoi(env) > 0) ? atoi(env) : 2;
    const int64_t max_useful = (rpb + kStagedWarps * rows_per_warp - 1) / (kStagedWarps * rows_per_warp);
    if (per > max_useful) per = static_cast<int>(max_useful);
    return per;
}

template <int NV, int kStages>
static int launch_ln_staged(const LnParams& p, int ctas_per_batch, int64_t nbatch, cudaStream_t stream) {
    const int n_vec = (p.scale1p ? 2 : 0) + (p.ln_w ? 2 : 0);
    const int smem = n_vec * 256 * NV * 4 + kStagedWarps * kStages * 256 * NV * 2 + kStagedWarps * kStages * 8;
    if (smem > 232448) return 1;  // does not fit (e.g. d = 5120 with affine AND modulation): the caller takes the register kernel
    static bool opted_in[64] = {};
    if (int rc = smem_opt_in(adaln_layernorm_staged_kernel<NV, kStages>, 232448, opted_in)) return rc;
    adaln_layernorm_staged_kernel<NV, kStages><<<static_cast<unsigned>(nbatch * ctas_per_batch), kStagedWarps * 32, smem, stream>>>(p, ctas_per_batch);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static bool staged_disabled() {
    const char* e = getenv("VAP_NORM_STAGED");
    return e && e[0] == '0';
}

int launch_adaln_layernorm(const LnParams& p, cudaStream_t stream) {
    VAP_REQUIRE(p.d % 8 == 0 && p.d >= 8 && p.d <= 8192, "adaln_layernorm: d=%d must be a multiple of 8 and <= 8192", p.d);
    VAP_REQUIRE(p.x_stride % 8 == 0 && p.out_stride % 8 == 0, "adaln_layernorm: row strides must be multiples of 8 elements");
    VAP_REQUIRE((p.ln_w == nullptr) == (p.ln_b == nullptr), "adaln_layernorm: ln_w and ln_b must both be given or both null");
    VAP_REQUIRE((p.scale1p == nullptr) == (p.shift == nullptr), "adaln_layernorm: scale1p and shift must both be given or both null");
    if (p.rows == 0) return 0;
    const int nvec = p.d / 8;
    if ((p.d == 5120 || p.d == 4096 || p.d == 3072) && !staged_disabled() && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) {
        int64_t nbatch = 1;
        LnParams q = p;
        if (!p.scale1p || p.mod_stride == 0 || p.rows_per_batch <= 0) q.rows_per_batch = p.rows;  // one set of vectors for all rows
        const int per = staged_ctas_per_batch(q.rows, q.rows_per_batch, &nbatch);
        if (per > 0) {
            const int rc = p.d == 5120 ? launch_ln_staged<20, 2>(q, per, nbatch, stream)
                         : p.d == 4096 ? launch_ln_staged<16, 2>(q, per, nbatch, stream) : launch_ln_staged<12, 3>(q, per, nbatch, stream);
            if (rc <= 0) return rc;
        }
    }
    const dim3 block(kWarpsPerBlock * 32);
    if (nvec <= 32 * 8) {  // d <= 2048: one warp per row
        const unsigned grid = static_cast<unsigned>((p.rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
        if (nvec <= 32 * 2)
            adaln_layernorm_kernel<2, 1><<<grid, block, 0, stream>>>(p);
        else
            adaln_layernorm_kernel<8, 1><<<grid, block, 0, stream>>>(p);
    } else {  // four warps per row
        constexpr int kRowsPerBlock = kWarpsPerBlock / 4;
        const unsigned grid = static_cast<unsigned>((p.rows + kRowsPerBlock - 1) / kRowsPerBlock);
        if (nvec <= 128 * 3)
            adaln_layernorm_kernel<3, 4><<<grid, block, 0, stream>>>(p);
        else if (nvec <= 128 * 5)
            adaln_layernorm_kernel<5, 4><<<grid, block, 0, stream>>>(p);
        else
            adaln_layernorm_kernel<8, 4><<<grid, block, 0, stream>>>(p);
    }
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// q/k norm + RoPE (in place)
// ------------------------------------------------------------------------------------------

// One work item = (token row, q or k); WPR warps share an item (WPR = 1 for d <= 2048, 4 above) so a thread holds at most
// 8 16-byte vectors and nothing spills (the one-warp-per-row version held 20 per lane at d = 5120 and did q then k serially).
// Every thread always owns the same 8 channels of a head (vector v = tl + 32*WPR*i -> channel-in-head 8*(v % (D/8)) =
// 8*(tl % (D/8)) because 32*WPR % (D/8) == 0), so its 4 RoPE (cos,sin) pairs and (Cog) its per-head LN affine values are
// loaded once; a head's D/8 vectors sit in consecutive lanes of ONE warp (D/8 divides 32), so the per-head statistics of
// the CogVideoX mode stay shuffle-only.
template <int MAXV, int WPR, bool COG>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (MAXV <= 5) ? 4 : 3) qk_norm_rope_kernel(QkParams p) {
    __shared__ float red[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5;
    const int tl = threadIdx.x & (32 * WPR - 1);
    const int group = warp / WPR;
    constexpr int kStride = 32 * WPR;
    const bool scatter = p.nsplit > 0;
    const int n_which = scatter ? 3 : (p.k ? 2 : 1);  // k == null: normalise q only (cross-attention queries)
    const int64_t item = static_cast<int64_t>(blockIdx.x) * (kWarpsPerBlock / WPR) + group;
    if (item >= p.rows * n_which) return;  // whole groups leave together (the named barrier below is per group)
    const int64_t row = item / n_which;
    const int which = static_cast<int>(item - row * n_which);
    const int d = p.heads * p.head_dim;
    const int nvec = d >> 3;
    const int vec_per_head = p.head_dim >> 3;
    const int ch = 8 * (tl % vec_per_head);  // channel offset inside the head

    const int64_t pos = p.rows_per_batch > 0 ? row % p.rows_per_batch : row;
    const bool rotate = (p.cos != nullptr) && pos >= p.rope_row0 && (pos - p.rope_row0) < p.rope_rows;
    float cs[4] = {1.f, 1.f, 1.f, 1.f}, sn[4] = {0.f, 0.f, 0.f, 0.f};
    if (rotate) {
        const int64_t t = (pos - p.rope_row0) * (p.head_dim >> 1) + (ch >> 1);
        const float4 c4 = *reinterpret_cast<const float4*>(p.cos + t);
        const float4 s4 = *reinterpret_cast<const float4*>(p.sin + t);
        cs[0] = c4.x, cs[1] = c4.y, cs[2] = c4.z, cs[3] = c4.w;
        sn[0] = s4.x, sn[1] = s4.y, sn[2] = s4.z, sn[3] = s4.w;
    }

    // Scatter mode (Ulysses exchange #1 fused into this kernel): the result vector of head h goes to rank s = h / (H/P), into
    // its receive buffer [slot = this rank][slot row][q|k|v][(H/P) * D] — a 16-byte store over NVLink instead of a store into
    // the local buffer followed by a pack kernel and an NCCL all-to-all.  The V item is a pure copy.
    const int hp = scatter ? d / p.nsplit : d;  // channels per destination rank
    auto dst_of = [&](int v) -> __nv_bfloat16* {
        const int c0 = 8 * v;
        const int s = c0 / hp;
        return p.dst[s] + ((p.dst_slot * p.slot_rows + p.dst_row0 + row) * 3 + which) * hp + (c0 - s * hp);
    };
    if (which == 2) {
        const __nv_bfloat16* vr = p.v + row * p.row_stride;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = tl + kStride * i;
            if (v < nvec) *reinterpret_cast<uint4*>(dst_of(v)) = ld_nc_v4(vr + 8 * v);
        }
        return;
    }
    {
        __nv_bfloat16* xr = (which == 0 ? p.q : p.k) + row * p.row_stride;
        const float* w = which == 0 ? p.wq : p.wk;
        const float* b = which == 0 ? p.bq : p.bk;
        uint4 raw[MAXV];
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = tl + kStride * i;
            if (v < nvec) raw[i] = *reinterpret_cast<const uint4*>(xr + 8 * v);
        }
        float rs = 0.f;  // Wan: rsqrt(mean(x^2) + eps) over the whole row
        if (!COG) {
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                if (tl + kStride * i < nvec) {
                    const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = bf16x2_to_float2(u[j]);
                        sq += f.x * f.x + f.y * f.y;
                    }
                }
            }
            sq = warp_sum(sq);
            if constexpr (WPR > 1) {
                if ((threadIdx.x & 31) == 0) red[warp] = sq;
                named_bar_sync(1 + group, 32 * WPR);
                sq = 0.f;
#pragma unroll
                for (int ww = 0; ww < WPR; ++ww) sq += red[group * WPR + ww];
            }
            rs = rsqrtf(sq / static_cast<float>(d) + p.eps);
        }
        float wl[8], bl[8];
        if (COG) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                wl[j] = w[ch + j];
                bl[j] = b[ch + j];
            }
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = tl + kStride * i;
            // NOTE: the shuffles below need all lanes of a head group; nvec is a multiple of vec_per_head and
            // vec_per_head divides 32, so a head is never split between an active and an inactive lane.
            const bool active = v < nvec;
            float y[8];
            {
                const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = active ? bf16x2_to_float2(u[j]) : make_float2(0.f, 0.f);
                    y[2 * j] = f.x;
                    y[2 * j + 1] = f.y;
                }
            }
            if (COG) {
                // per-head LayerNorm: reduce over the vec_per_head lanes holding this head
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += y[j];
                for (int o = vec_per_head >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s / static_cast<float>(p.head_dim);
                float q2 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a = y[j] - mean;
                    q2 += a * a;
                }
                for (int o = vec_per_head >> 1; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
                const float rstd = rsqrtf(q2 / static_cast<float>(p.head_dim) + p.eps);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = bf16_round((y[j] - mean) * rstd * wl[j] + bl[j]);
            } else {
                if (active) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + 8 * v), w1 = *reinterpret_cast<const float4*>(w + 8 * v + 4);
                    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = bf16_round(bf16_round(y[j] * rs) * ww[j]);
                }
            }
            if (active) {
                if (rotate) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xr_ = y[2 * j], xi_ = y[2 * j + 1];
                        y[2 * j] = xr_ * cs[j] - xi_ * sn[j];
                        y[2 * j + 1] = xi_ * cs[j] + xr_ * sn[j];
                    }
                }
                uint4 o;
                o.x = pack_bf16x2(y[0], y[1]);
                o.y = pack_bf16x2(y[2], y[3]);
                o.z = pack_bf16x2(y[4], y[5]);
                o.w = pack_bf16x2(y[6], y[7]);
                *reinterpret_cast<uint4*>(scatter ? dst_of(v) : xr + 8 * v) = o;
            }
        }
    }
}


// Staged variant of the Wan mode (RMSNorm across all heads + RoPE, in place, no scatter): same structure as
// adaln_layernorm_staged_kernel — one work item = (token row, q or k), a private bulk-copy ring per warp, the two norm-weight vectors
// staged in shared memory once per CTA.  Lane l always owns channels 8 l .. 8 l + 7 of a head (256 i mod D == 0), so its four
// (cos, sin) pairs are one float4 each per row.
template <int NV, int kStages>
__global__ void __launch_bounds__(kStagedWarps * 32, 1) qk_norm_rope_staged_kernel(QkParams p, int ctas_per_batch) {
    extern __shared__ __align__(128) uint8_t qk_smem[];
    constexpr int d = 256 * NV;
    constexpr uint32_t row_bytes = d * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_which = p.k ? 2 : 1;
    const StagedRows rg(p.rows, p.rows_per_batch, ctas_per_batch);
    float* vec = reinterpret_cast<float*>(qk_smem);  // wq | wk
    const uint32_t ring = smem_u32(qk_smem) + 2u * d * sizeof(float) + static_cast<uint32_t>(warp) * kStages * row_bytes;
    const uint32_t bars = smem_u32(qk_smem) + 2u * d * sizeof(float) + kStagedWarps * kStages * row_bytes + static_cast<uint32_t>(warp) * kStages * 8u;
    for (int f = threadIdx.x; f < d / 4; f += kStagedWarps * 32) {  // split lo / hi layout, see adaln_layernorm_staged_kernel
        const int o = (f & 1) * (d / 2) + (f >> 1) * 4;
        *reinterpret_cast<float4*>(vec + o) = *reinterpret_cast<const float4*>(p.wq + 4 * f);
        if (p.k) *reinterpret_cast<float4*>(vec + d + o) = *reinterpret_cast<const float4*>(p.wk + 4 * f);
    }
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + 8u * s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t n_items = (rg.r1 - rg.r0) * n_which;  // item t of this CTA: row r0 + t / n_which, tensor t % n_which
    auto item_ptr = [&](int64_t t) -> __nv_bfloat16* {
        const int64_t row = rg.r0 + t / n_which;
        return ((t % n_which) == 0 ? p.q : p.k) + row * p.row_stride;
    };
    auto issue = [&](int64_t k) {
        const int64_t t = warp + kStagedWarps * k;
        if (t < n_items && lane == 0) {
            const uint32_t st = static_cast<uint32_t>(k % kStages);
            mbar_arrive_expect_tx(bars + 8u * st, row_bytes);
            bulk_load_row(ring + st * row_bytes, item_ptr(t), row_bytes, bars + 8u * st);
        }
    };
#pragma unroll
    for (int k = 0; k < kStages; ++k) issue(k);
    const int ch = (8 * lane) % p.head_dim;  // channel offset inside the head, the same for every vector of this lane
    const float inv_d = 1.f / static_cast<float>(d);
    for (int64_t k = 0;; ++k) {
        const int64_t t = warp + kStagedWarps * k;
        if (t >= n_items) break;
        const int64_t row = rg.r0 + t / n_which;
        const int which = static_cast<int>(t % n_which);
        const int64_t pos = row - rg.batch * p.rows_per_batch;
        const bool rotate = (p.cos != nullptr) && pos >= p.rope_row0 && (pos - p.rope_row0) < p.rope_rows;
        float4 c4 = make_float4(1.f, 1.f, 1.f, 1.f), s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rotate) {
            const int64_t off = (pos - p.rope_row0) * (p.head_dim >> 1) + (ch >> 1);
            c4 = *reinterpret_cast<const float4*>(p.cos + off);
            s4 = *reinterpret_cast<const float4*>(p.sin + off);
        }
        const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t st = static_cast<uint32_t>(k % kStages);
        mbar_wait(bars + 8u * st, static_cast<uint32_t>(k / kStages) & 1u);
        uint4 raw[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) raw[i] = ld_shared_v4_u32(ring + st * row_bytes + (lane + 32 * i) * 16);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(u[j]);
                sq += f.x * f.x + f.y * f.y;
            }
        }
        __syncwarp();  // the row is in registers (sq consumed every vector): refill its stage with the item kStages ahead
        issue(k + kStages);
        const float rs = rsqrtf(warp_sum(sq) * inv_d + p.eps);
        __nv_bfloat16* xr = item_ptr(t);
        const float* w = vec + which * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 8 * (lane + 32 * i);
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
            const float4 w0 = *reinterpret_cast<const float4*>(w + c / 2), w1 = *reinterpret_cast<const float4*>(w + d / 2 + c / 2);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float y[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(u[j]);
                y[2 * j] = bf16_round(bf16_round(f.x * rs) * ww[2 * j]);
                y[2 * j + 1] = bf16_round(bf16_round(f.y * rs) * ww[2 * j + 1]);
            }
            if (rotate) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xr_ = y[2 * j], xi_ = y[2 * j + 1];
                    y[2 * j] = xr_ * cs[j] - xi_ * sn[j];
                    y[2 * j + 1] = xi_ * cs[j] + xr_ * sn[j];
                }
            }
            uint4 o;
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            o.z = pack_bf16x2(y[4], y[5]);
            o.w = pack_bf16x2(y[6], y[7]);
            st_v4(xr + c, o);
        }
    }
}

template <int NV, int kStages>
static int launch_qk_staged(const QkParams& p, int ctas_per_batch, int64_t nbatch, cudaStream_t stream) {
    constexpr int smem = 2 * 256 * NV * 4 + kStagedWarps * kStages * 256 * NV * 2 + kStagedWarps * kStages * 8;
    static_assert(smem <= 232448, "shared memory budget");
    static bool opted_in[64] = {};
    if (int rc = smem_opt_in(qk_norm_rope_staged_kernel<NV, kStages>, smem, opted_in)) return rc;
    qk_norm_rope_staged_kernel<NV, kStages><<<static_cast<unsigned>(nbatch * ctas_per_batch), kStagedWarps * 32, smem, stream>>>(p, ctas_per_batch);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_qk_norm_rope(const QkParams& p, int cog_mode, cudaStream_t stream) {
    const int d = p.heads * p.head_dim;
    VAP_REQUIRE(p.head_dim == 64 || p.head_dim == 128 || p.head_dim == 32 || p.head_dim == 256,
                "qk_norm_rope: head_dim=%d must be 32, 64, 128 or 256", p.head_dim);
    VAP_REQUIRE(d <= 8192, "qk_norm_rope: heads*head_dim=%d must be <= 8192", d);
    VAP_REQUIRE(p.row_stride % 8 == 0, "qk_norm_rope: row stride must be a multiple of 8 elements");
    VAP_REQUIRE((p.cos == nullptr) == (p.sin == nullptr), "qk_norm_rope: cos and sin must both be given or both null");
    VAP_REQUIRE(p.wq && (p.wk || !p.k), "qk_norm_rope: norm weights are required");
    if (cog_mode) VAP_REQUIRE(p.bq && (p.bk || !p.k), "qk_norm_rope: per-head LayerNorm needs biases");
    if (p.rows == 0) return 0;
    const int nvec = d / 8;
    if (p.nsplit > 0) {
        VAP_REQUIRE(p.nsplit <= 8 && p.heads % p.nsplit == 0, "qkv_scatter: %d heads are not divisible by %d ranks (max 8)", p.heads, p.nsplit);
        VAP_REQUIRE(p.k && p.v, "qkv_scatter: q, k and v are required");
        for (int s = 0; s < p.nsplit; ++s) VAP_REQUIRE(p.dst[s] && (reinterpret_cast<uintptr_t>(p.dst[s]) & 15) == 0, "qkv_scatter: bad destination %d", s);
    }
    if (!cog_mode && p.nsplit == 0 && (d == 5120 || d == 4096) && 256 % p.head_dim == 0 && !staged_disabled() &&
        (reinterpret_cast<uintptr_t>(p.q) & 15) == 0 && (!p.k || (reinterpret_cast<uintptr_t>(p.k) & 15) == 0)) {
        int64_t nbatch = 1;
        QkParams q = p;
        if (p.rows_per_batch <= 0 || p.rows_per_batch > p.rows) q.rows_per_batch = p.rows;
        const int per = staged_ctas_per_batch(q.rows, q.rows_per_batch, &nbatch);
        if (per > 0) return d == 5120 ? launch_qk_staged<20, 2>(q, per, nbatch, stream) : launch_qk_staged<16, 2>(q, per, nbatch, stream);
    }
    const int64_t items = p.rows * (p.nsplit > 0 ? 3 : (p.k ? 2 : 1));
    const dim3 block(kWarpsPerBlock * 32);
#define VAP_QK_LAUNCH(MV, WPR)                                                                                   \
    do {                                                                                                         \
        const unsigned grid = static_cast<unsigned>((items + kWarpsPerBlock / WPR - 1) / (kWarpsPerBlock / WPR)); \
        if (cog_mode)                                                                                            \
            qk_norm_rope_kernel<MV, WPR, true><<<grid, block, 0, stream>>>(p);                                   \
        else                                                                                                     \
            qk_norm_rope_kernel<MV, WPR, false><<<grid, block, 0, stream>>>(p);                                  \
    } while (0)
    if (nvec <= 32 * 2)
        VAP_QK_LAUNCH(2, 1);
    else if (nvec <= 32 * 8)
        VAP_QK_LAUNCH(8, 1);
    else if (nvec <= 128 * 3)
        VAP_QK_LAUNCH(3, 4);
    else if (nvec <= 128 * 5)
        VAP_QK_LAUNCH(5, 4);
    else
        VAP_QK_LAUNCH(8, 4);
#undef VAP_QK_LAUNCH
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Classifier-free guidance + FlowMatchEuler update of the Wan denoise loop in one pass.
//   n   = u + g * (c - u)          three bf16 tensor ops in the reference (pipeline_wan_i2v_mot.py:874): one rounding each
//   out = sample.float() + dt * n  dt * n is a bf16 tensor op (0-dim fp32 dt, bf16 n), the sum is fp32, the result is cast back
//                                  to the model's dtype (scheduling_flow_match_euler_discrete.py:433, 457, 462-467)
// HBM-bound: reads 2 + 2 (+2 / +4) bytes per element, writes 2; one thread per 8 elements, 16-byte accesses.  The output rows may
// sit inside the next step's transformer input [B, C_latent + C_cond, F, h, w] (out_batch_stride), which removes that torch.cat.
// __f*_rn intrinsics keep ptxas from contracting the separately rounded multiplies and adds into FMAs.
// ------------------------------------------------------------------------------------------------------------------------
template <bool kSampleF32>
__global__ void __launch_bounds__(256) cfg_flow_match_kernel(const StepParams p) {
    const int64_t vec_per_batch = p.inner / 8;
    const int64_t total = p.batch * vec_per_batch;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t b = idx / vec_per_batch;
        const int64_t e = (idx - b * vec_per_batch) * 8;  // element inside the batch
        const int64_t in_off = b * p.inner + e;
        float n[8], x[8];
        {
            const uint4 cu = ld_nc_v4(p.cond + in_off);
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cu);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(cw[j]);
                n[2 * j] = f.x, n[2 * j + 1] = f.y;
            }
        }
        if (p.uncond) {
            const uint4 uu = ld_nc_v4(p.uncond + in_off);
            const uint32_t* uw = reinterpret_cast<const uint32_t*>(&uu);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(uw[j]);
                const float u2[2] = {f.x, f.y};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float d = bf16_round(__fsub_rn(n[2 * j + h], u2[h]));
                    const float m = bf16_round(__fmul_rn(p.guidance, d));
                    n[2 * j + h] = bf16_round(__fadd_rn(u2[h], m));
                }
            }
        }
        if (kSampleF32) {
            const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(p.sample) + in_off);
            const float4 c = *reinterpret_cast<const float4*>(static_cast<const float*>(p.sample) + in_off + 4);
            x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = c.x, x[5] = c.y, x[6] = c.z, x[7] = c.w;
        } else {
            const uint4 su = ld_nc_v4(static_cast<const __nv_bfloat16*>(p.sample) + in_off);
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(&su);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf16x2_to_float2(sw[j]);
                x[2 * j] = f.x, x[2 * j + 1] = f.y;
            }
        }
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __fadd_rn(x[j], bf16_round(__fmul_rn(p.dt, n[j])));
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]);
        o.y = pack_bf16x2(y[2], y[3]);
        o.z = pack_bf16x2(y[4], y[5]);
        o.w = pack_bf16x2(y[6], y[7]);
        st_v4(p.out + b * p.out_batch_stride + e, o);
    }
}

int launch_cfg_flow_match_step(const StepParams& p, cudaStream_t stream) {
    VAP_REQUIRE(p.batch >= 0 && p.inner >= 0 && p.inner % 8 == 0, "cfg_flow_match_step: inner=%lld must be a non-negative multiple of 8",
                static_cast<long long>(p.inner));
    VAP_REQUIRE(p.out_batch_stride >= p.inner && p.out_batch_stride % 8 == 0, "cfg_flow_match_step: bad output batch stride");
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(p.cond) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.uncond) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.sample) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0,
                "cfg_flow_match_step: tensors must be 16-byte aligned");
    const int64_t total = p.batch * (p.inner / 8);
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    if (p.sample_is_f32)
        cfg_flow_match_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    else
        cfg_flow_match_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// W1: the adaLN modulation of a Wan block, (scale_shift_table[1, C, d] + temb[B, C, d].float()).chunk(C) with the "+ 1" of the two scale
// chunks folded in (transformer_wan_mot.py:606-616, 620-622, 680-689) — one launch instead of three eager ones per stream and block.
//   out[b, c, :] = float(table[c, :]) + float(temb[b, c, :]) (+ 1 when bit c of plus_one_mask is set)        fp32, like the reference
// table / temb are bf16 or fp32 (a from_pretrained model keeps scale_shift_table in fp32, a .to(bfloat16) one does not).
// ------------------------------------------------------------------------------------------------------------------------
template <bool kTableF32, bool kTembF32>
__global__ void __launch_bounds__(256) wan_modulation_kernel(const void* table, const void* temb, float* out, int64_t batch, int chunks, int d, unsigned plus_one_mask) {
    const int64_t per = static_cast<int64_t>(chunks) * d;
    const int64_t total = batch * per;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t e = idx % per;
        const int c = static_cast<int>(e / d);
        const float t = kTableF32 ? static_cast<const float*>(table)[e] : __bfloat162float(static_cast<const __nv_bfloat16*>(table)[e]);
        const float m = kTembF32 ? static_cast<const float*>(temb)[idx] : __bfloat162float(static_cast<const __nv_bfloat16*>(temb)[idx]);
        float v = __fadd_rn(t, m);
        if ((plus_one_mask >> c) & 1u) v = __fadd_rn(1.f, v);
        out[idx] = v;
    }
}

int launch_wan_modulation(const void* table, int table_is_f32, const void* temb, int temb_is_f32, float* out, int64_t batch, int chunks, int d,
                          unsigned plus_one_mask, cudaStream_t stream) {
    VAP_REQUIRE(batch >= 0 && chunks > 0 && chunks <= 32 && d > 0, "wan_modulation: bad shape batch=%lld chunks=%d d=%d", static_cast<long long>(batch), chunks, d);
    const int64_t total = batch * chunks * d;
    if (total == 0) return 0;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    if (table_is_f32) {
        if (temb_is_f32) wan_modulation_kernel<true, true><<<blocks, 256, 0, stream>>>(table, temb, out, batch, chunks, d, plus_one_mask);
        else wan_modulation_kernel<true, false><<<blocks, 256, 0, stream>>>(table, temb, out, batch, chunks, d, plus_one_mask);
    } else {
        if (temb_is_f32) wan_modulation_kernel<false, true><<<blocks, 256, 0, stream>>>(table, temb, out, batch, chunks, d, plus_one_mask);
        else wan_modulation_kernel<false, false><<<blocks, 256, 0, stream>>>(table, temb, out, batch, chunks, d, plus_one_mask);
    }
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vap

// probe_sm100.cu — single-CTA bring-up probe for the tcgen05 operand encodings used by gemm_sm100.cu and
// attn_sm100.cu: D[128,N] (fp32) = A[128,K] * B with
//   A from shared memory (K-major, SWIZZLE_128B)            or  A staged into TMEM as packed bf16 pairs,
//   B = [N,K] K-major (QK^T / Linear weights, SWIZZLE_128B) or  B = [K,N] MN-major (the V operand of P*V).
// The B descriptor's LBO / SBO / per-k-step advance / layout type are runtime arguments, so a test can
// confirm the encoding the production kernels hard-code (and sweep alternatives if a driver ever differs).
#include "vap_kernels.cuh"

namespace vap {


__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = smem_base;                 // up to 2 slabs of 128 x 128 B
    const uint32_t b_smem = smem_base + 2 * 16384;     // up to 64 KB
    const uint32_t bar_base = b_smem + 65536;
    const uint32_t load_bar = bar_base, mma_bar = bar_base + 8, tmem_ptr_addr = bar_base + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_slabs = p.K / 64;

    if (threadIdx.x == 0) {
        mbar_init(load_bar, 1);
        mbar_init(mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_addr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

    if (threadIdx.x == 0) {
        uint32_t bytes = 0;
        if (!p.a_in_tmem) bytes += k_slabs * 16384;
        bytes += p.b_mn_major ? (p.N / 64) * p.K * 128 : k_slabs * p.N * 128;
        mbar_arrive_expect_tx(load_bar, bytes);
        if (!p.a_in_tmem)
            for (int s = 0; s < k_slabs; ++s) tma_load_2d(a_smem + s * 16384, &tmA, load_bar, s * 64, 0);
        if (p.b_mn_major) {
            for (int s = 0; s < p.N / 64; ++s) tma_load_2d(b_smem + s * p.K * 128, &tmB, load_bar, s * 64, 0);
        } else {
            for (int s = 0; s < k_slabs; ++s) tma_load_2d(b_smem + s * p.N * 128, &tmB, load_bar, s * 64, 0);
        }
    }
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    if (p.a_in_tmem && p.lane16_shapes) {
        // 16-lane shapes: per half hf the warp writes lanes 32*warp + 16*hf + {t/4, t/4+8}, packed column 4g + t%4 (K = 128)
        for (int hf = 0; hf < 2; ++hf) {
            const int r0 = warp * 32 + 16 * hf + (lane >> 2);
            const uint32_t* a0 = reinterpret_cast<const uint32_t*>(p.A + static_cast<int64_t>(r0) * p.K);
            const uint32_t* a1 = reinterpret_cast<const uint32_t*>(p.A + static_cast<int64_t>(r0 + 8) * p.K);
            uint32_t v[32];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                v[2 * g] = a0[4 * g + (lane & 3)];
                v[2 * g + 1] = a1[4 * g + (lane & 3)];
            }
            tmem_st_16x128b_x16(tmem_base + (static_cast<uint32_t>(warp * 32 + 16 * hf) << 16) + 256, v);
        }
        tmem_st_wait();
    } else if (p.a_in_tmem) {
        // thread t owns row t: pack (A[t,2c], A[t,2c+1]) into TMEM column 256 + c
        const uint32_t* arow = reinterpret_cast<const uint32_t*>(p.A + static_cast<int64_t>(threadIdx.x) * p.K);
        for (int c0 = 0; c0 < p.K / 2; c0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = arow[c0 + e];
            tmem_st_x16(tmem_base + lane_addr + 256 + c0, v);
        }
        tmem_st_wait();
    }
    mbar_wait(load_bar, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, p.N, 0, p.b_mn_major ? 1 : 0);
        for (int k = 0; k < p.K / 16; ++k) {
            uint64_t db;
            if (p.b_mn_major)
                db = make_smem_desc(b_smem + k * p.kstep_b, p.lbo_b, p.sbo_b, p.layout_type);
            else
                db = make_smem_desc(b_smem + (k >> 2) * p.N * 128 + (k & 3) * 32, p.lbo_b, p.sbo_b, p.layout_type);
            if (p.a_in_tmem) {
                umma_ts(tmem_base, tmem_base + 256 + 8 * k, db, idesc, k != 0);
            } else {
                const uint64_t da = make_smem_desc(a_smem + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, kLayoutSw128);
                umma_ss(tmem_base, da, db, idesc, k != 0);
            }
        }
        umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    if (p.lane16_shapes) {
        // read D back with 16x256b: r[4g+0..1] = (lane t/4, cols 8g + 2(t%4) + {0,1}), r[4g+2..3] = lane t/4 + 8
        for (int hf = 0; hf < 2; ++hf) {
            const int r0 = warp * 32 + 16 * hf + (lane >> 2);
            for (int c0 = 0; c0 < p.N; c0 += 32) {
                uint32_t v[16];
                tmem_ld_16x256b_x4(tmem_base + (static_cast<uint32_t>(warp * 32 + 16 * hf) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        p.Dout[static_cast<int64_t>(r0 + (e >> 1) * 8) * p.N + c0 + 8 * g + 2 * (lane & 3) + (e & 1)] = __uint_as_float(v[4 * g + e]);
            }
        }
    } else {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_addr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) p.Dout[static_cast<int64_t>(row) * p.N + c0 + e] = __uint_as_float(v[e]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int launch_probe_umma(const __nv_bfloat16* A, const __nv_bfloat16* B, ProbeParams p, cudaStream_t stream) {
    VAP_REQUIRE(p.K == 64 || p.K == 128, "probe: K must be 64 or 128");
    VAP_REQUIRE(p.N % 64 == 0 && p.N >= 64 && p.N <= 256, "probe: N must be 64, 128, 192 or 256");
    VAP_REQUIRE(!p.b_mn_major || p.N <= 128, "probe: MN-major B supports N <= 128");
    VAP_REQUIRE(!p.lane16_shapes || (p.K == 128 && p.N % 32 == 0), "probe: the 16-lane TMEM shapes need K = 128");
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), 128};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.K)};
        const uint32_t box[2] = {64, 128};
        if (make_tmap_bf16(&tmA, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    }
    if (p.b_mn_major) {  // B [K, N], N contiguous: box = 64 columns x K rows
        const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.K)};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.N)};
        const uint32_t box[2] = {64, static_cast<uint32_t>(p.K)};
        if (make_tmap_bf16(&tmB, B, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    } else {  // B [N, K], K contiguous: box = 64 columns x N rows
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.K)};
        const uint32_t box[2] = {64, static_cast<uint32_t>(p.N)};
        if (make_tmap_bf16(&tmB, B, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    }
    const int smem = 2 * 16384 + 65536 + 1024 + 64;
    static bool opted_in[1][64] = {};
    if (int rc = smem_opt_in(probe_umma_kernel, smem, opted_in[0])) return rc;
    probe_umma_kernel<<<1, 128, smem, stream>>>(tmA, tmB, p);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vap

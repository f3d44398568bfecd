// mma_rate_probe.cu — how fast does ONE issuing thread drive tcgen05.mma (cta_group::1, kind::f16, M = 128) on sm_100a?
// Not part of the product: it measures the floors the attention kernel's MMA warp lives with.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate_probe.bin tools/mma_rate_probe.cu -lcuda
// One CTA per SM (all SMs busy, so power / clocks are realistic).  Operands are random bf16 in shared memory (SWIZZLE_128B
// tiles as the attention kernel lays them out) or in TMEM; the accumulators are garbage on purpose.  Prints SM clocks per
// MMA for each variant:
//   ss128    S = Q K^T style:  A smem K-major, B smem K-major, N = 128
//   ss64     the same with N = 64 (the double-buffered-S experiment)
//   ss256    N = 256
//   ts128    O += P V style:   A from TMEM, B smem MN-major, N = 128
//   ts64k    A from TMEM, B smem K-major, N = 64 (Q held in TMEM)
//   mix      4 x ts128 + 8 x ss64 (one KV step of the N = 64 kernel)
//   fa       8 x ts128 + 8 x ss128 (one KV step of the N = 128 kernel)
//   fa+c     fa with a tcgen05.commit after every 16 MMAs
//   fa+cw    fa+c plus a try_wait on an already completed mbarrier and tcgen05.fence::after_thread_sync per group
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../video-as-prompt_b200/csrc/vap_common.cuh"

using namespace vap;
namespace vap { void set_error(const char*, ...) {} }

enum { SS128, SS64, SS256, TS128, TS64K, MIX, FA, FA_C, FA_CW, NVAR };
static const char* kNames[NVAR] = {"ss128", "ss64", "ss256", "ts128", "ts64k", "mix", "fa", "fa+c", "fa+cw"};

constexpr int kReps = 64;  // groups per measurement

__global__ void __launch_bounds__(128, 1) probe(int variant, long long* out, int* mmas_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = base;                 // 2 x 32 KB
    const uint32_t kv_smem = base + 65536;        // 4 x 32 KB
    const uint32_t bar = base + 65536 + 131072;   // done barrier
    const uint32_t bar2 = bar + 8;                // dummy commit target
    const uint32_t bar3 = bar + 16;               // pre-completed barrier
    const uint32_t tptr = bar + 32;
    // fill shared memory with pseudo-random bf16 in [-1, 1)
    uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
    for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        const uint32_t lo = 0x3f00u | ((s >> 9) & 0x7fu) | ((s >> 3) & 0x8000u);
        const uint32_t hi = 0x3f00u | ((s >> 17) & 0x7fu) | ((s >> 11) & 0x8000u);
        w[i] = lo | (hi << 16);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        mbar_init(bar3, 1);
        fence_mbar_init();
        mbar_arrive(bar3);  // phase 0 complete
    }
    if (threadIdx.x < 32) {
        tmem_alloc(tptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
    // P-like data in TMEM columns 0..127
    {
        uint32_t v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0x3f003f00u + e * 0x00010001u;
        const uint32_t la = static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16;
        for (int c = 0; c < 128; c += 32) tmem_st_x32(tmem + la + c, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
        constexpr uint32_t id_ss128 = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t id_ss64 = make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t id_ss256 = make_idesc_bf16(128, 256, 0, 0);
        constexpr uint32_t id_ts128 = make_idesc_bf16(128, 128, 0, 1);
        auto ss = [&](int i, int tile, uint32_t idesc, uint32_t slab_b, uint32_t dcol) {  // 8 MMAs (K = 128)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_ss(tmem + dcol, make_smem_desc(q_smem + i * 32768 + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, kLayoutSw128),
                        make_smem_desc(kv_smem + tile * 32768 + (k >> 2) * slab_b + (k & 3) * 32, 0, 1024, kLayoutSw128), idesc, k != 0);
        };
        auto ts128 = [&](int tile, int nk, uint32_t slab_b, uint32_t dcol, uint32_t acol) {  // nk MMAs, B MN-major
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nk) umma_ts(tmem + dcol, tmem + acol + 8 * k, make_smem_desc(kv_smem + tile * 32768 + k * 2048, slab_b, 1024, kLayoutSw128), id_ts128, 1u);
        };
        auto ts64k = [&](int tile, uint32_t dcol, uint32_t acol) {  // 8 MMAs, A TMEM, B K-major N = 64
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_ts(tmem + dcol, tmem + acol + 8 * k, make_smem_desc(kv_smem + tile * 32768 + (k >> 2) * 8192 + (k & 3) * 32, 0, 1024, kLayoutSw128), id_ss64, k != 0);
        };
        int mmas = 0;
        const long long t0 = clock64();
        for (int r = 0; r < kReps; ++r) {
            const int t = r & 3;
            switch (variant) {
                case SS128: ss(r & 1, t, id_ss128, 16384, 128 * (r & 1)); ss(1 - (r & 1), t, id_ss128, 16384, 256); mmas += 16; break;
                case SS64: ss(r & 1, t, id_ss64, 8192, 64 * (r & 3)); ss(1 - (r & 1), t, id_ss64, 8192, 256); mmas += 16; break;
                case SS256: ss(r & 1, t & 1, id_ss256, 32768, 0); ss(1 - (r & 1), t & 1, id_ss256, 32768, 256); mmas += 16; break;
                case TS128: ts128(t, 8, 16384, 256, 0); ts128((t + 1) & 3, 8, 16384, 384, 64); mmas += 16; break;
                case TS64K: ts64k(t, 256, 0); ts64k((t + 1) & 3, 320, 64); mmas += 16; break;
                case MIX: ts128(t, 4, 8192, 256, 0); ss(0, (t + 1) & 3, id_ss64, 8192, 128); ts128(t, 4, 8192, 384, 64); ss(1, (t + 1) & 3, id_ss64, 8192, 192); mmas += 24; break;
                case FA:
                case FA_C:
                case FA_CW:
                    if (variant == FA_CW) {
                        while (!mbar_try_wait(bar3, 0)) {}
                        tc_fence_after();
                    }
                    ts128(t, 8, 16384, 256, 0);
                    ss(r & 1, (t + 1) & 3, id_ss128, 16384, 128);
                    if (variant != FA) umma_commit(bar2);
                    mmas += 16;
                    break;
            }
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
        if (blockIdx.x == 0) *mmas_out = mmas;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = 65536 + 131072 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d_out;
    int* d_mmas;
    cudaMalloc(&d_out, sizeof(long long) * 2 * sms);
    cudaMalloc(&d_mmas, sizeof(int));
    long long* h = static_cast<long long*>(malloc(sizeof(long long) * 2 * sms));
    printf("%-8s %10s %14s %14s\n", "variant", "MMAs", "clk/MMA issue", "clk/MMA done");
    for (int v = 0; v < NVAR; ++v) {
        int mmas = 0;
        double best_issue = 1e30, best_done = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            probe<<<sms, 128, smem>>>(v, d_out, d_mmas);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s: %s\n", kNames[v], cudaGetErrorString(e));
                return 1;
            }
            cudaMemcpy(h, d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
            cudaMemcpy(&mmas, d_mmas, sizeof(int), cudaMemcpyDeviceToHost);
            double si = 0, sd = 0;
            for (int i = 0; i < sms; ++i) si += h[2 * i], sd += h[2 * i + 1];
            si /= sms, sd /= sms;
            if (sd < best_done) best_done = sd, best_issue = si;
        }
        printf("%-8s %10d %14.1f %14.1f\n", kNames[v], mmas, best_issue / mmas, best_done / mmas);
    }
    return 0;
}

// softmax_pipe_probe.cu — micro-benchmark of the per-element instruction mix of the attention softmax on sm_100a.
// Not part of the product: it answers "which pipe paces p = 2^(s*c - m*c) -> bf16 pack -> row sum" on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/softmax_pipe_probe.bin tools/softmax_pipe_probe.cu
// Every variant works on 64 fp32 values held in registers per thread, repeated REPS times (one ALU-pipe LOP3 per value flips
// its mantissa LSB each repetition so ptxas cannot hoist the work; the 'empty loop' row is that overhead).  Prints SM cycles per repetition
// per warp and per "share" (= 64 elements x 32 lanes) for W = 1, 2, 4 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video-as-prompt_b200/csrc/vap_common.cuh"

using namespace vap;
namespace vap { void set_error(const char*, ...) {} }

constexpr int REPS = 512;

enum { V_MUFU_BF16X2, V_MUFU_F16X2, V_EMPTY, V_FFMA, V_FFMA2, V_FADD2, V_FMNMX, V_FMNMX3, V_F2FP, V_MUFU, V_IMAD, V_SCALE_MUFU_PACK, V_SOFTMAX, V_SOFTMAX_NOSUM, V_POLY_NOCLAMP };

template <int VARIANT, int POLY>
__global__ void __launch_bounds__(512, 1) probe(const float* in, uint32_t* out, long long* cycles) {
    float s[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) s[e] = in[(threadIdx.x * 64 + e) & 4095];
    const float c = in[4096], nm = in[4097];
    const uint64_t c2 = pack_f32x2(c, c), nm2 = pack_f32x2(nm, nm);
    uint64_t l2 = 0;
    uint32_t acc = 0;
    float mx = -1e30f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REPS; ++r) {
#pragma unroll
        for (int e = 0; e < 64; ++e) s[e] = __int_as_float(__float_as_int(s[e]) ^ (r & 1));  // ALU-pipe LOP3: fresh values every rep (ptxas hoists otherwise)
        if (VARIANT == V_EMPTY) {
        } else if (VARIANT == V_FFMA) {
#pragma unroll
            for (int e = 0; e < 64; ++e) s[e] = fmaf(s[e], c, nm);
        } else if (VARIANT == V_FFMA2) {
#pragma unroll
            for (int g = 0; g < 32; ++g) unpack_f32x2(fma_f32x2(pack_f32x2(s[2 * g], s[2 * g + 1]), c2, nm2), s[2 * g], s[2 * g + 1]);
        } else if (VARIANT == V_FADD2) {
#pragma unroll
            for (int g = 0; g < 32; ++g) unpack_f32x2(add_f32x2(pack_f32x2(s[2 * g], s[2 * g + 1]), nm2), s[2 * g], s[2 * g + 1]);
        } else if (VARIANT == V_FMNMX) {
            float a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3];
#pragma unroll
            for (int e = 4; e < 64; e += 4) a0 = fmaxf(a0, s[e]), a1 = fmaxf(a1, s[e + 1]), a2 = fmaxf(a2, s[e + 2]), a3 = fmaxf(a3, s[e + 3]);
            mx += fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        } else if (VARIANT == V_FMNMX3) {
            float a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3];
#pragma unroll
            for (int e = 4; e < 60; e += 8) {
                a0 = fmax3(a0, s[e], s[e + 1]), a1 = fmax3(a1, s[e + 2], s[e + 3]);
                a2 = fmax3(a2, s[e + 4], s[e + 5]), a3 = fmax3(a3, s[e + 6], s[e + 7]);
            }
            mx += fmax3(fmax3(a0, a1, s[60]), fmax3(a2, a3, s[61]), fmaxf(s[62], s[63]));
        } else if (VARIANT == V_F2FP) {
#pragma unroll
            for (int g = 0; g < 32; ++g) acc ^= pack_bf16x2(s[2 * g], s[2 * g + 1]);
        } else if (VARIANT == V_MUFU) {
#pragma unroll
            for (int e = 0; e < 64; ++e) s[e] = ex2_approx(s[e]);
        } else if (VARIANT == V_MUFU_BF16X2) {
#pragma unroll
            for (int g = 0; g < 32; ++g) {  // 64 elements as 32 packed pairs -> 64 MUFU.EX2.BF16
                uint32_t x = __float_as_uint(s[g]), y;
                asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
                s[g] = __uint_as_float(y);
            }
        } else if (VARIANT == V_MUFU_F16X2) {
#pragma unroll
            for (int g = 0; g < 32; ++g) {
                uint32_t x = __float_as_uint(s[g]), y;
                asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
                s[g] = __uint_as_float(y);
            }
        } else if (VARIANT == V_IMAD) {
#pragma unroll
            for (int e = 0; e < 64; ++e) s[e] = __int_as_float(__float_as_int(s[e]) * 0x800000 + __float_as_int(nm));
        } else {
            uint32_t pk[32];
#pragma unroll
            for (int g = 0; g < 32; ++g) {
                const uint64_t x2 = fma_f32x2(pack_f32x2(s[2 * g], s[2 * g + 1]), c2, nm2);
                float x0, x1, p0, p1;
                unpack_f32x2(x2, x0, x1);
                if ((g & 7) < POLY) {
                    if (VARIANT == V_POLY_NOCLAMP) {
                        const float kMagic = 12582912.f;
                        const uint64_t x = pack_f32x2(x0, x1);
                        const uint64_t rr = add_rm_f32x2(x, pack_f32x2(kMagic, kMagic));
                        const uint64_t n = sub_f32x2(rr, pack_f32x2(kMagic, kMagic));
                        const uint64_t f = sub_f32x2(x, n);
                        uint64_t pl = fma_f32x2(pack_f32x2(0.077119089663028717f, 0.077119089663028717f), f, pack_f32x2(0.227564394474029541f, 0.227564394474029541f));
                        pl = fma_f32x2(pl, f, pack_f32x2(0.695146143436431885f, 0.695146143436431885f));
                        pl = fma_f32x2(pl, f, pack_f32x2(1.0f, 1.0f));
                        float q0, q1, r0, r1;
                        unpack_f32x2(pl, q0, q1);
                        unpack_f32x2(rr, r0, r1);
                        p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
                        p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
                    } else {
                        ex2_poly_x2(x0, x1, p0, p1);
                    }
                } else {
                    p0 = ex2_approx(x0);
                    p1 = ex2_approx(x1);
                }
                if (VARIANT != V_SOFTMAX_NOSUM) l2 = add_f32x2(l2, pack_f32x2(p0, p1));
                pk[g] = pack_bf16x2(p0, p1);
            }
#pragma unroll
            for (int g = 0; g < 32; g += 2) acc += pk[g] ^ pk[g + 1];
        }
    }
    const long long t1 = clock64();
    float lo, hi;
    unpack_f32x2(l2, lo, hi);
    float sum = lo + hi + mx;
#pragma unroll
    for (int e = 0; e < 64; ++e) sum += s[e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(sum);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VARIANT, int POLY>
static void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
    printf("%-40s", name);
    for (int wps : {1, 2, 4}) {
        const int threads = wps * 4 * 32;
        probe<VARIANT, POLY><<<148, threads>>>(in, out, cyc);
        cudaDeviceSynchronize();
        probe<VARIANT, POLY><<<148, threads>>>(in, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" error %s", cudaGetErrorString(e)); continue; }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148.0 * REPS;
        printf("  W=%d: %7.1f clk/rep (%6.1f /share)", wps, avg, avg / wps);
    }
    printf("\n");
}

int main() {
    float* in;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&in, 4098 * sizeof(float));
    cudaMalloc(&out, 148 * 512 * sizeof(uint32_t));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    float h[4098];
    for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 977);
    h[4096] = 0.127f, h[4097] = -0.3f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("64 elements / thread / rep.  '/share' = sub-partition cycles per (64 elements x 32 lanes).\n");
    run<V_EMPTY, 0>("empty loop", in, out, cyc);
    run<V_FFMA, 0>("64 FFMA", in, out, cyc);
    run<V_FFMA2, 0>("32 FFMA2", in, out, cyc);
    run<V_FADD2, 0>("32 FADD2", in, out, cyc);
    run<V_FMNMX, 0>("63 FMNMX (row max)", in, out, cyc);
    run<V_FMNMX3, 0>("32 FMNMX3 (row max)", in, out, cyc);
    run<V_F2FP, 0>("32 F2FP.BF16 + 32 LOP3", in, out, cyc);
    run<V_MUFU, 0>("64 MUFU.EX2", in, out, cyc);
    run<V_IMAD, 0>("64 IMAD", in, out, cyc);
    run<V_MUFU_BF16X2, 0>("32 ex2.bf16x2 (64 MUFU.EX2.BF16)", in, out, cyc);
    run<V_MUFU_F16X2, 0>("32 ex2.f16x2 (64 MUFU.EX2.F16)", in, out, cyc);
    run<V_SOFTMAX_NOSUM, 0>("scale + MUFU + pack (no sum)", in, out, cyc);
    run<V_SOFTMAX, 0>("softmax poly 0/8", in, out, cyc);
    run<V_SOFTMAX, 2>("softmax poly 2/8", in, out, cyc);
    run<V_SOFTMAX, 3>("softmax poly 3/8", in, out, cyc);
    run<V_SOFTMAX, 4>("softmax poly 4/8", in, out, cyc);
    run<V_SOFTMAX, 8>("softmax poly 8/8", in, out, cyc);
    run<V_POLY_NOCLAMP, 3>("softmax poly 3/8 no clamp", in, out, cyc);
    run<V_POLY_NOCLAMP, 4>("softmax poly 4/8 no clamp", in, out, cyc);
    run<V_POLY_NOCLAMP, 8>("softmax poly 8/8 no clamp", in, out, cyc);
    return 0;
}

// capi.cu — extern "C" boundary of libvap_b200.so (declared in include/vap_b200.h) plus the host-side helpers
// shared by the kernels: thread-local error message, TMA descriptor encoding through the driver entry point
// (no link-time dependency on libcuda, so the library loads on a CPU-only box), Ulysses re-layout kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/vap_b200.h"
#include "vap_kernels.cuh"

namespace vap {

static thread_local char g_err[512] = "";
static long long* g_attn_trace = nullptr;  // debug: device buffer for the attention kernel's clock64() trace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !sym) {
            set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -3;
    cuuint64_t gdim[5];
    cuuint64_t gstride[4];
    cuuint32_t bdim[5];
    cuuint32_t estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estride[i] = 1;
        if (i > 0) gstride[i - 1] = strides_elems[i - 1] * sizeof(__nv_bfloat16);
    }
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstride,
                          bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", static_cast<int>(r), rank,
                  static_cast<unsigned long long>(dims[0]), static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0],
                  rank > 1 ? box[1] : 0);
        return -3;
    }
    return 0;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------
// Ulysses re-layout: strided [L, nsplit, chunk] <-> [nsplit, L, chunk], 16-byte vectors, grid-stride
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ulysses_permute_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t L,
                                                              int nsplit, int64_t chunk_v, int64_t wide_row_v, int64_t split_row_v,
                                                              int64_t split_stride_v, int unpack) {
    const int64_t total = L * nsplit * chunk_v;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t c = idx % chunk_v;
        const int64_t l = (idx / chunk_v) % L;
        const int64_t s = idx / (chunk_v * L);
        const int64_t wide = l * wide_row_v + s * chunk_v + c;               // [L, nsplit*chunk] side
        const int64_t split = s * split_stride_v + l * split_row_v + c;      // [nsplit, L, chunk] side
        if (unpack)
            dst[wide] = src[split];
        else
            dst[split] = src[wide];
    }
}

static int launch_ulysses(const void* src, void* dst, int64_t L, int nsplit, int64_t chunk, int64_t wide_row, int64_t split_row,
                          int64_t split_stride, int unpack, cudaStream_t stream) {
    VAP_REQUIRE(chunk > 0 && chunk % 8 == 0 && wide_row % 8 == 0 && split_row % 8 == 0 && split_stride % 8 == 0,
                "ulysses: chunk and strides must be multiples of 8 elements");
    VAP_REQUIRE(nsplit > 0 && L >= 0 && wide_row >= static_cast<int64_t>(nsplit) * chunk && split_row >= chunk, "ulysses: bad shape");
    VAP_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "ulysses: buffers must be 16-byte aligned");
    const int64_t total = L * nsplit * (chunk / 8);
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    ulysses_permute_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), L,
                                                                             nsplit, chunk / 8, wide_row / 8, split_row / 8,
                                                                             split_stride / 8, unpack);
    VAP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vap

using namespace vap;

extern "C" {

int vap_version(void) { return VAP_B200_VERSION; }
const char* vap_last_error(void) { return g_err; }
int vap_sm_count(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("vap_sm_count: no CUDA device");
        return -2;
    }
    return sm_count();
}

int vap_adaln_layernorm(const void* x, void* out, int64_t rows, int d, int64_t x_row_stride, int64_t out_row_stride,
                        const float* ln_w, const float* ln_b, const float* scale1p, const float* shift, int64_t mod_stride,
                        int64_t rows_per_batch, float eps, int rounding, void* stream) {
    VAP_REQUIRE(x && out, "vap_adaln_layernorm: null tensor");
    VAP_REQUIRE(rows >= 0, "vap_adaln_layernorm: rows < 0");
    LnParams p{static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), rows, d, x_row_stride, out_row_stride, ln_w, ln_b,
               scale1p, shift, mod_stride, rows_per_batch, eps, rounding};
    return launch_adaln_layernorm(p, static_cast<cudaStream_t>(stream));
}

int vap_qk_norm_rope(void* q, void* k, int64_t rows, int heads, int head_dim, int64_t row_stride, const float* wq, const float* bq,
                     const float* wk, const float* bk, const float* cos, const float* sin, int64_t rows_per_batch, int64_t rope_row0,
                     int64_t rope_rows, float eps, int mode, void* stream) {
    VAP_REQUIRE(q, "vap_qk_norm_rope: null tensor");
    VAP_REQUIRE(rows >= 0 && heads > 0, "vap_qk_norm_rope: bad shape");
    VAP_REQUIRE(mode == 0 || mode == 1, "vap_qk_norm_rope: mode must be 0 (Wan) or 1 (CogVideoX)");
    QkParams p{};
    p.q = static_cast<__nv_bfloat16*>(q), p.k = static_cast<__nv_bfloat16*>(k);
    p.rows = rows, p.heads = heads, p.head_dim = head_dim, p.row_stride = row_stride;
    p.wq = wq, p.bq = bq, p.wk = wk, p.bk = bk, p.cos = cos, p.sin = sin;
    p.rows_per_batch = rows_per_batch, p.rope_row0 = rope_row0, p.rope_rows = rope_rows, p.eps = eps;
    return launch_qk_norm_rope(p, mode, static_cast<cudaStream_t>(stream));
}

int vap_qkv_scatter(const void* q, const void* k, const void* v, int64_t rows, int heads, int head_dim, int64_t row_stride, const float* wq,
                    const float* bq, const float* wk, const float* bk, const float* cos, const float* sin, int64_t rows_per_batch,
                    int64_t rope_row0, int64_t rope_rows, float eps, int mode, void* const* dst, int nsplit, int64_t dst_slot,
                    int64_t slot_rows, int64_t dst_row0, void* stream) {
    VAP_REQUIRE(q && k && v && dst, "vap_qkv_scatter: null tensor");
    VAP_REQUIRE(rows >= 0 && heads > 0, "vap_qkv_scatter: bad shape");
    VAP_REQUIRE(mode == 0 || mode == 1, "vap_qkv_scatter: mode must be 0 (Wan) or 1 (CogVideoX)");
    VAP_REQUIRE(nsplit >= 1 && nsplit <= 8, "vap_qkv_scatter: nsplit=%d must be in [1, 8]", nsplit);
    VAP_REQUIRE(dst_slot >= 0 && slot_rows > 0 && dst_row0 >= 0 && dst_row0 + rows <= slot_rows, "vap_qkv_scatter: rows do not fit the slot");
    QkParams p{};
    // q and k are only READ in scatter mode (the kernel stores to dst)
    p.q = static_cast<__nv_bfloat16*>(const_cast<void*>(q)), p.k = static_cast<__nv_bfloat16*>(const_cast<void*>(k));
    p.v = static_cast<const __nv_bfloat16*>(v);
    p.rows = rows, p.heads = heads, p.head_dim = head_dim, p.row_stride = row_stride;
    p.wq = wq, p.bq = bq, p.wk = wk, p.bk = bk, p.cos = cos, p.sin = sin;
    p.rows_per_batch = rows_per_batch, p.rope_row0 = rope_row0, p.rope_rows = rope_rows, p.eps = eps;
    for (int s = 0; s < nsplit; ++s) p.dst[s] = static_cast<__nv_bfloat16*>(dst[s]);
    p.nsplit = nsplit, p.dst_slot = dst_slot, p.slot_rows = slot_rows, p.dst_row0 = dst_row0;
    return launch_qk_norm_rope(p, mode, static_cast<cudaStream_t>(stream));
}

static int attention_fwd_plain(const char* who, int accumulate, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Lq, int Lkv, int D,
                               int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                               int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale, void* stream) {
    VAP_REQUIRE(q && k && v && o, "%s: null tensor", who);
    AttnParams p{};
    p.B = B, p.H = H, p.Lq = Lq, p.Lkv = Lkv;
    p.o = static_cast<__nv_bfloat16*>(o);
    p.o_sb = o_sb, p.o_sh = o_sh, p.o_sl = o_sl;
    p.lse = lse;
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.trace = g_attn_trace;
    p.accumulate = accumulate;
    const AttnTensor tq{static_cast<const __nv_bfloat16*>(q), q_sb, q_sh, q_sl};
    const AttnTensor tk{static_cast<const __nv_bfloat16*>(k), k_sb, k_sh, k_sl};
    const AttnTensor tv{static_cast<const __nv_bfloat16*>(v), v_sb, v_sh, v_sl};
    return launch_attention_fwd(tq, tk, tv, p, D, static_cast<cudaStream_t>(stream));
}

int vap_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Lq, int Lkv, int D, int64_t q_sb,
                      int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                      int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale, void* stream) {
    return attention_fwd_plain("vap_attention_fwd", 0, q, k, v, o, lse, B, H, Lq, Lkv, D, q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, o_sb, o_sh, o_sl,
                               scale, stream);
}

int vap_attention_fwd_accumulate(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Lq, int Lkv, int D, int64_t q_sb,
                                 int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                                 int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale, void* stream) {
    return attention_fwd_plain("vap_attention_fwd_accumulate", 1, q, k, v, o, lse, B, H, Lq, Lkv, D, q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, o_sb,
                               o_sh, o_sl, scale, stream);
}

int vap_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_peers, int npeers, int o_rows_per_peer, float* lse,
                              int B, int H, int Lq, int Lkv, int D, int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh,
                              int64_t k_sl, int64_t v_sb, int64_t v_sh, int64_t v_sl, int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale,
                              void* stream) {
    VAP_REQUIRE(q && k && v && o_peers, "vap_attention_fwd_scatter: null tensor");
    VAP_REQUIRE(npeers >= 1 && npeers <= 8 && o_rows_per_peer > 0 && static_cast<int64_t>(npeers) * o_rows_per_peer >= Lq,
                "vap_attention_fwd_scatter: %d peers x %d rows do not cover Lq=%d", npeers, o_rows_per_peer, Lq);
    AttnParams p{};
    p.B = B, p.H = H, p.Lq = Lq, p.Lkv = Lkv;
    p.o = nullptr;
    for (int r = 0; r < npeers; ++r) p.o_peer[r] = static_cast<__nv_bfloat16*>(o_peers[r]);
    p.o_rows_per_peer = o_rows_per_peer;
    p.o_sb = o_sb, p.o_sh = o_sh, p.o_sl = o_sl;
    p.lse = lse;
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.trace = g_attn_trace;
    const AttnTensor tq{static_cast<const __nv_bfloat16*>(q), q_sb, q_sh, q_sl};
    const AttnTensor tk{static_cast<const __nv_bfloat16*>(k), k_sb, k_sh, k_sl};
    const AttnTensor tv{static_cast<const __nv_bfloat16*>(v), v_sb, v_sh, v_sl};
    return launch_attention_fwd(tq, tk, tv, p, D, static_cast<cudaStream_t>(stream));
}

int vap_attention_fwd_splitkv(const void* q, const void* k, const void* v, void* o_part, float* lse_part, int kv_splits, int B, int H, int Lq,
                              int Lkv, int D, int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb,
                              int64_t v_sh, int64_t v_sl, float scale, void* stream) {
    VAP_REQUIRE(q && k && v && o_part && lse_part, "vap_attention_fwd_splitkv: null tensor");
    VAP_REQUIRE(kv_splits >= 1 && kv_splits <= 8, "vap_attention_fwd_splitkv: kv_splits=%d must be in [1, 8]", kv_splits);
    AttnParams p{};
    p.B = B, p.H = H, p.Lq = Lq, p.Lkv = Lkv;
    p.o = static_cast<__nv_bfloat16*>(o_part);  // [kv_splits, B, Lq, H, D]
    p.o_sl = static_cast<int64_t>(H) * D, p.o_sh = D, p.o_sb = static_cast<int64_t>(Lq) * H * D;
    p.kv_splits = kv_splits;
    p.o_split_stride = static_cast<int64_t>(B) * Lq * H * D;
    p.lse = lse_part;  // [kv_splits, B, H, Lq]
    p.lse_split_stride = static_cast<int64_t>(B) * H * Lq;
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.trace = g_attn_trace;
    const AttnTensor tq{static_cast<const __nv_bfloat16*>(q), q_sb, q_sh, q_sl};
    const AttnTensor tk{static_cast<const __nv_bfloat16*>(k), k_sb, k_sh, k_sl};
    const AttnTensor tv{static_cast<const __nv_bfloat16*>(v), v_sb, v_sh, v_sl};
    return launch_attention_fwd(tq, tk, tv, p, D, static_cast<cudaStream_t>(stream));
}

int vap_attention_combine(const void* o_part, const float* lse_part, int kv_splits, int B, int H, int Lq, int D, void* o, void* const* o_peers,
                          int npeers, int o_rows_per_peer, float* lse, int64_t o_sb, int64_t o_sh, int64_t o_sl, void* stream) {
    VAP_REQUIRE(o_part && lse_part, "vap_attention_combine: null partials");
    VAP_REQUIRE((o != nullptr) != (o_peers != nullptr), "vap_attention_combine: pass either o or o_peers");
    AttnCombineParams p{};
    p.o_part = static_cast<const __nv_bfloat16*>(o_part);
    p.lse_part = lse_part;
    p.splits = kv_splits, p.D = D;
    p.dst.B = B, p.dst.H = H, p.dst.Lq = Lq;
    p.dst.o = static_cast<__nv_bfloat16*>(o);
    if (o_peers) {
        VAP_REQUIRE(npeers >= 1 && npeers <= 8 && o_rows_per_peer > 0 && static_cast<int64_t>(npeers) * o_rows_per_peer >= Lq,
                    "vap_attention_combine: %d peers x %d rows do not cover Lq=%d", npeers, o_rows_per_peer, Lq);
        for (int r = 0; r < npeers; ++r) p.dst.o_peer[r] = static_cast<__nv_bfloat16*>(o_peers[r]);
        p.dst.o_rows_per_peer = o_rows_per_peer;
    }
    p.dst.o_sb = o_sb, p.dst.o_sh = o_sh, p.dst.o_sl = o_sl;
    p.dst.lse = lse;
    return launch_attention_combine(p, static_cast<cudaStream_t>(stream));
}

int vap_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K, const void* bias,
                  int epilogue, const void* R, int64_t ldr, const float* gate, int64_t gate_stride, int64_t rows_per_batch, void* stream) {
    VAP_REQUIRE(A && W && C, "vap_gemm_bf16: null tensor");
    GemmParams p{};
    p.M = M, p.N = N, p.K = K;
    p.C = static_cast<__nv_bfloat16*>(C);
    p.ldc = ldc;
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.epilogue = epilogue;
    p.R = static_cast<const __nv_bfloat16*>(R);
    p.ldr = ldr;
    p.gate = gate;
    p.gate_stride = gate_stride;
    p.rows_per_batch = rows_per_batch;
    return launch_gemm_bf16(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), ldw, p,
                            static_cast<cudaStream_t>(stream));
}

int vap_ulysses_pack(const void* src, void* dst, int64_t L, int nsplit, int64_t chunk, int64_t src_row_stride, int64_t dst_row_stride,
                     int64_t dst_split_stride, void* stream) {
    VAP_REQUIRE(src && dst, "vap_ulysses_pack: null tensor");
    return launch_ulysses(src, dst, L, nsplit, chunk, src_row_stride, dst_row_stride, dst_split_stride, 0, static_cast<cudaStream_t>(stream));
}
int vap_ulysses_unpack(const void* src, void* dst, int64_t L, int nsplit, int64_t chunk, int64_t src_row_stride, int64_t src_split_stride,
                       int64_t dst_row_stride, void* stream) {
    VAP_REQUIRE(src && dst, "vap_ulysses_unpack: null tensor");
    return launch_ulysses(src, dst, L, nsplit, chunk, dst_row_stride, src_row_stride, src_split_stride, 1, static_cast<cudaStream_t>(stream));
}

int vap_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                      float* delta_ws, int B, int H, int Lq, int Lkv, int D, const int64_t* strides, float scale, void* stream) {
    VAP_REQUIRE(q && k && v && o && dout && lse && dq && dk && dv && delta_ws && strides, "vap_attention_bwd: null argument");
    AttnBwdArgs a{};
    a.B = B, a.H = H, a.Lq = Lq, a.Lkv = Lkv;
    const int64_t* s = strides;  // (batch, head, token) element strides of q, k, v, o, dout, dq, dk, dv in that order
    a.q = AttnTensor{static_cast<const __nv_bfloat16*>(q), s[0], s[1], s[2]};
    a.k = AttnTensor{static_cast<const __nv_bfloat16*>(k), s[3], s[4], s[5]};
    a.v = AttnTensor{static_cast<const __nv_bfloat16*>(v), s[6], s[7], s[8]};
    a.o = AttnTensor{static_cast<const __nv_bfloat16*>(o), s[9], s[10], s[11]};
    a.dout = AttnTensor{static_cast<const __nv_bfloat16*>(dout), s[12], s[13], s[14]};
    a.dq = AttnGrad{static_cast<__nv_bfloat16*>(dq), s[15], s[16], s[17]};
    a.dk = AttnGrad{static_cast<__nv_bfloat16*>(dk), s[18], s[19], s[20]};
    a.dv = AttnGrad{static_cast<__nv_bfloat16*>(dv), s[21], s[22], s[23]};
    a.lse = lse, a.delta = delta_ws;
    a.scale = scale;
    return launch_attention_bwd(a, D, static_cast<cudaStream_t>(stream));
}

int vap_cfg_flow_match_step(const void* noise_cond, const void* noise_uncond, const void* sample, int sample_is_f32, void* out, int64_t batch,
                            int64_t inner, int64_t out_batch_stride, float guidance_scale, float dt, void* stream) {
    VAP_REQUIRE(noise_cond && sample && out, "vap_cfg_flow_match_step: null tensor");
    StepParams p{};
    p.cond = static_cast<const __nv_bfloat16*>(noise_cond), p.uncond = static_cast<const __nv_bfloat16*>(noise_uncond);
    p.sample = sample, p.sample_is_f32 = sample_is_f32;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.batch = batch, p.inner = inner, p.out_batch_stride = out_batch_stride;
    p.guidance = guidance_scale, p.dt = dt;
    return launch_cfg_flow_match_step(p, static_cast<cudaStream_t>(stream));
}

int vap_wan_modulation(const void* table, int table_is_f32, const void* temb, int temb_is_f32, float* out, int64_t batch, int chunks, int d,
                       int plus_one_mask, void* stream) {
    VAP_REQUIRE(table && temb && out, "vap_wan_modulation: null tensor");
    return launch_wan_modulation(table, table_is_f32, temb, temb_is_f32, out, batch, chunks, d, static_cast<unsigned>(plus_one_mask),
                                 static_cast<cudaStream_t>(stream));
}

int vap_debug_set_attention_trace(void* device_buffer) {
    g_attn_trace = static_cast<long long*>(device_buffer);
    return 0;
}

int vap_probe_umma(const void* A, const void* B, float* Dout, int N, int K, int a_in_tmem, int b_mn_major, int lbo_b, int sbo_b,
                   int kstep_b, int layout_type, void* stream) {
    VAP_REQUIRE(A && B && Dout, "vap_probe_umma: null tensor");
    ProbeParams p{};
    p.A = static_cast<const __nv_bfloat16*>(A);
    p.Dout = Dout;
    p.N = N, p.K = K;
    p.a_in_tmem = a_in_tmem & 1, p.b_mn_major = b_mn_major;
    p.lane16_shapes = (a_in_tmem >> 1) & 1;
    // defaults = the encodings gemm_sm100.cu / attn_sm100.cu use
    p.lbo_b = lbo_b >= 0 ? lbo_b : (b_mn_major ? K * 128 : 0);
    p.sbo_b = sbo_b >= 0 ? sbo_b : 1024;
    p.kstep_b = kstep_b >= 0 ? kstep_b : 2048;
    p.layout_type = layout_type >= 0 ? layout_type : kLayoutSw128;
    return launch_probe_umma(static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(B), p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

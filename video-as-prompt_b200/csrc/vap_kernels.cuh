// vap_kernels.cuh — parameter structs and launcher declarations shared by the kernel translation units and capi.cu.
#pragma once
#include "vap_common.cuh"

namespace vap {

struct LnParams {
    const __nv_bfloat16* x;
    __nv_bfloat16* out;
    int64_t rows;
    int d;
    int64_t x_stride, out_stride;  // elements
    const float* ln_w;             // [d] or null
    const float* ln_b;             // [d] or null
    const float* scale1p;          // [(nbatch), d] (1 + scale) or null
    const float* shift;            // [(nbatch), d] or null
    int64_t mod_stride;            // elements between batches' modulation vectors
    int64_t rows_per_batch;
    float eps;
    int cog_rounding;  // 1: round to bf16 after LN-affine, after the scale multiply and after the shift add
};
int launch_adaln_layernorm(const LnParams& p, cudaStream_t stream);

struct QkParams {
    __nv_bfloat16* q;
    __nv_bfloat16* k;
    int64_t rows;  // B * L
    int heads, head_dim;
    int64_t row_stride;  // elements, same for q and k
    const float* wq;     // Wan: [H*D]; Cog: [D]
    const float* bq;     // Cog only [D]
    const float* wk;
    const float* bk;
    const float* cos;  // [rope_rows, D/2] fp32 (pair i: cos[i], sin[i]); null -> no RoPE
    const float* sin;
    int64_t rows_per_batch;  // L
    int64_t rope_row0;       // positions < rope_row0 inside a batch are not rotated (Cog text tokens)
    int64_t rope_rows;       // rows in the table
    float eps;
    // --- fused all-to-all dispatch (Ulysses exchange #1 by peer stores); nsplit == 0: plain in-place mode ---
    const __nv_bfloat16* v;    // third item per row: copied unchanged (the V columns of the QKV projection)
    __nv_bfloat16* dst[8];     // dst[s]: rank s's receive buffer [slots, slot_rows, 3, (heads/nsplit)*head_dim]
    int nsplit;                // ranks
    int64_t dst_slot;          // slot (= this rank) inside every receive buffer
    int64_t slot_rows;         // rows per slot (local rows of both streams)
    int64_t dst_row0;          // first row of this call inside the slot (stream offset)
};
int launch_qk_norm_rope(const QkParams& p, int cog_mode, cudaStream_t stream);

struct AttnParams {
    int B, H, Lq, Lkv;
    __nv_bfloat16* o;
    int64_t o_sb, o_sh, o_sl;  // element strides of O
    float* lse;                // [B, H, Lq] or null
    float scale_log2;          // softmax scale * log2(e)
    float scale;
    long long* trace;          // optional clock64() trace of CTA (0,0,0): [3 roles][64 iterations][8 stamps], or null
    // --- fused all-to-all combine (Ulysses exchange #2 by peer stores); o_rows_per_peer == 0: plain mode (o above) ---
    __nv_bfloat16* o_peer[8];  // query rows [r*rpp, (r+1)*rpp) belong to rank r and are written to o_peer[r] + (row - r*rpp)*o_sl
    int o_rows_per_peer;
    // --- split-KV (kv_splits > 1): grid z = batch * kv_splits + split; split s attends to its share of the KV tiles and writes a
    //     normalised partial O at o + s * o_split_stride and its log-sum-exp at lse + s * lse_split_stride (lse is then required);
    //     launch_attention_combine merges the partials.  Fills the SMs when B * H * ceil(Lq / 256) is a poor multiple of their number.
    int kv_splits;
    int64_t o_split_stride, lse_split_stride;  // elements
    // --- accumulate != 0 (plain mode only): o <- bf16(float(o) + float(bf16(O / l))), the bf16 tensor add the reference performs on the
    //     outputs of its two cross-attention softmaxes (transformer_wan_mot.py:186), fused into the second launch's epilogue
    int accumulate;
};
struct AttnTensor {
    const __nv_bfloat16* ptr;
    int64_t sb, sh, sl;  // element strides (batch, head, token); the head_dim axis is contiguous
};
int launch_attention_fwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, AttnParams p, int D, cudaStream_t stream);

// Attention backward (attn_bwd_sm100.cu): dQ, dK, dV from dO, O and the forward's log-sum-exp.
struct AttnGrad {
    __nv_bfloat16* ptr;
    int64_t sb, sh, sl;
};
struct AttnBwdArgs {
    int B, H, Lq, Lkv;
    AttnTensor q, k, v, o, dout;  // q / o / dout [B,H,Lq,D], k / v [B,H,Lkv,D] by element strides
    const float* lse;             // [B, H, Lq] natural-log LSE of the scaled scores (vap_attention_fwd's optional output)
    float* delta;                 // [B, H, Lq] workspace: rowsum(dO o O)
    AttnGrad dq, dk, dv;
    float scale;
};
int launch_attention_bwd(const AttnBwdArgs& a, int D, cudaStream_t stream);

// Merge of split-KV partials: O = sum_s w_s O_s, w_s = exp(lse_s - lse) ; the result goes to `dst` exactly as the attention
// epilogue would have written it (plain strided tensor or rows scattered to their owning peers; dst.lse optional).
struct AttnCombineParams {
    const __nv_bfloat16* o_part;  // [splits, B, Lq, H, D] contiguous
    const float* lse_part;        // [splits, B, H, Lq]
    int splits, D;
    AttnParams dst;               // B, H, Lq, o / o_peer / o_rows_per_peer / strides / lse of the destination
};
int launch_attention_combine(const AttnCombineParams& p, cudaStream_t stream);

struct GemmParams {
    int M, N, K;
    __nv_bfloat16* C;
    int64_t ldc;
    const __nv_bfloat16* bias;  // [N] or null
    int epilogue;
    const __nv_bfloat16* R;  // residual [M, ldr]
    int64_t ldr;
    const float* gate;  // [(nbatch), N] fp32
    int64_t gate_stride;
    int64_t rows_per_batch;
    int m_blocks, n_blocks;
};
int launch_gemm_bf16(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, GemmParams p, cudaStream_t stream);

// CFG + FlowMatchEuler update (one elementwise pass over the latents)
struct StepParams {
    const __nv_bfloat16* cond;    // [batch, inner] noise prediction of the conditional pass
    const __nv_bfloat16* uncond;  // same shape, or null (no classifier-free guidance)
    const void* sample;           // [batch, inner] fp32 or bf16 latents
    int sample_is_f32;
    __nv_bfloat16* out;           // batch rows of `inner` elements, out_batch_stride apart
    int64_t batch, inner, out_batch_stride;
    float guidance, dt;
};
int launch_cfg_flow_match_step(const StepParams& p, cudaStream_t stream);

// W1: out[b, c, :] = float(table[c, :]) + float(temb[b, c, :]) (+ 1 for the chunks in plus_one_mask), fp32
int launch_wan_modulation(const void* table, int table_is_f32, const void* temb, int temb_is_f32, float* out, int64_t batch, int chunks, int d,
                          unsigned plus_one_mask, cudaStream_t stream);

struct ProbeParams {
    const __nv_bfloat16* A;  // [128, K] row-major (used directly when a_in_tmem)
    float* Dout;             // [128, N]
    int N, K;
    int a_in_tmem, b_mn_major;
    int lane16_shapes;  // stage A / read D back with the 16-lane TMEM shapes (16x128b st, 16x256b ld) the attention softmax uses
    uint32_t lbo_b, sbo_b, kstep_b, layout_type;
};
int launch_probe_umma(const __nv_bfloat16* A, const __nv_bfloat16* B, ProbeParams p, cudaStream_t stream);

}  // namespace vap

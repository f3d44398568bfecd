/* vap_b200.h — C ABI of libvap_b200.so: the B200 (sm_100a) kernels of the Video-As-Prompt MoT denoise hot path.
 *
 * The reference (bytedance/Video-As-Prompt) is pure Python/PyTorch and has no FFI of its own; every entry
 * point below replaces one PyTorch library call site (or a fused group of them) inside the reference's MoT
 * block forward.  "Ref:" cites the replaced interface, paths relative to
 * /root/reference/diffusers/src/diffusers.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch allocates every buffer incl. outputs);
 *     the library never allocates, frees or synchronises a stream.  The only state it keeps between calls is per-device
 *     bookkeeping that does not change results: which devices have been opted in to the kernels' dynamic shared memory
 *     (cudaFuncSetAttribute is a per-device setting), the SM count / cluster occupancy of each device, and the driver's
 *     tensor-map entry point.  Developer switches (VAP_ATTN_*, VAP_GEMM_* environment variables) select among kernels
 *     that compute the same function; a process may drive several GPUs (set the device before the call);
 *   - bf16 tensors are `uint16_t`-sized elements (torch.bfloat16); vectors of per-channel parameters
 *     (norm weights, modulation, gates, RoPE tables) are fp32;
 *   - sizes/strides are in ELEMENTS; the innermost dimension is contiguous;
 *   - `stream` is a `cudaStream_t` passed as `void*` (torch.cuda.current_stream().cuda_stream);
 *   - return 0 on success, <0 on error (-1 invalid argument, -2 CUDA runtime error, -3 TMA descriptor
 *     error); `vap_last_error()` returns the thread-local message.  Nothing is ever computed on the CPU.
 */
#ifndef VAP_B200_H
#define VAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAP_B200_VERSION 400 /* major*10000 + minor*100 + patch */

/* Library version (VAP_B200_VERSION of the build). */
int vap_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* vap_last_error(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int vap_sm_count(void);

/* (1) adaLN-modulated LayerNorm over the last dim, one pass (read x once, write out once).
 *   y = LayerNorm(x, eps)  [* ln_w + ln_b]  [* scale1p[b] + shift[b]]     b = row / rows_per_batch
 * rounding = 0 (Wan): everything in fp32, one bf16 rounding at the end.
 *     Ref: transformer_wan_mot.py:620-623, 680-689 (FP32LayerNorm no-affine + (1+scale)/shift modulation),
 *          :668-669 (norm2: FP32LayerNorm affine, no modulation), normalization.py:85-94.
 * rounding = 1 (CogVideoX): bf16 roundings after the affine LayerNorm, after the scale multiply and after
 *     the shift add, like the reference's bf16 tensor ops.
 *     Ref: CogVideoXLayerNormZero.forward, normalization.py:464-471; nn.LayerNorm norm_final :1045.
 * scale1p is (1 + scale), precomputed by the caller in the reference's dtype. */
int vap_adaln_layernorm(const void* x, void* out, int64_t rows, int d, int64_t x_row_stride, int64_t out_row_stride,
                        const float* ln_w, const float* ln_b, const float* scale1p, const float* shift, int64_t mod_stride,
                        int64_t rows_per_batch, float eps, int rounding, void* stream);

/* (2) q/k normalisation + rotary embedding, in place on q and k (rows = B * rows_per_batch tokens, each row
 *     heads*head_dim contiguous channels, consecutive rows row_stride apart — q and k may be column slices
 *     of the fused QKV projection output).
 * mode = 0 (Wan): RMSNorm over ALL heads*head_dim channels with weight wq/wk [heads*head_dim] (bq/bk NULL);
 *     Ref: WanAttnMOTProcessor2_0 transformer_wan_mot.py:218-236, RMSNorm normalization.py:554-568.
 * mode = 1 (CogVideoX): per-head LayerNorm(head_dim) with weight/bias [head_dim];
 *     Ref: CogVideoXAttnMOTProcessor2_0 attention_processor.py:2934-2945, apply_rotary_emb embeddings.py:1229-1248.
 * RoPE: pairs (x[2i], x[2i+1]) of every head are rotated by (cos[t, i], sin[t, i]), tables [rope_rows, head_dim/2]
 *     fp32, t = (row % rows_per_batch) - rope_row0; tokens with t < 0 (CogVideoX text tokens) are not
 *     rotated.  cos = sin = NULL disables RoPE; k = NULL normalises q only (cross-attention queries / keys). */
int vap_qk_norm_rope(void* q, void* k, int64_t rows, int heads, int head_dim, int64_t row_stride, const float* wq,
                     const float* bq, const float* wk, const float* bk, const float* cos, const float* sin,
                     int64_t rows_per_batch, int64_t rope_row0, int64_t rope_rows, float eps, int mode, void* stream);

/* (3) Joint attention forward: O = softmax(Q K^T * scale) V, no mask, no dropout, non-causal.
 *     q [B,H,Lq,D], k/v [B,H,Lkv,D], o [B,H,Lq,D] addressed by element strides (batch, head, token), D
 *     contiguous, D in {64, 128}; lse (optional, may be NULL) [B,H,Lq] fp32 = log-sum-exp of the scaled scores.
 *     Ref: F.scaled_dot_product_attention at transformer_wan_mot.py:637-644 (joint [target|ref]),
 *          cogvideox_transformer_3d_mot.py:424-431 (joint [text|target|text_ref|ref]),
 *          transformer_wan_mot.py:163-179 (cross-attention); finetrainers attention_dispatch.py:416-458. */
int vap_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Lq, int Lkv, int D,
                      int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb,
                      int64_t v_sh, int64_t v_sl, int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale, void* stream);

/* (3a) vap_attention_fwd whose epilogue ADDS to the output already in `o`:  o <- bf16(float(o) + float(bf16(softmax(QK^T) V))) —
 *     the bf16 tensor add of the reference's two cross-attention softmaxes (image tokens + text tokens), fused into the second
 *     launch.  Same arguments as vap_attention_fwd.
 *     Ref: transformer_wan_mot.py:163-186 (hidden_states = hidden_states_img + hidden_states). */
int vap_attention_fwd_accumulate(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Lq, int Lkv, int D,
                      int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb,
                      int64_t v_sh, int64_t v_sl, int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale, void* stream);

/* (4) Linear layer  C[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias[N])  (bf16 in/out, fp32 accumulate).
 *     Ref: nn.Linear call sites transformer_wan_mot.py:214-216, 241-243; attention.py:1245-1251 (FeedForward);
 *          attention_processor.py:2923-2925, 2952.
 * epilogue: 0 bias only;
 *           1 bias + GELU(tanh)                                                   (activations.py:65-91)
 *           2 C = bf16(R + bf16(AW+b) * gate)   fp32 gate, fp32 math            (transformer_wan_mot.py:658-663, 684, 693-697)
 *           3 C = bf16(R + bf16(AW+b))                                            (transformer_wan_mot.py:675-676)
 *           4 C = bf16(R + bf16(gate * bf16(AW+b)))  bf16-valued gate             (cogvideox_transformer_3d_mot.py:445-446, 457-458)
 * gate [nbatch, N] fp32 with gate_stride elements between batches, batch = row / rows_per_batch. */
int vap_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                  const void* bias, int epilogue, const void* R, int64_t ldr, const float* gate, int64_t gate_stride,
                  int64_t rows_per_batch, void* stream);

/* (5) Ulysses sequence-parallel re-layout helpers (head <-> sequence exchange around the joint attention;
 *     replaces the reference's ring attention, finetrainers attention_dispatch.py:686-773, parallel/ptd.py:515-679).
 *   pack  : dst[s][l][0:chunk] = src[l][s*chunk : (s+1)*chunk]      s < nsplit, l < L
 *           src rows are src_row_stride apart; dst rows dst_row_stride apart, dst splits dst_split_stride apart
 *           (builds the all-to-all #1 send buffer [P, L_loc, 3, H/P*D] from q, k, v column slices of the QKV output)
 *   unpack: dst[l][s*chunk : (s+1)*chunk] = src[s][l][0:chunk]      (after all-to-all #2: [P, L_loc, H/P*D] -> [L_loc, H*D])
 *   chunk and all strides are in bf16 elements and must be multiples of 8. */
int vap_ulysses_pack(const void* src, void* dst, int64_t L, int nsplit, int64_t chunk, int64_t src_row_stride,
                     int64_t dst_row_stride, int64_t dst_split_stride, void* stream);
int vap_ulysses_unpack(const void* src, void* dst, int64_t L, int nsplit, int64_t chunk, int64_t src_row_stride,
                       int64_t src_split_stride, int64_t dst_row_stride, void* stream);

/* (5b) Ulysses exchange FUSED into the producing kernels, over NVLink peer memory (the pointer tables hold device pointers of
 *      every rank's symmetric buffer, mapped into this process; `dst` / `o_peers` are HOST arrays of nsplit / npeers entries).
 *   vap_qkv_scatter: vap_qk_norm_rope (same arguments) on the q and k columns of the local QKV projection output plus an
 *      unchanged copy of the v columns, with every 16-byte result vector of head h stored straight into rank
 *      s = h / (heads/nsplit)'s receive buffer  dst[s][dst_slot][dst_row0 + row][q|k|v][(heads/nsplit)*head_dim]
 *      (slot_rows rows per slot) — all-to-all #1 without a pack kernel or a collective call.  q/k/v are not modified.
 *   vap_attention_fwd_scatter: vap_attention_fwd whose epilogue stores query row `row` into
 *      o_peers[row / o_rows_per_peer] at local row `row % o_rows_per_peer` (strides o_sb/o_sh/o_sl apply inside each peer
 *      buffer; the caller offsets every peer pointer to this rank's head columns) — all-to-all #2 and the unpack fused
 *      into the attention kernel.
 *   The caller orders the exchange with a device-side barrier between the ranks (e.g. torch symmetric-memory barrier). */
int vap_qkv_scatter(const void* q, const void* k, const void* v, int64_t rows, int heads, int head_dim, int64_t row_stride, const float* wq,
                    const float* bq, const float* wk, const float* bk, const float* cos, const float* sin, int64_t rows_per_batch,
                    int64_t rope_row0, int64_t rope_rows, float eps, int mode, void* const* dst, int nsplit, int64_t dst_slot,
                    int64_t slot_rows, int64_t dst_row0, void* stream);
int vap_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_peers, int npeers, int o_rows_per_peer, float* lse,
                              int B, int H, int Lq, int Lkv, int D, int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh,
                              int64_t k_sl, int64_t v_sb, int64_t v_sh, int64_t v_sl, int64_t o_sb, int64_t o_sh, int64_t o_sl, float scale,
                              void* stream);

/* (3b) Split-KV joint attention: when B * H * ceil(Lq / 256) work items are a poor multiple of the SM count (e.g. 5 heads per
 *      rank under 8-way Ulysses: 795 items = 5.4 waves on 148 SMs), the KV sequence is cut into kv_splits ranges; item
 *      (batch, head, 256 query rows, range) writes a NORMALISED partial O and its log-sum-exp, and the combine kernel merges
 *      them: O = sum_s exp(lse_s - lse) O_s — to a plain strided output (o) or, for the fused Ulysses exchange #2, straight into
 *      the owning ranks' buffers (o_peers / o_rows_per_peer as in vap_attention_fwd_scatter; pass exactly one of o / o_peers).
 *      o_part [kv_splits, B, Lq, H, D] bf16 and lse_part [kv_splits, B, H, Lq] fp32 are caller-allocated workspaces.
 *      Same reference call sites as (3). */
int vap_attention_fwd_splitkv(const void* q, const void* k, const void* v, void* o_part, float* lse_part, int kv_splits, int B, int H, int Lq,
                              int Lkv, int D, int64_t q_sb, int64_t q_sh, int64_t q_sl, int64_t k_sb, int64_t k_sh, int64_t k_sl, int64_t v_sb,
                              int64_t v_sh, int64_t v_sl, float scale, void* stream);
int vap_attention_combine(const void* o_part, const float* lse_part, int kv_splits, int B, int H, int Lq, int D, void* o, void* const* o_peers,
                          int npeers, int o_rows_per_peer, float* lse, int64_t o_sb, int64_t o_sh, int64_t o_sl, void* stream);

/* (3c) Joint attention BACKWARD (written without GPU access at the end of round 1 — its parity check sits in
 *      tests/gpu_checks.py:CHECKS_PENDING until it has run on a B200): dq, dk, dv from dout, the forward's output o and its
 *      log-sum-exp lse [B,H,Lq] (the optional output of vap_attention_fwd).  Same layouts as (3): q / o / dout / dq [B,H,Lq,D],
 *      k / v / dk / dv [B,H,Lkv,D] addressed by element strides, D in {64, 128} contiguous.  `strides` is a HOST array of 24 int64:
 *      (batch, head, token) strides of q, k, v, o, dout, dq, dk, dv in that order.  delta_ws [B,H,Lq] fp32 is a caller-allocated
 *      workspace (rowsum(dout * o)).  Three launches on `stream`: delta, dQ (per 128 query rows, streams K/V), dK+dV (per 128 KV
 *      rows, streams Q/dO); deterministic, no atomics.
 *      Ref: the autograd of F.scaled_dot_product_attention at transformer_wan_mot.py:637-644 / cogvideox_transformer_3d_mot.py:424-431
 *      as the trainer runs it (finetrainers/trainer/sft_trainer/trainer.py:674-714, attention_dispatch.py:416-458). */
int vap_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                      float* delta_ws, int B, int H, int Lq, int Lkv, int D, const int64_t* strides, float scale, void* stream);

/* (8) Classifier-free guidance + FlowMatchEuler scheduler update of the Wan denoise loop, one elementwise pass (HBM-bound):
 *       n   = noise_uncond ? bf16(u + bf16(g * bf16(c - u))) : c     Ref: pipelines/wan/pipeline_wan_i2v_mot.py:874 (three bf16 tensor ops)
 *       out = bf16(float(sample) + bf16(dt * n))                       Ref: schedulers/scheduling_flow_match_euler_discrete.py:433, 457, 462-467
 *                                                                          (sample upcast to fp32; dt * model_output is a bf16 tensor op; cast back)
 *     noise_cond / noise_uncond [batch, inner] bf16 contiguous (noise_uncond may be NULL: no guidance); sample [batch, inner] contiguous,
 *     fp32 (sample_is_f32 = 1: the pipeline's initial latents) or bf16 (the later steps); out: batch rows of `inner` bf16 elements,
 *     out_batch_stride apart — e.g. the latent channels of the NEXT step's transformer input [B, 16 + 20, F, h, w], which saves the
 *     reference's torch.cat([latents, condition]) (:815).  dt: the value torch multiplies by — the scheduler's dt is a 0-dim fp32 TENSOR,
 *     which torch casts to the other operand's dtype (bf16) before the multiply, so the caller passes float(bf16(sigma_next - sigma));
 *     the kernel itself multiplies by whatever fp32 value it is given.  inner and out_batch_stride must be
 *     multiples of 8, all pointers 16-byte aligned. */
int vap_cfg_flow_match_step(const void* noise_cond, const void* noise_uncond, const void* sample, int sample_is_f32, void* out, int64_t batch,
                            int64_t inner, int64_t out_batch_stride, float guidance_scale, float dt, void* stream);

/* (5b) adaLN modulation vectors of a Wan block in one launch (W1):
 *   out[b, c, :] = float(table[c, :]) + float(temb[b, c, :])   (+ 1 when bit c of plus_one_mask is set)      fp32 [batch, chunks, d]
 * table [chunks, d] and temb [batch, chunks, d] are contiguous, bf16 (is_f32 = 0) or fp32 (1).  The six chunks are
 * shift / scale / gate / c_shift / c_scale / c_gate; mask 0b010010 yields (1 + scale) and (1 + c_scale) ready for
 * vap_adaln_layernorm's scale1p.
 *     Ref: transformer_wan_mot.py:606-616 ((scale_shift_table + temb.float()).chunk(6)), :620-622, :680-689 ("1 + scale"). */
int vap_wan_modulation(const void* table, int table_is_f32, const void* temb, int temb_is_f32, float* out, int64_t batch, int chunks, int d,
                       int plus_one_mask, void* stream);

/* (6) Bring-up probe for the tcgen05 descriptors: one CTA computes D[128,N] = A[128,K] * B, fp32 out.
 *     a_in_tmem: bit 0: 0 = A from shared memory (K-major, SWIZZLE_128B), 1 = A staged to TMEM as packed bf16;
 *                bit 1: stage A / read D back with the 16-lane TMEM shapes (tcgen05.st 16x128b, tcgen05.ld 16x256b).
 *     b_mn_major: 0 = B is [N,K] (K contiguous), 1 = B is [K,N] (N contiguous; the V operand of P*V).
 *     Descriptor fields are runtime arguments so alternative encodings can be swept from the test-suite. */
int vap_probe_umma(const void* A, const void* B, float* Dout, int N, int K, int a_in_tmem, int b_mn_major, int lbo_b, int sbo_b,
                   int kstep_b, int layout_type, void* stream);

/* (7) Debug: when set to a device buffer of 3*64*8 int64, CTA (0,0,0) of every following vap_attention_fwd records
 *     clock64() stamps of its softmax warps and MMA issuer (see tools/attn_trace.py).  NULL disables it. */
int vap_debug_set_attention_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* VAP_B200_H */

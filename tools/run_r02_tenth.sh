# three-stage P publish (attention) + two-CUDA-stream block schedule: parity first, then A/B timings.  One GPU.
set -x
ATT=attn_d128,attn_d128_multi_tile,attn_d64,attn_cross,attn_cross_512,attn_one_tile,attn_peaky,attn_accumulate,attn_accumulate_d64,attn_splitkv_2,attn_splitkv_3_d64,attn_splitkv_uneven,ulysses_p2p_emulated_wan,ulysses_p2p_emulated_cog,attn_full_size
for v in l16p3a l16p3b; do
  VAP_B200_LIB=$PWD/build_variants/libvap_$v.so timeout 500 python tools/gpu_diag.py --only $ATT --stop-on-fail --timeout 60 > gpurun_out/p3_checks_$v.log 2>&1; echo "$v checks rc=$?"; cut -c1-200 gpurun_out/p3_checks_$v.log | tail -18
done
# dual-stream schedule (default on) through the block / model / denoise / reference-class checks, with the in-tree library
timeout 600 python tools/gpu_diag.py --only wan_blocks,wan_model,wan_denoise,cog_blocks_small,cog_blocks_multi,cog_model_config1,processor_level,wan_denoise_cached,wan_dead_ref_skip --timeout 120 > gpurun_out/dual_checks.log 2>&1; echo "dual checks rc=$?"; cut -c1-220 gpurun_out/dual_checks.log | tail -12
timeout 400 python tools/attn_ab.py --rounds 3 --shapes wan,cog --modes lane16 --only intree,l16old,l16p3a,l16p3b,l16p3spin,l16p3poly2 > gpurun_out/attn_ab7.json 2> gpurun_out/attn_ab7.err; echo "attn_ab rc=$?"; tail -c 2500 gpurun_out/attn_ab7.json; tail -3 gpurun_out/attn_ab7.err
timeout 400 python tools/dual_stream_ab.py --iters 8 > gpurun_out/dual_stream_ab.json 2> gpurun_out/dual_stream_ab.err; echo "dual_ab rc=$?"; cat gpurun_out/dual_stream_ab.json; tail -5 gpurun_out/dual_stream_ab.err

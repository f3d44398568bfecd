# Round 2, third GPU call: the one-thread-per-row attention kernel after the setmaxnreg fix (ONE small check first; everything that depends
# on it is skipped if that fails), A/B of its variants, then the staged norm kernels (parity, then bandwidth A/B) and a block-level regression.
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
if VAP_ATTN_SOFTMAX=row timeout 120 python tools/gpu_diag.py --only attn_one_tile --timeout 60 > gpurun_out/row_smoke.log 2>&1; then
  cut -c1-300 gpurun_out/row_smoke.log
  ATTN="attn_d128,attn_d128_multi_tile,attn_d64,attn_cross,attn_cross_512,attn_peaky,attn_splitkv_2,attn_splitkv_3_d64,attn_splitkv_uneven,attn_splitkv_peers,ulysses_p2p_emulated_wan,ulysses_p2p_emulated_cog,attn_full_size"
  VAP_ATTN_SOFTMAX=row python tools/gpu_diag.py --only $ATTN --stop-on-fail --timeout 90 > gpurun_out/row_checks.log 2>&1; echo "row checks rc=$?"; cut -c1-250 gpurun_out/row_checks.log
  cp gpurun_out/diag.json gpurun_out/diag_row.json
  VAP_ATTN_SOFTMAX=row VAP_ATTN_CLUSTER=2 python tools/gpu_diag.py --only attn_d128_multi_tile,attn_d64 --stop-on-fail --timeout 90 > gpurun_out/row_cl2_checks.log 2>&1; echo "row cl2 checks rc=$?"; cut -c1-250 gpurun_out/row_cl2_checks.log
  timeout 400 python tools/attn_ab.py --rounds 2 --shapes wan,cog > gpurun_out/attn_ab.json 2> gpurun_out/attn_ab.err; echo "attn_ab rc=$?"; tail -c 4500 gpurun_out/attn_ab.json; tail -5 gpurun_out/attn_ab.err
else
  echo "ROW SMOKE FAILED"; tail -c 1500 gpurun_out/row_smoke.log
fi
python tools/gpu_diag.py --only wan_modulation,ln_wan_staged,ln_wan_staged_affine,ln_wan_staged_batch2,ln_cog_staged,qk_wan_staged --stop-on-fail --timeout 90 > gpurun_out/staged_checks.log 2>&1; rc=$?; echo "staged checks rc=$rc"; cut -c1-300 gpurun_out/staged_checks.log
if [ $rc -eq 0 ]; then
  timeout 300 python tools/kernel_bench.py --mem > gpurun_out/kernel_bench_mem.log 2>&1; cat gpurun_out/kernel_bench_mem.log | cut -c1-400
  python tools/gpu_diag.py --only ln_wan,ln_wan_batch,ln_cog,qk_wan,qk_cog,wan_blocks,wan_model,wan_denoise,cog_blocks_small,cog_model_config1 --timeout 120 > gpurun_out/regress_checks.log 2>&1; echo "regress rc=$?"; cut -c1-300 gpurun_out/regress_checks.log
fi

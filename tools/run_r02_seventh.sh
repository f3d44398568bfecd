set -x
if timeout 300 python tools/gpu_diag.py --only attn_d64,attn_splitkv_3_d64,ulysses_p2p_emulated_cog,attn_d128 --stop-on-fail --timeout 40 > gpurun_out/dec_smoke.log 2>&1; then
  cut -c1-200 gpurun_out/dec_smoke.log
  VAP_ATTN_CLUSTER=2 timeout 100 python tools/gpu_diag.py --only attn_d64 --stop-on-fail --timeout 40 > gpurun_out/dec_cl2.log 2>&1; echo "cl2 rc=$?"; cut -c1-200 gpurun_out/dec_cl2.log
  timeout 400 python tools/attn_ab.py --rounds 3 --shapes cog,wan > gpurun_out/attn_ab5.json 2> gpurun_out/attn_ab5.err; echo "attn_ab rc=$?"; tail -c 1200 gpurun_out/attn_ab5.json; tail -5 gpurun_out/attn_ab5.err
  python tools/gpu_diag.py --only cog_blocks_small,cog_blocks_multi,cog_model_config1,cog_denoise,qk_cog --timeout 120 > gpurun_out/cog_checks.log 2>&1; echo "cog rc=$?"; cut -c1-200 gpurun_out/cog_checks.log
else
  echo "DECOUPLED SMOKE FAILED"; tail -c 1500 gpurun_out/dec_smoke.log
fi

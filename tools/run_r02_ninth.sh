set -x
if VAP_ATTN_PAIR=1 timeout 200 python tools/gpu_diag.py --only attn_d128 --stop-on-fail --timeout 40 > gpurun_out/pair_smoke.log 2>&1; then
  cut -c1-200 gpurun_out/pair_smoke.log
  VAP_ATTN_PAIR=1 timeout 400 python tools/gpu_diag.py --only attn_d128_multi_tile,attn_cross,attn_cross_512,attn_peaky,attn_accumulate,attn_splitkv_2,attn_splitkv_uneven,attn_splitkv_peers,ulysses_p2p_emulated_wan,ulysses_p2p_emulated_p8,attn_full_size,attn_bwd_d128 --stop-on-fail --timeout 60 > gpurun_out/pair_checks.log 2>&1; echo "pair checks rc=$?"; cut -c1-220 gpurun_out/pair_checks.log
  timeout 300 python tools/attn_ab.py --rounds 3 --shapes wan,j16k > gpurun_out/attn_ab6.json 2> gpurun_out/attn_ab6.err; echo "attn_ab rc=$?"; tail -c 1500 gpurun_out/attn_ab6.json; tail -5 gpurun_out/attn_ab6.err
else
  echo "PAIR SMOKE FAILED"; tail -c 2500 gpurun_out/pair_smoke.log; cat gpurun_out/diag_attn_d128.log | tail -20
fi

# Round 2, fifth GPU call: new MMA issue order (bookkeeping off the P -> PV path) + softmax latency hiding: smoke, parity, trace, A/B; staged norm kernels v2.
set -x
if timeout 200 python tools/gpu_diag.py --only attn_one_tile,attn_d128_multi_tile,attn_d64,attn_cross --stop-on-fail --timeout 40 > gpurun_out/order_smoke.log 2>&1; then
  cut -c1-200 gpurun_out/order_smoke.log
  VAP_B200_LIB=$PWD/build_variants/libvap_l16ed.so timeout 200 python tools/gpu_diag.py --only attn_one_tile,attn_d128_multi_tile,attn_d64,attn_peaky,attn_cross --stop-on-fail --timeout 40 > gpurun_out/ed_smoke.log 2>&1; echo "ed smoke rc=$?"; cut -c1-200 gpurun_out/ed_smoke.log
  VAP_ATTN_SOFTMAX=row timeout 200 python tools/gpu_diag.py --only attn_one_tile,attn_d128_multi_tile,attn_d64 --stop-on-fail --timeout 40 > gpurun_out/row_smoke.log 2>&1; echo "row smoke rc=$?"; cut -c1-200 gpurun_out/row_smoke.log
  VAP_B200_LIB=$PWD/build_variants/libvap_l16trace.so timeout 120 python tools/attn_trace.py 128 > gpurun_out/attn_trace2_d128.txt 2>&1; sed -n 4,8p gpurun_out/attn_trace2_d128.txt; sed -n 15,19p gpurun_out/attn_trace2_d128.txt; sed -n 26,30p gpurun_out/attn_trace2_d128.txt
  VAP_B200_LIB=$PWD/build_variants/libvap_l16edtrace.so timeout 120 python tools/attn_trace.py 128 > gpurun_out/attn_trace3_d128.txt 2>&1; sed -n 4,8p gpurun_out/attn_trace3_d128.txt; sed -n 15,19p gpurun_out/attn_trace3_d128.txt; sed -n 26,30p gpurun_out/attn_trace3_d128.txt
  timeout 400 python tools/attn_ab.py --rounds 3 --shapes wan,cog > gpurun_out/attn_ab3.json 2> gpurun_out/attn_ab3.err; echo "attn_ab rc=$?"; tail -c 2600 gpurun_out/attn_ab3.json; tail -5 gpurun_out/attn_ab3.err
else
  echo "ORDER SMOKE FAILED"; tail -c 1500 gpurun_out/order_smoke.log
fi
python tools/gpu_diag.py --only ln_wan_staged,ln_wan_staged_affine,ln_wan_staged_batch2,ln_cog_staged,qk_wan_staged --stop-on-fail --timeout 60 > gpurun_out/staged_checks.log 2>&1; rc=$?; echo "staged checks rc=$rc"; cut -c1-200 gpurun_out/staged_checks.log
[ $rc -eq 0 ] && timeout 300 python tools/kernel_bench.py --mem > gpurun_out/kernel_bench_mem2.log 2>&1; cut -c1-330 gpurun_out/kernel_bench_mem2.log

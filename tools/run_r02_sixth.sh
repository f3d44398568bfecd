set -x
python tools/gpu_diag.py --only ln_wan_staged,ln_wan_staged_affine,ln_wan_staged_batch2,ln_cog_staged,qk_wan_staged --stop-on-fail --timeout 60 > gpurun_out/staged_checks.log 2>&1; rc=$?; echo "staged checks rc=$rc"; cut -c1-200 gpurun_out/staged_checks.log
[ $rc -eq 0 ] && timeout 300 python tools/kernel_bench.py --mem > gpurun_out/kernel_bench_mem3.log 2>&1; cut -c1-330 gpurun_out/kernel_bench_mem3.log
timeout 400 python tools/attn_ab.py --rounds 3 --shapes wan,cog > gpurun_out/attn_ab4.json 2> gpurun_out/attn_ab4.err; echo "attn_ab rc=$?"; tail -c 1200 gpurun_out/attn_ab4.json; tail -5 gpurun_out/attn_ab4.err

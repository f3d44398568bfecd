# First GPU call for the code written without GPU access (tests/gpu_checks.py:CHECKS_PENDING).  Ordered from harmless to risky, every
# check in its own subprocess (tools/gpu_diag.py); the attention-backward kernels start with the single-tile case and stop at the first
# failure so that a faulting kernel is launched once, not five times.  `nvidia-smi` Xid lines are printed after each stage.
set -x
xid() { dmesg 2>/dev/null | grep -i xid | tail -3; nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader; }
python tools/gpu_diag.py --pending --only wan_denoise_cached,wan_dead_ref_skip,cfg_flow_match_step,wan_denoise_fused; xid
for c in attn_bwd_one_tile attn_bwd_d128 attn_bwd_d64 attn_bwd_tails attn_bwd_multi_tile attn_bwd_split2_d128 attn_bwd_split2_d64 attn_bwd_prefetch_d128 attn_bwd_prefetch_d64; do
  python tools/gpu_diag.py --pending --only $c || { echo "STOP at $c"; cat gpurun_out/diag_$c.log | tail -20; xid; exit 1; }
  xid
done
timeout 300 python tools/kernel_bench.py --bwd --quick
VAP_ATTN_BWD_SPLIT=2 timeout 300 python tools/kernel_bench.py --bwd --quick
VAP_ATTN_BWD_SPLIT=2 VAP_ATTN_BWD_PREFETCH=1 timeout 300 python tools/kernel_bench.py --bwd --quick

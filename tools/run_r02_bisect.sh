set -x
for cfg in "A=1" "VAP_DUAL_STREAM=0" "VAP_ATTN_SHORT=0"; do
  env $cfg timeout 300 python tools/gpu_diag.py --reference --only ref_wan14b_blocks --timeout 200 > gpurun_out/bisect_$cfg.log 2>&1; echo "$cfg rc=$?"; grep -o "'per_block': {[^}]*}" gpurun_out/bisect_$cfg.log gpurun_out/diag_ref_wan14b_blocks.log 2>/dev/null | head -2; grep -o '"per_block": {[^}]*}' gpurun_out/bisect_$cfg.log | head -1
done
env A=1 timeout 300 python tools/gpu_diag.py --reference --only ref_wan14b_blocks --timeout 200 > gpurun_out/bisect_again.log 2>&1; echo "again rc=$?"; grep -o '"per_block": {[^}]*}' gpurun_out/bisect_again.log | head -1

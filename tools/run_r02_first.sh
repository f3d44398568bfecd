# Round 2, first GPU call: (1) first hardware run of CHECKS_PENDING, (2) the one-thread-per-row attention kernel: parity, then A/B of its
# variants against the 16-lane kernel and cuDNN, (3) parity + timing against the staged reference's own classes.
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
bash tools/run_pending_gpu.sh > gpurun_out/pending_r02.log 2>&1; echo "pending rc=$?"; grep -E "^\{|checks passed|STOP" gpurun_out/pending_r02.log | cut -c1-400
ATTN="attn_d128,attn_d128_multi_tile,attn_d64,attn_cross,attn_cross_512,attn_one_tile,attn_peaky,attn_splitkv_2,attn_splitkv_3_d64,attn_splitkv_uneven,attn_splitkv_peers,ulysses_p2p_emulated_wan,ulysses_p2p_emulated_cog,attn_full_size"
VAP_ATTN_SOFTMAX=row python tools/gpu_diag.py --only $ATTN > gpurun_out/row_checks.log 2>&1; echo "row checks rc=$?"; cut -c1-300 gpurun_out/row_checks.log
cp gpurun_out/diag.json gpurun_out/diag_row.json
VAP_ATTN_SOFTMAX=row VAP_ATTN_CLUSTER=2 python tools/gpu_diag.py --only attn_d128_multi_tile,attn_d64,attn_full_size > gpurun_out/row_cl2_checks.log 2>&1; echo "row cl2 checks rc=$?"; cut -c1-300 gpurun_out/row_cl2_checks.log
timeout 600 python tools/attn_ab.py --rounds 2 --shapes wan,cog > gpurun_out/attn_ab.json 2> gpurun_out/attn_ab.err; echo "attn_ab rc=$?"; tail -c 3000 gpurun_out/attn_ab.json; tail -5 gpurun_out/attn_ab.err
python tools/gpu_diag.py --reference --timeout 900 > gpurun_out/reference_r02.log 2>&1; echo "reference rc=$?"; cut -c1-1500 gpurun_out/reference_r02.log

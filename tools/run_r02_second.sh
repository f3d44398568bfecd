# Round 2, second GPU call: the one-thread-per-row attention kernel (parity, then A/B of its variants against the 16-lane kernel and
# cuDNN on the same box), then the bench line with its reference_gpu / cpu_baseline legs and the CPU reference arm.
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
ATTN="attn_d128,attn_d128_multi_tile,attn_d64,attn_cross,attn_cross_512,attn_one_tile,attn_peaky,attn_splitkv_2,attn_splitkv_3_d64,attn_splitkv_uneven,attn_splitkv_peers,ulysses_p2p_emulated_wan,ulysses_p2p_emulated_cog,attn_full_size"
VAP_ATTN_SOFTMAX=row python tools/gpu_diag.py --only $ATTN > gpurun_out/row_checks.log 2>&1; echo "row checks rc=$?"; cut -c1-300 gpurun_out/row_checks.log
cp gpurun_out/diag.json gpurun_out/diag_row.json
VAP_ATTN_SOFTMAX=row VAP_ATTN_CLUSTER=2 python tools/gpu_diag.py --only attn_d128_multi_tile,attn_d64,attn_full_size > gpurun_out/row_cl2_checks.log 2>&1; echo "row cl2 checks rc=$?"; cut -c1-300 gpurun_out/row_cl2_checks.log
timeout 600 python tools/attn_ab.py --rounds 2 --shapes wan,cog > gpurun_out/attn_ab.json 2> gpurun_out/attn_ab.err; echo "attn_ab rc=$?"; tail -c 3500 gpurun_out/attn_ab.json; tail -5 gpurun_out/attn_ab.err
python bench.py > gpurun_out/bench_r02a.log 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r02a.log | cut -c1-6000; tail -3 gpurun_out/bench_r02a.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r02a.log 2>&1; tail -1 gpurun_out/bench_ref_r02a.log | cut -c1-2500

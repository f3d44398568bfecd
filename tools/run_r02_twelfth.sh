# GEMM residual-epilogue prefetch (parity + A/B per epilogue) and the block-level trainer seam against the staged reference.  One GPU.
set -x
timeout 400 python tools/gpu_diag.py --only gemm_small,gemm_n128,gemm_tails,gemm_gelu,gemm_gate_f32,gemm_res_add,gemm_gate_bf16,gemm_nobias_k5120,gemm_large,wan_blocks,cog_blocks_small --stop-on-fail --timeout 90 > gpurun_out/gemm_checks.log 2>&1; echo "gemm checks rc=$?"; cut -c1-200 gpurun_out/gemm_checks.log | tail -13
timeout 500 python tools/gemm_epi_ab.py --rounds 3 > gpurun_out/gemm_epi_ab.json 2> gpurun_out/gemm_epi_ab.err; echo "gemm_epi_ab rc=$?"; cat gpurun_out/gemm_epi_ab.json | cut -c1-3000; tail -3 gpurun_out/gemm_epi_ab.err
timeout 700 python tools/gpu_diag.py --reference --only ref_wan14b_train_block,ref_cog5b_train_block --timeout 300 > gpurun_out/train_block_checks.log 2>&1; echo "train checks rc=$?"; cut -c1-600 gpurun_out/train_block_checks.log | tail -8

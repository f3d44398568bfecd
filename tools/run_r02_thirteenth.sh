# GEMM: eight epilogue warps (in-tree) vs four (build_variants/libvap_epi4.so): parity, then per-epilogue A/B.  One GPU.
set -x
timeout 500 python tools/gpu_diag.py --only gemm_small,gemm_n128,gemm_tails,gemm_gelu,gemm_gate_f32,gemm_res_add,gemm_gate_bf16,gemm_nobias_k5120,gemm_large,wan_blocks,wan_model,cog_blocks_small,cog_blocks_multi,cog_model_config1 --stop-on-fail --timeout 90 > gpurun_out/gemm_checks2.log 2>&1; echo "gemm checks rc=$?"; cut -c1-160 gpurun_out/gemm_checks2.log | tail -16
timeout 500 python tools/gemm_epi_ab.py --rounds 3 > gpurun_out/gemm_epi_ab.json 2> gpurun_out/gemm_epi_ab.err; echo "gemm_epi_ab rc=$?"; tail -3 gpurun_out/gemm_epi_ab.err
timeout 300 python tools/dual_stream_ab.py --iters 8 > gpurun_out/dual_stream_ab2.json 2> gpurun_out/dual_stream_ab2.err; echo "dual_ab rc=$?"; cat gpurun_out/dual_stream_ab2.json; tail -3 gpurun_out/dual_stream_ab2.err

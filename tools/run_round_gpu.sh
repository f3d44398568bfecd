# One round-end GPU pass: parity tests (incl. the staged reference's own classes), bench (own arm with its reference_gpu / cpu_baseline legs +
# the CPU reference arm), ncu launch list of one timed step (2-block model: per-block kernel shares are identical), ncu --set full of the
# joint-attention launch and of one CTA-pair GEMM launch inside that step, and of the attention micro-target with source-level stall sampling.
# Every ncu command runs only after the same command has exited 0 without ncu.
set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-1200
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
B2L="python bench.py --config wan14b_2l --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu --profile"
$B2L > gpurun_out/plain_bench2l.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B2L > gpurun_out/ncu_bench.log 2>&1
$B2L > gpurun_out/plain_bench2l_b.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/attn_step -f $B2L > gpurun_out/ncu_attn_step.log 2>&1
$B2L > gpurun_out/plain_bench2l_c.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_pair -c 1 -o gpurun_out/gemm_pair_step -f $B2L > gpurun_out/ncu_gemm_step.log 2>&1
$B2L > gpurun_out/plain_bench2l_d.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:adaln_layernorm_staged -c 1 -o gpurun_out/ln_step -f $B2L > gpurun_out/ncu_ln_step.log 2>&1
echo done

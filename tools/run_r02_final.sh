# end-of-round pass on one GPU: short-KV kernel A/B (after the early V release), full pytest -m gpu, smoke, the headline bench and the other workloads
set -x
timeout 200 python tools/attn_short_ab.py > gpurun_out/attn_short_ab2.json 2> gpurun_out/attn_short_ab2.err; echo "short ab rc=$?"; cat gpurun_out/attn_short_ab2.json | head -4; tail -2 gpurun_out/attn_short_ab2.err
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-900
python bench.py --config cog5b --no-cpu-baseline > gpurun_out/bench_cog5b_final.log 2> gpurun_out/bench_cog5b_final.err; echo "bench cog rc=$?"; tail -1 gpurun_out/bench_cog5b_final.log | cut -c1-700
python bench.py --config wan14b_d20 --no-cpu-baseline > gpurun_out/bench_d20_final.log 2> gpurun_out/bench_d20_final.err; echo "bench d20 rc=$?"; tail -1 gpurun_out/bench_d20_final.log | cut -c1-700

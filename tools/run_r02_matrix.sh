# Round 2 measurement matrix (one B200): every workload bench.py knows, driver-comparable settings, with the stock reference on the same
# box where it fits; kernel micro-benchmarks incl. the per-rank Ulysses shapes.  Lines land in gpurun_out/ and are copied to profiles/.
set -x
for cfg in wan14b_d20 wan14b_d10 cog5b; do
  python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.log 2> gpurun_out/bench_$cfg.err; echo "$cfg rc=$?"; tail -1 gpurun_out/bench_$cfg.log | cut -c1-700
done
python bench.py --config wan14b_720p --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_wan14b_720p.log 2> gpurun_out/bench_wan14b_720p.err; echo "720p rc=$?"; tail -1 gpurun_out/bench_wan14b_720p.log | cut -c1-700
timeout 600 python tools/kernel_bench.py --attn --gemm --mem > gpurun_out/kernel_bench_r02.log 2>&1; cp gpurun_out/kernel_bench.json gpurun_out/kernel_bench_r02.json; cut -c1-420 gpurun_out/kernel_bench_r02.log
timeout 600 python tools/kernel_bench.py --sp > gpurun_out/kernel_bench_sp_r02.log 2>&1; cp gpurun_out/kernel_bench.json gpurun_out/kernel_bench_sp_r02.json; cut -c1-330 gpurun_out/kernel_bench_sp_r02.log

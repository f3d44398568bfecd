set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/attn_trace.py > gpurun_out/attn_trace128.log 2>&1
python tools/attn_trace.py 64 > gpurun_out/attn_trace64.log 2>&1
python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
python tools/kernel_bench.py > gpurun_out/kernel_bench.json 2> gpurun_out/kernel_bench.err
python bench.py --config wan14b_2l --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench2l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --config wan14b_2l --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/r01_attn_v3 -f python tools/prof_target.py attn > gpurun_out/ncu_attn.log 2>&1
echo done

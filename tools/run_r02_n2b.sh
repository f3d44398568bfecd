# 2-GPU call after the two-stream schedule / GEMM epilogue changes: Ulysses parity on real GPUs (both transports, Wan + CogVideoX), the bench
# (eager + CUDA-graph replay of the step, which now captures the fork / join of the side stream) and the same bench with the one-stream schedule.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/ulysses_multi_gpu_check.py > gpurun_out/ulysses_check_wan_n2b.log 2>&1; echo "wan check rc=$?"; grep '^{' gpurun_out/ulysses_check_wan_n2b.log
timeout 300 $TR --master-port 29512 tools/ulysses_multi_gpu_check.py --family cog > gpurun_out/ulysses_check_cog_n2b.log 2>&1; echo "cog check rc=$?"; grep '^{' gpurun_out/ulysses_check_cog_n2b.log
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --graph on > gpurun_out/bench_n2b.log 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2b.log | cut -c1-2500; tail -3 gpurun_out/bench_n2b.err
VAP_DUAL_STREAM=0 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2b_single.log 2> gpurun_out/bench_n2b_single.err; echo "bench n2 one-stream rc=$?"; grep '^{' gpurun_out/bench_n2b_single.log | cut -c1-600; tail -3 gpurun_out/bench_n2b_single.err

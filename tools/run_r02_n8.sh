# 8-GPU call: Ulysses parity on eight real GPUs (both transports, incl. a B = 2 batch), the headline bench (eager + graph replay in one run, with
# parity_vs_single), and the 81f 720p workload of BASELINE.json configs[3].
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 tools/ulysses_multi_gpu_check.py --heads 8 > gpurun_out/ulysses_check_wan_n8.log 2>&1; echo "wan check rc=$?"; grep '^{' gpurun_out/ulysses_check_wan_n8.log
timeout 420 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --graph on > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; grep '^{' gpurun_out/bench_n8.log | cut -c1-2600; tail -3 gpurun_out/bench_n8.err | cut -c1-300
timeout 420 $TR --master-port 29523 bench.py --gpus 8 --config wan14b_720p --steps 2 --warmup 3 > gpurun_out/bench_n8_720p.log 2> gpurun_out/bench_n8_720p.err; echo "bench n8 720p rc=$?"; grep '^{' gpurun_out/bench_n8_720p.log | cut -c1-1500; tail -3 gpurun_out/bench_n8_720p.err | cut -c1-300
